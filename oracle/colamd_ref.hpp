// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the int32 COLAMD / SYMAMD column ordering used by the reference
// (`colamd_rs`, a transliteration of SuiteSparse COLAMD @9759b8c).  Every routine cites the
// reference lines it follows (paths relative to /root/reference).  The restatement keeps the
// reference's workspace arithmetic (`Alen`, garbage-collection trigger, hash function, dense
// thresholds) because the permutation has to be bit-exact.
//
// Pinned by: colamd_rs/src/lib.rs:253-321 (three permutations), colamd_rs/src/status.rs:167-237
// (required size 57, error codes, jumbled statistics) — see tests/test_oracle_colamd.py.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orc {
namespace colamd {

// colamd_rs/src/colamd.rs:104-126
enum : int {
    STATS = 20,
    DENSE_ROW = 0,
    DENSE_COL = 1,
    DEFRAG_COUNT = 2,
    STATUS = 3,
    INFO1 = 4,
    INFO2 = 5,
    INFO3 = 6,
    OK = 0,
    OK_BUT_JUMBLED = 1,
    ERROR_NROW_NEGATIVE = -3,
    ERROR_NCOL_NEGATIVE = -4,
    ERROR_NNZ_NEGATIVE = -5,
    ERROR_P0_NONZERO = -6,
    ERROR_A_TOO_SMALL = -7,
    ERROR_COL_LENGTH_NEGATIVE = -8,
    ERROR_ROW_INDEX_OUT_OF_BOUNDS = -9,
};
static const int EMPTY = -1;
static const int ALIVE = 0;
static const int DEAD = -1;
static const int DEAD_PRINCIPAL = -1;
static const int DEAD_NON_PRINCIPAL = -2;

// colamd_rs/src/colamd.rs:33-102 — the C unions are kept as unions so that the aliasing of
// score/order, thickness/parent, prev/hash/headhash and degree_next/hash_next is identical.
struct RowT {
    int start, length;
    union { int degree, p; } s1;
    union { int mark, first_column; } s2;
};
struct ColT {
    int start, length;
    union { int thickness, parent; } s1;
    union { int score, order; } s2;
    union { int headhash, hash, prev; } s3;
    union { int degree_next, hash_next; } s4;
};

// colamd_rs/src/options.rs:25-29
struct Options {
    double dense_row_control = 10.0;
    double dense_column_control = 10.0;
    bool aggressive_row_absorption = true;
};

// colamd_rs/src/colamd.rs:139-158.  Returns false on negative input / overflow.
inline bool recommended(int nnz, int n_row, int n_col, size_t* out) {
    if (nnz < 0 || n_row < 0 || n_col < 0) return false;
    size_t c = ((size_t)n_col + 1) * sizeof(ColT) / sizeof(int);
    size_t r = ((size_t)n_row + 1) * sizeof(RowT) / sizeof(int);
    *out = (size_t)nnz * 2 + c + r + (size_t)n_col + (size_t)nnz / 5;
    return true;
}

// colamd_rs/src/colamd.rs:1303-1326
inline int clear_mark(int tag_mark, int max_mark, int n_row, RowT* Row) {
    if (tag_mark <= 0 || tag_mark >= max_mark) {
        for (int r = 0; r < n_row; r++)
            if (Row[r].s2.mark >= ALIVE) Row[r].s2.mark = 0;
        tag_mark = 1;
    }
    return tag_mark;
}

// colamd_rs/src/colamd.rs:502-655
inline bool init_rows_cols(int n_row, int n_col, RowT* Row, ColT* Col, int* A, int* p, int* stats) {
    for (int col = 0; col < n_col; col++) {
        Col[col].start = p[col];
        Col[col].length = p[col + 1] - p[col];
        if (Col[col].length < 0) {
            stats[STATUS] = ERROR_COL_LENGTH_NEGATIVE;
            stats[INFO1] = col;
            stats[INFO2] = Col[col].length;
            return false;
        }
        Col[col].s1.thickness = 1;
        Col[col].s2.score = 0;
        Col[col].s3.prev = EMPTY;
        Col[col].s4.degree_next = EMPTY;
    }
    stats[INFO3] = 0;
    for (int r = 0; r < n_row; r++) {
        Row[r].length = 0;
        Row[r].s2.mark = -1;
    }
    for (int col = 0; col < n_col; col++) {
        int last_row = -1;
        for (int cp = p[col]; cp < p[col + 1]; cp++) {
            int row = A[cp];
            if (row < 0 || row >= n_row) {
                stats[STATUS] = ERROR_ROW_INDEX_OUT_OF_BOUNDS;
                stats[INFO1] = col;
                stats[INFO2] = row;
                stats[INFO3] = n_row;
                return false;
            }
            if (row <= last_row || Row[row].s2.mark == col) {
                stats[STATUS] = OK_BUT_JUMBLED;
                stats[INFO1] = col;
                stats[INFO2] = row;
                stats[INFO3] += 1;
            }
            if (Row[row].s2.mark != col) {
                Row[row].length += 1;
            } else {
                Col[col].length -= 1;
            }
            Row[row].s2.mark = col;
            last_row = row;
        }
    }
    // row pointers — colamd.rs:584-592 (the reference indexes rows[0] unconditionally; its row
    // slice always has n_row+1 entries)
    Row[0].start = p[n_col];
    Row[0].s1.p = Row[0].start;
    Row[0].s2.mark = -1;
    for (int row = 1; row < n_row; row++) {
        Row[row].start = Row[row - 1].start + Row[row - 1].length;
        Row[row].s1.p = Row[row].start;
        Row[row].s2.mark = -1;
    }
    // row form — colamd.rs:596-618
    if (stats[STATUS] == OK_BUT_JUMBLED) {
        for (int col = 0; col < n_col; col++)
            for (int cp = p[col]; cp < p[col + 1]; cp++) {
                int row = A[cp];
                if (Row[row].s2.mark != col) {
                    A[Row[row].s1.p] = col;
                    Row[row].s1.p += 1;
                    Row[row].s2.mark = col;
                }
            }
    } else {
        for (int col = 0; col < n_col; col++)
            for (int cp = p[col]; cp < p[col + 1]; cp++) {
                int row = A[cp];
                A[Row[row].s1.p] = col;
                Row[row].s1.p += 1;
            }
    }
    for (int r = 0; r < n_row; r++) {
        Row[r].s2.mark = 0;
        Row[r].s1.degree = Row[r].length;
    }
    // re-create the column form for jumbled input — colamd.rs:629-651.
    // NOTE (reference quirk kept): the reference iterates `for rp in A[start]..A[start]+length`,
    // i.e. over the VALUE range starting at the row's first column index, not over the stored
    // column indices (the C original walks the stored indices).  The LM path never reaches this
    // branch: its CSC input is sorted and duplicate-free (solvi/src/sparse_col_mat.rs:690-737).
    if (stats[STATUS] == OK_BUT_JUMBLED) {
        Col[0].start = 0;
        p[0] = Col[0].start;
        for (int col = 1; col < n_col; col++) {
            Col[col].start = Col[col - 1].start + Col[col - 1].length;
            p[col] = Col[col].start;
        }
        for (int row = 0; row < n_row; row++) {
            int first = A[Row[row].start];
            for (int rp = first; rp < first + Row[row].length; rp++) {
                A[p[rp]] = row;
                p[rp] += 1;
            }
        }
    }
    return true;
}

// colamd_rs/src/colamd.rs:656-809
inline void init_scoring(int n_row, int n_col, RowT* Row, ColT* Col, int* A, int* head,
                         const Options& opt, int* p_n_row2, int* p_n_col2, int* p_max_deg) {
    int dense_row_count, dense_col_count;
    if (opt.dense_row_control < 0.) {
        dense_row_count = n_col - 1;
    } else {
        double v = opt.dense_row_control * std::sqrt((double)n_col);
        dense_row_count = (int)(16.0 > v ? 16.0 : v);
    }
    if (opt.dense_column_control < 0.) {
        dense_col_count = n_row - 1;
    } else {
        double v = opt.dense_column_control * std::sqrt((double)(n_row < n_col ? n_row : n_col));
        dense_col_count = (int)(16.0 > v ? 16.0 : v);
    }
    int max_deg = 0, n_col2 = n_col, n_row2 = n_row;
    // kill empty columns
    for (int c = n_col - 1; c >= 0; c--) {
        if (Col[c].length == 0) {
            Col[c].s2.order = --n_col2;
            Col[c].start = DEAD_PRINCIPAL;
        }
    }
    // kill dense columns
    for (int c = n_col - 1; c >= 0; c--) {
        if (Col[c].start < ALIVE) continue;
        if (Col[c].length > dense_col_count) {
            Col[c].s2.order = --n_col2;
            int* cp = &A[Col[c].start];
            int* cp_end = cp + Col[c].length;
            while (cp < cp_end) Row[*cp++].s1.degree--;
            Col[c].start = DEAD_PRINCIPAL;
        }
    }
    // kill dense and empty rows
    for (int r = 0; r < n_row; r++) {
        int deg = Row[r].s1.degree;
        if (deg > dense_row_count || deg == 0) {
            Row[r].s2.mark = DEAD;
            --n_row2;
        } else {
            max_deg = max_deg > deg ? max_deg : deg;
        }
    }
    // initial column scores
    for (int c = n_col - 1; c >= 0; c--) {
        if (Col[c].start < ALIVE) continue;
        int score = 0;
        int* cp = &A[Col[c].start];
        int* new_cp = cp;
        int* cp_end = cp + Col[c].length;
        while (cp < cp_end) {
            int row = *cp++;
            if (Row[row].s2.mark < ALIVE) continue;
            *new_cp++ = row;
            score += Row[row].s1.degree - 1;
            score = score < n_col ? score : n_col;
        }
        int col_length = (int)(new_cp - &A[Col[c].start]);
        if (col_length == 0) {
            Col[c].s2.order = --n_col2;
            Col[c].start = DEAD_PRINCIPAL;
        } else {
            Col[c].length = col_length;
            Col[c].s2.score = score;
        }
    }
    // degree lists
    for (int c = 0; c <= n_col; c++) head[c] = EMPTY;
    for (int c = n_col - 1; c >= 0; c--) {
        if (Col[c].start < ALIVE) continue;
        int score = Col[c].s2.score;
        int next_col = head[score];
        Col[c].s3.prev = EMPTY;
        Col[c].s4.degree_next = next_col;
        if (next_col != EMPTY) Col[next_col].s3.prev = c;
        head[score] = c;
    }
    *p_n_col2 = n_col2;
    *p_n_row2 = n_row2;
    *p_max_deg = max_deg;
}

// colamd_rs/src/colamd.rs:1222-1302
inline int garbage_collection(int n_row, int n_col, RowT* Row, ColT* Col, int* A, int* pfree) {
    int* pdest = &A[0];
    for (int c = 0; c < n_col; c++) {
        if (Col[c].start < ALIVE) continue;
        int* psrc = &A[Col[c].start];
        Col[c].start = (int)(pdest - &A[0]);
        int length = Col[c].length;
        for (int j = 0; j < length; j++) {
            int r = *psrc++;
            if (Row[r].s2.mark >= ALIVE) *pdest++ = r;
        }
        Col[c].length = (int)(pdest - &A[Col[c].start]);
    }
    for (int r = 0; r < n_row; r++) {
        if (Row[r].s2.mark < ALIVE || Row[r].length == 0) {
            Row[r].s2.mark = DEAD;
        } else {
            int* psrc = &A[Row[r].start];
            Row[r].s2.first_column = *psrc;
            *psrc = -r - 1;
        }
    }
    int* psrc = pdest;
    while (psrc < pfree) {
        if (*psrc++ < 0) {
            psrc--;
            int r = -(*psrc) - 1;
            *psrc = Row[r].s2.first_column;
            Row[r].start = (int)(pdest - &A[0]);
            int length = Row[r].length;
            for (int j = 0; j < length; j++) {
                int c = *psrc++;
                if (Col[c].start >= ALIVE) *pdest++ = c;
            }
            Row[r].length = (int)(pdest - &A[Row[r].start]);
        }
    }
    return (int)(pdest - &A[0]);
}

// colamd_rs/src/colamd.rs:1139-1221
inline void detect_super_cols(ColT* Col, int* A, int* head, int row_start, int row_length) {
    int* rp = &A[row_start];
    int* rp_end = rp + row_length;
    while (rp < rp_end) {
        int col = *rp++;
        if (Col[col].start < ALIVE) continue;
        int hash = Col[col].s3.hash;
        int head_column = head[hash];
        int first_col;
        if (head_column > EMPTY) first_col = Col[head_column].s3.headhash;
        else first_col = -(head_column + 2);
        for (int super_c = first_col; super_c != EMPTY; super_c = Col[super_c].s4.hash_next) {
            int length = Col[super_c].length;
            int prev_c = super_c;
            for (int c = Col[super_c].s4.hash_next; c != EMPTY; c = Col[c].s4.hash_next) {
                if (Col[c].length != length || Col[c].s2.score != Col[super_c].s2.score) {
                    prev_c = c;
                    continue;
                }
                int* cp1 = &A[Col[super_c].start];
                int* cp2 = &A[Col[c].start];
                int i;
                for (i = 0; i < length; i++)
                    if (*cp1++ != *cp2++) break;
                if (i != length) {
                    prev_c = c;
                    continue;
                }
                Col[super_c].s1.thickness += Col[c].s1.thickness;
                Col[c].s1.parent = super_c;
                Col[c].start = DEAD_NON_PRINCIPAL;
                Col[c].s2.order = EMPTY;
                Col[prev_c].s4.hash_next = Col[c].s4.hash_next;
            }
        }
        if (head_column > EMPTY) Col[head_column].s3.headhash = EMPTY;
        else head[hash] = EMPTY;
    }
}

// colamd_rs/src/colamd.rs:810-1074
inline int find_ordering(int n_row, int n_col, int Alen, RowT* Row, ColT* Col, int* A, int* head,
                         int n_col2, int max_deg, int pfree, bool aggressive) {
    const int max_mark = INT32_MAX - n_col;
    int tag_mark = clear_mark(0, max_mark, n_row, Row);
    int min_score = 0;
    int ngarbage = 0;
    for (int k = 0; k < n_col2;) {
        // select pivot column
        while (head[min_score] == EMPTY && min_score < n_col) min_score++;
        int pivot_col = head[min_score];
        int next_col = Col[pivot_col].s4.degree_next;
        head[min_score] = next_col;
        if (next_col != EMPTY) Col[next_col].s3.prev = EMPTY;
        int pivot_col_score = Col[pivot_col].s2.score;
        Col[pivot_col].s2.order = k;
        int pivot_col_thickness = Col[pivot_col].s1.thickness;
        k += pivot_col_thickness;
        // garbage collection if necessary
        int needed_memory = pivot_col_score < n_col - k ? pivot_col_score : n_col - k;
        if (pfree + needed_memory >= Alen) {
            pfree = garbage_collection(n_row, n_col, Row, Col, A, &A[pfree]);
            ngarbage++;
            tag_mark = clear_mark(0, max_mark, n_row, Row);
        }
        // construct pivot row pattern
        int pivot_row_start = pfree;
        int pivot_row_degree = 0;
        Col[pivot_col].s1.thickness = -pivot_col_thickness;
        {
            int* cp = &A[Col[pivot_col].start];
            int* cp_end = cp + Col[pivot_col].length;
            while (cp < cp_end) {
                int row = *cp++;
                if (Row[row].s2.mark < ALIVE) continue;
                int* rp = &A[Row[row].start];
                int* rp_end = rp + Row[row].length;
                while (rp < rp_end) {
                    int col = *rp++;
                    int col_thickness = Col[col].s1.thickness;
                    if (col_thickness > 0 && Col[col].start >= ALIVE) {
                        Col[col].s1.thickness = -col_thickness;
                        A[pfree++] = col;
                        pivot_row_degree += col_thickness;
                    }
                }
            }
        }
        Col[pivot_col].s1.thickness = pivot_col_thickness;
        max_deg = max_deg > pivot_row_degree ? max_deg : pivot_row_degree;
        // kill all rows used to construct pivot row
        {
            int* cp = &A[Col[pivot_col].start];
            int* cp_end = cp + Col[pivot_col].length;
            while (cp < cp_end) Row[*cp++].s2.mark = DEAD;
        }
        // select a row index to use as the new pivot row
        int pivot_row_length = pfree - pivot_row_start;
        int pivot_row = pivot_row_length > 0 ? A[Col[pivot_col].start] : EMPTY;
        // approximate degree computation: set differences
        {
            int* rp = &A[pivot_row_start];
            int* rp_end = rp + pivot_row_length;
            while (rp < rp_end) {
                int col = *rp++;
                int col_thickness = -Col[col].s1.thickness;
                Col[col].s1.thickness = col_thickness;
                int cur_score = Col[col].s2.score;
                int prev_col = Col[col].s3.prev;
                next_col = Col[col].s4.degree_next;
                if (prev_col == EMPTY) head[cur_score] = next_col;
                else Col[prev_col].s4.degree_next = next_col;
                if (next_col != EMPTY) Col[next_col].s3.prev = prev_col;
                int* cp = &A[Col[col].start];
                int* cp_end = cp + Col[col].length;
                while (cp < cp_end) {
                    int row = *cp++;
                    int row_mark = Row[row].s2.mark;
                    if (row_mark < ALIVE) continue;
                    int set_difference = row_mark - tag_mark;
                    if (set_difference < 0) set_difference = Row[row].s1.degree;
                    set_difference -= col_thickness;
                    if (set_difference == 0 && aggressive) Row[row].s2.mark = DEAD;
                    else Row[row].s2.mark = set_difference + tag_mark;
                }
            }
        }
        // add up set differences for each column
        {
            int* rp = &A[pivot_row_start];
            int* rp_end = rp + pivot_row_length;
            while (rp < rp_end) {
                int col = *rp++;
                uint32_t hash = 0;
                int cur_score = 0;
                int* cp = &A[Col[col].start];
                int* new_cp = cp;
                int* cp_end = cp + Col[col].length;
                while (cp < cp_end) {
                    int row = *cp++;
                    int row_mark = Row[row].s2.mark;
                    if (row_mark < ALIVE) continue;
                    *new_cp++ = row;
                    hash += (uint32_t)row;
                    cur_score += row_mark - tag_mark;
                    cur_score = cur_score < n_col ? cur_score : n_col;
                }
                Col[col].length = (int)(new_cp - &A[Col[col].start]);
                if (Col[col].length == 0) {
                    Col[col].start = DEAD_PRINCIPAL;
                    pivot_row_degree -= Col[col].s1.thickness;
                    Col[col].s2.order = k;
                    k += Col[col].s1.thickness;
                } else {
                    Col[col].s2.score = cur_score;
                    hash %= (uint32_t)(n_col + 1);
                    int head_column = head[hash];
                    int first_col;
                    if (head_column > EMPTY) {
                        first_col = Col[head_column].s3.headhash;
                        Col[head_column].s3.headhash = col;
                    } else {
                        first_col = -(head_column + 2);
                        head[hash] = -(col + 2);
                    }
                    Col[col].s4.hash_next = first_col;
                    Col[col].s3.hash = (int)hash;
                }
            }
        }
        detect_super_cols(Col, A, head, pivot_row_start, pivot_row_length);
        Col[pivot_col].start = DEAD_PRINCIPAL;
        tag_mark = clear_mark(tag_mark + max_deg + 1, max_mark, n_row, Row);
        // finalize the new pivot row and column scores
        int* rp = &A[pivot_row_start];
        int* new_rp = rp;
        int* rp_end = rp + pivot_row_length;
        while (rp < rp_end) {
            int col = *rp++;
            if (Col[col].start < ALIVE) continue;
            *new_rp++ = col;
            A[Col[col].start + Col[col].length] = pivot_row;
            Col[col].length++;
            int cur_score = Col[col].s2.score + pivot_row_degree;
            int max_score = n_col - k - Col[col].s1.thickness;
            cur_score -= Col[col].s1.thickness;
            cur_score = cur_score < max_score ? cur_score : max_score;
            Col[col].s2.score = cur_score;
            next_col = head[cur_score];
            Col[col].s4.degree_next = next_col;
            Col[col].s3.prev = EMPTY;
            if (next_col != EMPTY) Col[next_col].s3.prev = col;
            head[cur_score] = col;
            min_score = min_score < cur_score ? min_score : cur_score;
        }
        if (pivot_row_degree > 0) {
            Row[pivot_row].start = pivot_row_start;
            Row[pivot_row].length = (int)(new_rp - &A[pivot_row_start]);
            Row[pivot_row].s1.degree = pivot_row_degree;
            Row[pivot_row].s2.mark = 0;
        }
    }
    return ngarbage;
}

// colamd_rs/src/colamd.rs:1087-1138
inline void order_children(int n_col, ColT* Col, int* p) {
    for (int i = 0; i < n_col; i++) {
        if (Col[i].start != DEAD_PRINCIPAL && Col[i].s2.order == EMPTY) {
            int parent = i;
            do {
                parent = Col[parent].s1.parent;
            } while (Col[parent].start != DEAD_PRINCIPAL);
            int c = i;
            int order = Col[parent].s2.order;
            do {
                Col[c].s2.order = order++;
                // colamd.rs:1116-1119: the reference (like the C original) collapses first and
                // then reads `col[c].shared1.parent`, i.e. it jumps straight to the principal.
                Col[c].s1.parent = parent;
                c = Col[c].s1.parent;
            } while (Col[c].s2.order == EMPTY);
            Col[parent].s2.order = order;
        }
    }
    for (int c = 0; c < n_col; c++) p[Col[c].s2.order] = c;
}

// colamd_rs/src/colamd.rs:354-494.  `a` has length `a_len` (>= recommended), `p` n_col+1.
// On success p[0..n_col) is the permutation.  Returns true on success.
inline bool colamd(int n_row, int n_col, size_t a_len, int* a, int* p, const Options& opt,
                   int* stats) {
    for (int i = 0; i < STATS; i++) stats[i] = 0;
    stats[STATUS] = OK;
    stats[INFO1] = -1;
    stats[INFO2] = -1;
    if (n_row < 0) { stats[STATUS] = ERROR_NROW_NEGATIVE; stats[INFO1] = n_row; return false; }
    if (n_col < 0) { stats[STATUS] = ERROR_NCOL_NEGATIVE; stats[INFO1] = n_col; return false; }
    int nnz = p[n_col];
    if (nnz < 0) { stats[STATUS] = ERROR_NNZ_NEGATIVE; stats[INFO1] = nnz; return false; }
    if (p[0] != 0) { stats[STATUS] = ERROR_P0_NONZERO; stats[INFO1] = p[0]; return false; }
    size_t col_size = ((size_t)n_col + 1) * sizeof(ColT) / sizeof(int);
    size_t row_size = ((size_t)n_row + 1) * sizeof(RowT) / sizeof(int);
    size_t need = (size_t)nnz * 2 + (size_t)n_col + col_size + row_size;
    if (need > a_len) {
        stats[STATUS] = ERROR_A_TOO_SMALL;
        stats[INFO1] = (int)need;
        stats[INFO2] = (int)a_len;
        return false;
    }
    size_t Alen = a_len - (col_size + row_size);
    // The reference carves Col/Row out of the tail of `a` (bytemuck cast, colamd.rs:445-449);
    // separate vectors give the same behaviour because Alen is what the algorithm observes.
    std::vector<ColT> cols((size_t)n_col + 1);
    std::vector<RowT> rows((size_t)n_row + 1);
    if (!init_rows_cols(n_row, n_col, rows.data(), cols.data(), a, p, stats)) return false;
    int n_row2, n_col2, max_deg;
    init_scoring(n_row, n_col, rows.data(), cols.data(), a, p, opt, &n_row2, &n_col2, &max_deg);
    int ngarbage = find_ordering(n_row, n_col, (int)Alen, rows.data(), cols.data(), a, p, n_col2,
                                 max_deg, 2 * nnz, opt.aggressive_row_absorption);
    order_children(n_col, cols.data(), p);
    stats[DENSE_ROW] = n_row - n_row2;
    stats[DENSE_COL] = n_col - n_col2;
    stats[DEFRAG_COUNT] = ngarbage;
    return true;
}

// colamd_rs/src/colamd.rs:162-352.  perm has n+1 entries.
inline bool symamd(int n, const int* a, const int* p, int* perm, Options opt, int* stats) {
    for (int i = 0; i < STATS; i++) stats[i] = 0;
    stats[STATUS] = OK;
    stats[INFO1] = -1;
    stats[INFO2] = -1;
    if (n < 0) { stats[STATUS] = ERROR_NCOL_NEGATIVE; stats[INFO1] = n; return false; }
    int nnz = p[n];
    if (nnz < 0) { stats[STATUS] = ERROR_NNZ_NEGATIVE; stats[INFO1] = nnz; return false; }
    if (p[0] != 0) { stats[STATUS] = ERROR_P0_NONZERO; stats[INFO1] = p[0]; return false; }
    std::vector<int> count((size_t)n + 1, 0), mark((size_t)n + 1, 0);
    stats[INFO3] = 0;
    for (int i = 0; i < n; i++) mark[i] = -1;
    for (int j = 0; j < n; j++) {
        int last_row = -1;
        int length = p[j + 1] - p[j];
        if (length < 0) {
            stats[STATUS] = ERROR_COL_LENGTH_NEGATIVE;
            stats[INFO1] = j;
            stats[INFO2] = length;
            return false;
        }
        for (int pp = p[j]; pp < p[j + 1]; pp++) {
            int i = a[pp];
            if (i < 0 || i >= n) {
                stats[STATUS] = ERROR_ROW_INDEX_OUT_OF_BOUNDS;
                stats[INFO1] = j;
                stats[INFO2] = i;
                stats[INFO3] = n;
                return false;
            }
            if (i <= last_row || mark[i] == j) {
                stats[STATUS] = OK_BUT_JUMBLED;
                stats[INFO1] = j;
                stats[INFO2] = i;
                stats[INFO3] += 1;
            }
            if (i > j && mark[i] != j) {
                count[i]++;
                count[j]++;
            }
            mark[i] = j;
            last_row = i;
        }
    }
    perm[0] = 0;
    for (int j = 1; j <= n; j++) perm[j] = perm[j - 1] + count[j - 1];
    for (int j = 0; j < n; j++) count[j] = perm[j];
    int mnz = perm[n];
    int n_row = mnz / 2;
    size_t m_len;
    if (!recommended(mnz, n_row, n, &m_len)) return false;
    std::vector<int> m(m_len, 0);
    int k = 0;
    if (stats[STATUS] == OK) {
        for (int j = 0; j < n; j++)
            for (int pp = p[j]; pp < p[j + 1]; pp++) {
                int i = a[pp];
                if (i > j) {
                    m[count[i]++] = k;
                    m[count[j]++] = k;
                    k++;
                }
            }
    } else {
        for (int i = 0; i < n; i++) mark[i] = -1;
        for (int j = 0; j < n; j++)
            for (int pp = p[j]; pp < p[j + 1]; pp++) {
                int i = a[pp];
                if (i > j && mark[i] != j) {
                    m[count[i]++] = k;
                    m[count[j]++] = k;
                    k++;
                    mark[i] = j;
                }
            }
    }
    opt.dense_column_control = opt.dense_row_control;
    opt.dense_row_control = -1.;
    colamd(n_row, n, m_len, m.data(), perm, opt, stats);
    stats[DENSE_ROW] = stats[DENSE_COL];
    return true;
}

}  // namespace colamd
}  // namespace orc
