// Column approximate-minimum-degree ordering (COLAMD, Davis/Gilbert/Larimore/Ng 2004) for the
// product's symbolic pipeline.  The permutation has to be bit-identical to what the reference
// obtains from colamd_rs::colamd (colamd_rs/src/colamd.rs:354-1326, default knobs
// colamd_rs/src/options.rs:25-29) on the augmented Jacobian pattern, because it fixes the
// elimination order of everything downstream.  Input here is always a clean CSC pattern (sorted,
// duplicate-free — the pipeline builds it itself), so the reference's input validation and
// "jumbled" repair paths have no counterpart.
//
// Layout: struct-of-arrays for the row/column records instead of the reference's packed unions
// carved out of the tail of the index workspace; the workspace length that drives garbage
// collection is computed exactly as colamd_recommended does (colamd.rs:139-158,445).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace fk {

class ColamdOrder {
public:
    struct Stats {
        int dense_rows = 0, dense_cols = 0, garbage_collections = 0;
    };

    // colptr[n_col+1], rowidx[nnz]; returns perm[n_col]: perm[k] = original column placed k-th.
    static std::vector<int32_t> order(int n_row, int n_col, const uint32_t* colptr,
                                      const uint32_t* rowidx, Stats* stats = nullptr) {
        ColamdOrder c(n_row, n_col, colptr, rowidx);
        c.score_initial();
        c.eliminate();
        c.place_absorbed();
        if (stats) *stats = c.stats_;
        std::vector<int32_t> perm(n_col);
        for (int j = 0; j < n_col; j++) perm[c.col_rank_[j]] = j;
        return perm;
    }

private:
    static constexpr int kNone = -1;

    int n_row_, n_col_, nnz_;
    int ws_len_;            // usable index workspace ("Alen" after the record area is removed)
    int free_top_;          // first unused workspace slot
    std::vector<int> ws_;   // column lists, then row lists, then scratch for pivot rows

    // column records
    std::vector<int> col_at_, col_len_;
    std::vector<int> col_thick_;   // thickness while alive; parent column once absorbed
    std::vector<int> col_rank_;    // score while alive; output position once ordered
    std::vector<int> col_prev_;    // degree-list back link / hash value / hash-bucket head
    std::vector<int> col_next_;    // degree-list forward link / hash-chain link
    // row records
    std::vector<int> row_at_, row_len_, row_deg_, row_tag_;
    // degree-list heads, indexed by score 0..n_col; doubles as the hash table during detection
    std::vector<int> head_;

    int live_cols_ = 0, live_rows_ = 0, max_deg_ = 0;
    Stats stats_;

    bool col_alive(int c) const { return col_at_[c] >= 0; }
    bool row_alive(int r) const { return row_tag_[r] >= 0; }
    void kill_principal(int c) { col_at_[c] = -1; }
    void kill_absorbed(int c) { col_at_[c] = -2; }

    ColamdOrder(int n_row, int n_col, const uint32_t* colptr, const uint32_t* rowidx)
        : n_row_(n_row), n_col_(n_col), nnz_((int)colptr[n_col]) {
        // same arithmetic as colamd_recommended: 2*nnz + n_col + nnz/5 usable ints
        ws_len_ = 2 * nnz_ + n_col_ + nnz_ / 5;
        ws_.assign((size_t)ws_len_ + 1, 0);
        col_at_.resize(n_col + 1); col_len_.resize(n_col + 1); col_thick_.assign(n_col + 1, 1);
        col_rank_.assign(n_col + 1, 0); col_prev_.assign(n_col + 1, kNone); col_next_.assign(n_col + 1, kNone);
        row_at_.assign(n_row + 1, 0); row_len_.assign(n_row + 1, 0); row_deg_.assign(n_row + 1, 0);
        row_tag_.assign(n_row + 1, 0);
        head_.assign(n_col + 1, kNone);
        for (int c = 0; c < n_col; c++) {
            col_at_[c] = (int)colptr[c];
            col_len_[c] = (int)(colptr[c + 1] - colptr[c]);
        }
        for (int k = 0; k < nnz_; k++) {
            ws_[k] = (int)rowidx[k];
            row_len_[rowidx[k]]++;
        }
        // row form directly behind the column form
        int at = nnz_;
        for (int r = 0; r < n_row; r++) {
            row_at_[r] = at;
            at += row_len_[r];
        }
        std::vector<int> fill(row_at_.begin(), row_at_.begin() + n_row);
        for (int c = 0; c < n_col; c++)
            for (int k = col_at_[c]; k < col_at_[c] + col_len_[c]; k++) ws_[fill[ws_[k]]++] = c;
        for (int r = 0; r < n_row; r++) row_deg_[r] = row_len_[r];
        free_top_ = 2 * nnz_;
    }

    // Dense/empty removal, initial scores, degree lists (colamd.rs:656-809).
    void score_initial() {
        const double kDense = 10.0;
        double dr = kDense * std::sqrt((double)n_col_);
        double dc = kDense * std::sqrt((double)(n_row_ < n_col_ ? n_row_ : n_col_));
        const int dense_row = (int)(16.0 > dr ? 16.0 : dr);
        const int dense_col = (int)(16.0 > dc ? 16.0 : dc);
        live_cols_ = n_col_;
        live_rows_ = n_row_;
        for (int c = n_col_ - 1; c >= 0; c--)
            if (col_len_[c] == 0) {
                col_rank_[c] = --live_cols_;
                kill_principal(c);
            }
        for (int c = n_col_ - 1; c >= 0; c--) {
            if (!col_alive(c) || col_len_[c] <= dense_col) continue;
            col_rank_[c] = --live_cols_;
            for (int k = col_at_[c]; k < col_at_[c] + col_len_[c]; k++) row_deg_[ws_[k]]--;
            kill_principal(c);
        }
        for (int r = 0; r < n_row_; r++) {
            int d = row_deg_[r];
            if (d > dense_row || d == 0) {
                row_tag_[r] = -1;
                live_rows_--;
            } else if (d > max_deg_) {
                max_deg_ = d;
            }
        }
        for (int c = n_col_ - 1; c >= 0; c--) {
            if (!col_alive(c)) continue;
            int score = 0, w = col_at_[c];
            for (int k = col_at_[c]; k < col_at_[c] + col_len_[c]; k++) {
                int r = ws_[k];
                if (!row_alive(r)) continue;
                ws_[w++] = r;
                score += row_deg_[r] - 1;
                if (score > n_col_) score = n_col_;
            }
            int len = w - col_at_[c];
            if (len == 0) {
                col_rank_[c] = --live_cols_;
                kill_principal(c);
            } else {
                col_len_[c] = len;
                col_rank_[c] = score;
            }
        }
        for (int c = n_col_ - 1; c >= 0; c--)
            if (col_alive(c)) push_degree_list(c, col_rank_[c]);
        stats_.dense_rows = n_row_ - live_rows_;
        stats_.dense_cols = n_col_ - live_cols_;
    }

    void push_degree_list(int c, int score) {
        int nx = head_[score];
        col_prev_[c] = kNone;
        col_next_[c] = nx;
        if (nx != kNone) col_prev_[nx] = c;
        head_[score] = c;
    }
    void unlink_degree_list(int c) {
        int pv = col_prev_[c], nx = col_next_[c];
        if (pv == kNone) head_[col_rank_[c]] = nx;
        else col_next_[pv] = nx;
        if (nx != kNone) col_prev_[nx] = pv;
    }

    int reset_tags_if_needed(int tag, int limit) {
        if (tag <= 0 || tag >= limit) {
            for (int r = 0; r < n_row_; r++)
                if (row_alive(r)) row_tag_[r] = 0;
            tag = 1;
        }
        return tag;
    }

    // Compact the workspace (colamd.rs:1222-1302).  Order inside every list is preserved.
    int compact() {
        int w = 0;
        for (int c = 0; c < n_col_; c++) {
            if (!col_alive(c)) continue;
            int src = col_at_[c], len = col_len_[c];
            col_at_[c] = w;
            for (int k = 0; k < len; k++) {
                int r = ws_[src + k];
                if (row_alive(r)) ws_[w++] = r;
            }
            col_len_[c] = w - col_at_[c];
        }
        // tag the first slot of every surviving row list so the lists can be found in storage order
        std::vector<int> first(n_row_, 0);
        for (int r = 0; r < n_row_; r++) {
            if (!row_alive(r) || row_len_[r] == 0) {
                row_tag_[r] = -1;
            } else {
                first[r] = ws_[row_at_[r]];
                ws_[row_at_[r]] = -r - 1;
            }
        }
        int src = w;
        while (src < free_top_) {
            if (ws_[src] >= 0) {
                src++;
                continue;
            }
            int r = -ws_[src] - 1;
            ws_[src] = first[r];
            row_at_[r] = w;
            int len = row_len_[r];
            for (int k = 0; k < len; k++) {
                int c = ws_[src++];
                if (col_alive(c)) ws_[w++] = c;
            }
            row_len_[r] = w - row_at_[r];
        }
        return w;
    }

    // Main elimination loop (colamd.rs:810-1074).
    void eliminate() {
        const int tag_limit = INT32_MAX - n_col_;
        int tag = reset_tags_if_needed(0, tag_limit);
        int min_score = 0;
        int k = 0;
        while (k < live_cols_) {
            while (head_[min_score] == kNone && min_score < n_col_) min_score++;
            const int pivot = head_[min_score];
            {
                int nx = col_next_[pivot];
                head_[min_score] = nx;
                if (nx != kNone) col_prev_[nx] = kNone;
            }
            const int pivot_score = col_rank_[pivot];
            const int pivot_thick = col_thick_[pivot];
            col_rank_[pivot] = k;
            k += pivot_thick;

            int need = pivot_score < n_col_ - k ? pivot_score : n_col_ - k;
            if (free_top_ + need >= ws_len_) {
                free_top_ = compact();
                stats_.garbage_collections++;
                tag = reset_tags_if_needed(0, tag_limit);
            }

            // pivot row = union of the live rows of the pivot column
            const int prow_at = free_top_;
            int prow_deg = 0;
            col_thick_[pivot] = -pivot_thick;
            for (int a = col_at_[pivot]; a < col_at_[pivot] + col_len_[pivot]; a++) {
                int r = ws_[a];
                if (!row_alive(r)) continue;
                for (int b = row_at_[r]; b < row_at_[r] + row_len_[r]; b++) {
                    int c = ws_[b];
                    int t = col_thick_[c];
                    if (t > 0 && col_alive(c)) {
                        col_thick_[c] = -t;
                        ws_[free_top_++] = c;
                        prow_deg += t;
                    }
                }
            }
            col_thick_[pivot] = pivot_thick;
            if (prow_deg > max_deg_) max_deg_ = prow_deg;
            for (int a = col_at_[pivot]; a < col_at_[pivot] + col_len_[pivot]; a++) row_tag_[ws_[a]] = -1;
            const int prow_len = free_top_ - prow_at;
            const int prow = prow_len > 0 ? ws_[col_at_[pivot]] : kNone;

            // set differences |row \ pivot row| accumulated in the tags
            for (int b = prow_at; b < prow_at + prow_len; b++) {
                int c = ws_[b];
                int t = -col_thick_[c];
                col_thick_[c] = t;
                unlink_degree_list(c);
                for (int a = col_at_[c]; a < col_at_[c] + col_len_[c]; a++) {
                    int r = ws_[a];
                    int rt = row_tag_[r];
                    if (rt < 0) continue;
                    int diff = rt - tag;
                    if (diff < 0) diff = row_deg_[r];
                    diff -= t;
                    row_tag_[r] = diff == 0 ? -1 /* aggressive absorption */ : diff + tag;
                }
            }
            // new scores + hash for supercolumn detection
            for (int b = prow_at; b < prow_at + prow_len; b++) {
                int c = ws_[b];
                uint32_t hash = 0;
                int score = 0, w = col_at_[c];
                for (int a = col_at_[c]; a < col_at_[c] + col_len_[c]; a++) {
                    int r = ws_[a];
                    int rt = row_tag_[r];
                    if (rt < 0) continue;
                    ws_[w++] = r;
                    hash += (uint32_t)r;
                    score += rt - tag;
                    if (score > n_col_) score = n_col_;
                }
                col_len_[c] = w - col_at_[c];
                if (col_len_[c] == 0) {
                    kill_principal(c);
                    prow_deg -= col_thick_[c];
                    col_rank_[c] = k;
                    k += col_thick_[c];
                } else {
                    col_rank_[c] = score;
                    int h = (int)(hash % (uint32_t)(n_col_ + 1));
                    int hd = head_[h], first;
                    if (hd > kNone) {  // bucket shares its slot with a degree list: park the chain in its head
                        first = col_prev_[hd];
                        col_prev_[hd] = c;
                    } else {
                        first = -(hd + 2);
                        head_[h] = -(c + 2);
                    }
                    col_next_[c] = first;
                    col_prev_[c] = h;
                }
            }
            merge_identical_columns(prow_at, prow_len);
            kill_principal(pivot);
            tag = reset_tags_if_needed(tag + max_deg_ + 1, tag_limit);

            // final scores, append the pivot row to the surviving columns, back into the lists
            int w = prow_at;
            for (int b = prow_at; b < prow_at + prow_len; b++) {
                int c = ws_[b];
                if (!col_alive(c)) continue;
                ws_[w++] = c;
                ws_[col_at_[c] + col_len_[c]++] = prow;
                int score = col_rank_[c] + prow_deg;
                int cap = n_col_ - k - col_thick_[c];
                score -= col_thick_[c];
                if (score > cap) score = cap;
                col_rank_[c] = score;
                push_degree_list(c, score);
                if (score < min_score) min_score = score;
            }
            if (prow_deg > 0) {
                row_at_[prow] = prow_at;
                row_len_[prow] = w - prow_at;
                row_deg_[prow] = prow_deg;
                row_tag_[prow] = 0;
            }
        }
    }

    // Supercolumn detection over the hash buckets filled above (colamd.rs:1139-1221).
    void merge_identical_columns(int prow_at, int prow_len) {
        for (int b = prow_at; b < prow_at + prow_len; b++) {
            int c0 = ws_[b];
            if (!col_alive(c0)) continue;
            int h = col_prev_[c0];
            int hd = head_[h];
            int first = hd > kNone ? col_prev_[hd] : -(hd + 2);
            for (int rep = first; rep != kNone; rep = col_next_[rep]) {
                int len = col_len_[rep], prev = rep;
                for (int c = col_next_[rep]; c != kNone; c = col_next_[c]) {
                    bool same = col_len_[c] == len && col_rank_[c] == col_rank_[rep];
                    if (same) {
                        const int* x = &ws_[col_at_[rep]];
                        const int* y = &ws_[col_at_[c]];
                        for (int i = 0; i < len && same; i++) same = x[i] == y[i];
                    }
                    if (!same) {
                        prev = c;
                        continue;
                    }
                    col_thick_[rep] += col_thick_[c];
                    col_thick_[c] = rep;  // parent
                    kill_absorbed(c);
                    col_rank_[c] = kNone;
                    col_next_[prev] = col_next_[c];
                }
            }
            if (hd > kNone) col_prev_[hd] = kNone;
            else head_[h] = kNone;
        }
    }

    // Give absorbed columns their positions right before their representative (colamd.rs:1087-1138).
    void place_absorbed() {
        for (int i = 0; i < n_col_; i++) {
            if (col_at_[i] == -1 || col_rank_[i] != kNone) continue;
            int rep = i;
            do rep = col_thick_[rep]; while (col_at_[rep] != -1);
            int c = i, pos = col_rank_[rep];
            do {
                col_rank_[c] = pos++;
                col_thick_[c] = rep;
                c = col_thick_[c];
            } while (col_rank_[c] == kNone);
            col_rank_[rep] = pos;
        }
    }
};

}  // namespace fk
