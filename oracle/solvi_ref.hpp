// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the parts of `solvi` (the reference's sparse linear algebra crate) that the
// Levenberg–Marquardt path uses: COO→CSC assembly, column permutation, column elimination tree,
// post-order, Gilbert–Ng–Peyton row/column counts, R- and Householder-pattern construction,
// left-looking sparse Householder QR, Qᵀb, back-substitution.  Every routine cites the reference
// lines it follows (paths relative to /root/reference).
//
// Pinned by the reference's own known-answer tests: solvi/src/decomposition/sparse/
// cholesky.rs:602-796, qr.rs:376-652, sparse_col_mat.rs:835-869, utils.rs doctests,
// permutation.rs:96-125 — see tests/test_oracle_solvi.py.
//
// Deliberate differences (results identical):
//  * `CholeskyCounts::build`'s recursive DSU `find` (cholesky.rs:257-262) is iterative here (same
//    full path compression) so that long elimination-tree chains cannot overflow the stack.
//  * `from_triplet_mat` uses a stable sort; the reference's `sort_unstable_by_key` leaves the
//    summation order of 3+ duplicates unspecified, two duplicates commute.
//  * The inverse column permutation is applied as a gather instead of a swap sequence
//    (permutation.rs:41-80); a permutation moves values without arithmetic.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "colamd_ref.hpp"

namespace orc {
namespace solvi {

static const size_t NONE = SIZE_MAX;  // usize::MAX sentinel of the reference
using idx_vec = std::vector<size_t>;

// solvi/src/triplet_mat.rs:31-37,99-117
struct TripletMat {
    size_t nrows = 0, ncols = 0;
    idx_vec row_indices, col_indices;
    std::vector<double> values;
    TripletMat() {}
    TripletMat(size_t r, size_t c) : nrows(r), ncols(c) {}
    void push_triplet(size_t row, size_t col, double v) {
        nrows = std::max(nrows, row + 1);
        ncols = std::max(ncols, col + 1);
        row_indices.push_back(row);
        col_indices.push_back(col);
        values.push_back(v);
    }
    void clear() {
        nrows = 0;
        ncols = 0;
        row_indices.clear();
        col_indices.clear();
        values.clear();
    }
};

// solvi/src/sparse_col_mat.rs (SparseColMatStructure)
struct Structure {
    size_t nrows = 0, ncols = 0;
    idx_vec row_indices;
    idx_vec column_pointers;
    const size_t* col_begin(size_t j) const { return row_indices.data() + column_pointers[j]; }
    const size_t* col_end(size_t j) const { return row_indices.data() + column_pointers[j + 1]; }
    size_t col_len(size_t j) const { return column_pointers[j + 1] - column_pointers[j]; }

    // sparse_col_mat.rs:456-502
    Structure permute_columns(const idx_vec& perm) const {
        Structure out;
        out.nrows = nrows;
        out.ncols = perm.size();
        out.row_indices.reserve(row_indices.size());
        out.column_pointers.reserve(perm.size() + 1);
        out.column_pointers.push_back(0);
        for (size_t idx = 0; idx < perm.size(); idx++) {
            size_t j = perm[idx];
            out.column_pointers.push_back(out.column_pointers[idx] + col_len(j));
            out.row_indices.insert(out.row_indices.end(), col_begin(j), col_end(j));
        }
        return out;
    }
};

struct SparseColMat {
    Structure structure;
    std::vector<double> values;

    // sparse_col_mat.rs:690-737
    static SparseColMat from_triplet_mat(const TripletMat& a) {
        size_t nnz = a.values.size();
        SparseColMat out;
        out.structure.nrows = a.nrows;
        out.structure.ncols = a.ncols;
        out.structure.column_pointers.assign(a.ncols + 1, 0);
        out.structure.row_indices.reserve(nnz);
        out.values.reserve(nnz);
        idx_vec indices(nnz);
        std::iota(indices.begin(), indices.end(), (size_t)0);
        std::stable_sort(indices.begin(), indices.end(), [&](size_t x, size_t y) {
            if (a.col_indices[x] != a.col_indices[y]) return a.col_indices[x] < a.col_indices[y];
            return a.row_indices[x] < a.row_indices[y];
        });
        size_t prev_row = NONE, prev_col = NONE;
        for (size_t idx : indices) {
            size_t row = a.row_indices[idx], col = a.col_indices[idx];
            if (row == prev_row && col == prev_col) {
                out.values.back() += a.values[idx];
            } else {
                if (col != prev_col)
                    for (size_t c = prev_col + 1; c <= col; c++)  // NONE + 1 wraps to 0
                        out.structure.column_pointers[c] = out.values.size();
                out.values.push_back(a.values[idx]);
                out.structure.row_indices.push_back(row);
            }
            prev_row = row;
            prev_col = col;
        }
        for (size_t c = prev_col + 1; c <= a.ncols; c++)
            out.structure.column_pointers[c] = out.values.size();
        return out;
    }

    // sparse_col_mat.rs:788-826.  b has length nrows.
    bool solve_upper_triangular_mut(double* b) const {
        for (size_t ii = structure.nrows; ii-- > 0;) {
            size_t i = ii;
            size_t lo = structure.column_pointers[i], hi = structure.column_pointers[i + 1];
            double diag = 0.0;
            if (hi > lo && structure.row_indices[hi - 1] == i) diag = values[hi - 1];
            if (diag == 0.0) return false;
            double coeff = b[i] / diag;
            b[i] = coeff;
            for (size_t k = lo; k < hi; k++) {
                size_t row = structure.row_indices[k];
                if (!(row < i)) break;  // take_while(row < i)
                b[row] = b[row] - coeff * values[k];
            }
        }
        return true;
    }
};

// solvi/src/permutation.rs:41-80, as a gather: out[i] = a[permutation[i]].
inline void gather_permute(const idx_vec& permutation, double* a) {
    std::vector<double> tmp(permutation.size());
    for (size_t i = 0; i < permutation.size(); i++) tmp[i] = a[permutation[i]];
    for (size_t i = 0; i < permutation.size(); i++) a[i] = tmp[i];
}
// The reference's swap-sequence construction, kept for its own unit test (permutation.rs:41-66).
inline std::vector<std::pair<size_t, size_t>> permutation_swaps(const idx_vec& permutation) {
    std::vector<std::pair<size_t, size_t>> swaps;
    std::vector<char> seen(permutation.size(), 0);
    idx_vec stack;
    for (size_t start : permutation) {
        size_t i = start;
        while (!seen[i]) {
            stack.push_back(i);
            seen[i] = 1;
            i = permutation[i];
        }
        if (!stack.empty()) {
            size_t pivot = stack[0];
            for (size_t k = stack.size(); k-- > 1;) swaps.push_back({pivot, stack[k]});
            stack.clear();
        }
    }
    return swaps;
}

// solvi/src/utils.rs:49-117
inline idx_vec post_order(const idx_vec& parents) {
    size_t n = parents.size();
    idx_vec head(n, NONE), next(n, NONE), post, stack;
    post.reserve(n);
    for (size_t node = 0; node < n; node++) {
        size_t parent = parents[node];
        if (parent != NONE) {
            next[node] = head[parent];
            head[parent] = node;
        }
    }
    for (size_t root = 0; root < n; root++) {
        if (parents[root] != NONE) continue;
        size_t node = root;
        for (;;) {
            while (head[node] != NONE) {
                size_t child = head[node];
                head[node] = next[child];
                stack.push_back(node);
                node = child;
            }
            post.push_back(node);
            if (stack.empty()) break;
            node = stack.back();
            stack.pop_back();
        }
    }
    return post;
}

// solvi/src/utils.rs:153-188
inline idx_vec node_depth_levels(const idx_vec& parents, size_t* max_level_out = nullptr) {
    size_t n = parents.size(), max_level = 0;
    idx_vec levels(n, 0), path;
    for (size_t start = 0; start < n; start++) {
        size_t node = start;
        while (levels[node] == 0 && parents[node] != NONE) {
            path.push_back(node);
            node = parents[node];
        }
        size_t level = levels[node];
        while (!path.empty()) {
            size_t v = path.back();
            path.pop_back();
            level += 1;
            levels[v] = level;
            max_level = std::max(max_level, level);
        }
    }
    if (max_level_out) *max_level_out = max_level;
    return levels;
}

// solvi/src/decomposition/sparse/cholesky.rs:31-84 with SYMMETRIC == false (tree of AᵀA).
inline idx_vec elimination_tree_ata(const Structure& a) {
    size_t m = a.nrows, n = a.ncols;
    idx_vec parents(n, NONE), ancestors(n, NONE), prev_col(m, NONE);
    for (size_t col = 0; col < n; col++) {
        for (const size_t* rp = a.col_begin(col); rp != a.col_end(col); ++rp) {
            size_t row = *rp;
            size_t k = prev_col[row];
            while (k != NONE) {
                if (k >= col) break;
                size_t col_next = ancestors[k];
                ancestors[k] = col;
                if (col_next == NONE) parents[k] = col;
                k = col_next;
            }
            prev_col[row] = col;
        }
    }
    return parents;
}
// Same routine with SYMMETRIC == true (tree of a symmetric A, upper part), for completeness.
inline idx_vec elimination_tree_sym(const Structure& a) {
    size_t n = a.ncols;
    idx_vec parents(n, NONE), ancestors(n, NONE);
    for (size_t col = 0; col < n; col++) {
        for (const size_t* rp = a.col_begin(col); rp != a.col_end(col); ++rp) {
            size_t k = *rp;
            while (k != NONE) {
                if (k >= col) break;
                size_t col_next = ancestors[k];
                ancestors[k] = col;
                if (col_next == NONE) parents[k] = col;
                k = col_next;
            }
        }
    }
    return parents;
}

// cholesky.rs:98-105,149-332
struct CholeskyCounts {
    idx_vec row_counts, col_counts, levels, first_columns;

    static CholeskyCounts build(const Structure& a, const idx_vec& parents, const idx_vec& postorder) {
        size_t m = a.nrows, n = a.ncols;
        CholeskyCounts out;
        out.levels = node_depth_levels(parents);
        const idx_vec& levels = out.levels;
        idx_vec places_in_postorder(n, 0);
        for (size_t place = 0; place < n; place++) places_in_postorder[postorder[place]] = place;
        idx_vec subtree_size(n, 1);
        for (size_t j : postorder) {
            size_t parent = parents[j];
            if (parent != NONE) subtree_size[parent] += subtree_size[j];
        }
        idx_vec first_descendants(n, 0);
        for (size_t place = 0; place < n; place++) {
            size_t j = postorder[place];
            first_descendants[j] = postorder[place + 1 - subtree_size[j]];
        }
        idx_vec& first_columns = out.first_columns;
        first_columns.assign(m, NONE);
        for (size_t j : postorder)
            for (const size_t* rp = a.col_begin(j); rp != a.col_end(j); ++rp)
                if (first_columns[*rp] == NONE) first_columns[*rp] = j;
        std::vector<idx_vec> hadj_f(n);
        for (size_t place = 0; place < n; place++) {
            size_t j = postorder[place];
            for (const size_t* rp = a.col_begin(j); rp != a.col_end(j); ++rp) {
                size_t f = first_columns[*rp];
                if (place > places_in_postorder[f]) hadj_f[f].push_back(j);
            }
        }
        std::vector<ptrdiff_t> vertex_weights(n, 0);
        for (size_t j = 0; j < n; j++) vertex_weights[j] = subtree_size[j] == 1 ? 1 : 0;
        idx_vec& col_counts = out.col_counts;
        col_counts.assign(n, 1);
        idx_vec prev_nbr(n, NONE), prev_f(n, NONE), dsu(n);
        std::iota(dsu.begin(), dsu.end(), (size_t)0);
        auto find = [&](size_t x) {
            size_t root = x;
            while (dsu[root] != root) root = dsu[root];
            while (dsu[x] != root) {
                size_t nx = dsu[x];
                dsu[x] = root;
                x = nx;
            }
            return root;
        };
        for (size_t j_place = 0; j_place < n; j_place++) {
            size_t j = postorder[j_place];
            if (parents[j] != NONE) vertex_weights[parents[j]] -= 1;
            size_t first_desc_place = places_in_postorder[first_descendants[j]];
            for (size_t u : hadj_f[j]) {
                // `prev_nbr[u].wrapping_add(1)`: NONE + 1 == 0 encodes "first time seen".
                if (first_desc_place + 1 > prev_nbr[u] + 1) {
                    vertex_weights[j] += 1;
                    size_t p_leaf = prev_f[u];
                    if (p_leaf != NONE) {
                        size_t q = find(p_leaf);
                        col_counts[u] += levels[j] - levels[q];
                        vertex_weights[q] -= 1;
                    } else {
                        col_counts[u] += levels[j] - levels[u];
                    }
                    prev_f[u] = j;
                }
                prev_nbr[u] = j_place;
            }
            if (parents[j] != NONE) dsu[j] = parents[j];
        }
        for (size_t j = 0; j < n; j++)
            if (parents[j] != NONE) vertex_weights[parents[j]] += vertex_weights[j];
        out.row_counts.resize(n);
        for (size_t j = 0; j < n; j++) out.row_counts[j] = (size_t)vertex_weights[j];
        return out;
    }
};

// cholesky.rs:118-137,359-595
struct CholeskyStructure {
    Structure l_structure;  // pattern of R = Lᵀ, n×n, per column ascending rows, diagonal last
    idx_vec row_permutation;  // length m+n
    Structure h_structure;  // Householder pattern, m×n, rows in permuted numbering

    static CholeskyStructure build(const Structure& a, const idx_vec& parents,
                                   const idx_vec& postorder, const CholeskyCounts& cc) {
        size_t m = a.nrows, n = a.ncols;
        const idx_vec& col_counts = cc.col_counts;
        const idx_vec& first_columns = cc.first_columns;
        CholeskyStructure out;
        size_t m_fictitious = m;
        idx_vec& row_permutation = out.row_permutation;
        row_permutation.assign(m + n, NONE);
        {
            idx_vec next(m, 0), head(n, NONE), tail(n, NONE);
            std::vector<ptrdiff_t> nqueue(n, 0);
            for (size_t ii = m; ii-- > 0;) {
                size_t i = ii, k = first_columns[i];
                if (k == NONE) continue;
                if (nqueue[k] == 0) tail[k] = i;
                nqueue[k] += 1;
                next[i] = head[k];
                head[k] = i;
            }
            for (size_t k = 0; k < n; k++) {
                size_t i;
                if (head[k] == NONE) i = m_fictitious++;
                else i = head[k];
                row_permutation[i] = k;
                nqueue[k] -= 1;
                if (nqueue[k] <= 0) continue;
                size_t parent = parents[k];
                if (parent != NONE) {
                    if (nqueue[parent] == 0) tail[parent] = tail[k];
                    next[tail[k]] = head[parent];
                    head[parent] = next[i];
                    nqueue[parent] += nqueue[k];
                }
            }
            size_t k = n;
            for (size_t i = 0; i < m; i++)
                if (row_permutation[i] == NONE) row_permutation[i] = k++;
        }
        std::vector<idx_vec> h_rows(n);
        size_t num_non_zero = 0;
        for (size_t c : col_counts) num_non_zero += c;
        idx_vec row_indices(num_non_zero, 0), stack, marker(m + n, 0);
        size_t start = 0;
        for (size_t j = 0; j < n; j++) {
            marker[j] = j + 1;
            h_rows[j].push_back(j);
            for (const size_t* rp = a.col_begin(j); rp != a.col_end(j); ++rp) {
                size_t i = *rp;
                size_t k = first_columns[i];
                while (k != NONE && k < j && marker[k] != j + 1) {
                    stack.push_back(k);
                    marker[k] = j + 1;
                    k = parents[k];
                }
                size_t pi = row_permutation[i];
                if (pi > j && marker[pi] < j + 1) {
                    h_rows[j].push_back(pi);
                    marker[pi] = j + 1;
                }
            }
            size_t idx = start;
            while (!stack.empty()) {
                size_t k = stack.back();
                stack.pop_back();
                row_indices[idx++] = k;
                if (parents[k] == j) {
                    for (size_t row : h_rows[k])
                        if (marker[row] < j + 1) {
                            marker[row] = j + 1;
                            h_rows[j].push_back(row);
                        }
                }
            }
            std::sort(row_indices.begin() + start, row_indices.begin() + idx);
            row_indices[idx] = j;
            start += col_counts[j];
        }
        // cholesky.rs:494-561 derive Householder row counts / vertex weights that are not part of
        // the returned structure (dead values); omitted.
        (void)postorder;
        for (auto& hr : h_rows) std::sort(hr.begin(), hr.end());
        out.l_structure.nrows = n;
        out.l_structure.ncols = n;
        out.l_structure.column_pointers.assign(1, 0);
        for (size_t j = 0; j < n; j++)
            out.l_structure.column_pointers.push_back(out.l_structure.column_pointers.back() + col_counts[j]);
        out.l_structure.row_indices = std::move(row_indices);
        out.h_structure.nrows = m;
        out.h_structure.ncols = n;
        out.h_structure.column_pointers.assign(1, 0);
        for (size_t j = 0; j < n; j++) {
            out.h_structure.column_pointers.push_back(out.h_structure.column_pointers.back() + h_rows[j].size());
            out.h_structure.row_indices.insert(out.h_structure.row_indices.end(), h_rows[j].begin(), h_rows[j].end());
        }
        return out;
    }
};

enum class QrOrdering { Natural = 0, Colamd = 1 };

// solvi/src/decomposition/sparse/qr.rs:68-76,118-206
struct SymbolicQr {
    idx_vec row_permutation;
    Structure r_structure, h_structure;
    idx_vec col_permutation, inv_col_permutation;
    // extra artefacts kept for the parity probes
    idx_vec parents, postorder;
    CholeskyCounts counts;
    int colamd_stats[colamd::STATS] = {0};

    static SymbolicQr build(const Structure& a, QrOrdering ordering) {
        SymbolicQr s;
        Structure permuted;
        const Structure* ap = &a;
        if (ordering == QrOrdering::Colamd) {
            size_t a_len = 0;
            if (!colamd::recommended((int)a.row_indices.size(), (int)a.nrows, (int)a.ncols, &a_len))
                throw std::runtime_error("overflow");
            std::vector<int> scratch(a_len, 0);
            for (size_t k = 0; k < a.row_indices.size(); k++) scratch[k] = (int)a.row_indices[k];
            std::vector<int> p(a.column_pointers.size());
            for (size_t k = 0; k < p.size(); k++) p[k] = (int)a.column_pointers[k];
            colamd::Options opt;
            if (!colamd::colamd((int)a.nrows, (int)a.ncols, a_len, scratch.data(), p.data(), opt,
                                s.colamd_stats))
                throw std::runtime_error("valid column ordering");
            s.col_permutation.resize(a.ncols);
            for (size_t k = 0; k < a.ncols; k++) s.col_permutation[k] = (size_t)p[k];
            permuted = a.permute_columns(s.col_permutation);
            ap = &permuted;
            s.inv_col_permutation.assign(a.ncols, 0);
            for (size_t idx = 0; idx < a.ncols; idx++) s.inv_col_permutation[s.col_permutation[idx]] = idx;
        } else {
            s.col_permutation.resize(a.ncols);
            std::iota(s.col_permutation.begin(), s.col_permutation.end(), (size_t)0);
            s.inv_col_permutation = s.col_permutation;
        }
        s.parents = elimination_tree_ata(*ap);
        s.postorder = post_order(s.parents);
        s.counts = CholeskyCounts::build(*ap, s.parents, s.postorder);
        CholeskyStructure cs = CholeskyStructure::build(*ap, s.parents, s.postorder, s.counts);
        s.row_permutation = std::move(cs.row_permutation);
        s.r_structure = std::move(cs.l_structure);
        s.h_structure = std::move(cs.h_structure);
        return s;
    }
};

// qr.rs:226-240
inline void apply_householder(double* x, double beta, const size_t* rows, size_t nrows_h,
                              const double* hv) {
    double tau = 0.0;
    for (size_t k = 0; k < nrows_h; k++) tau = tau + hv[k] * x[rows[k]];
    tau = tau * beta;
    for (size_t k = 0; k < nrows_h; k++) x[rows[k]] = x[rows[k]] - hv[k] * tau;
}

// Same operations in the same order over a 32-bit copy of the row indices (fast mode only: the factor
// is memory bound on the index stream, results do not depend on the index width).
inline void apply_householder32(double* x, double beta, const uint32_t* rows, size_t nrows_h, const double* hv) {
    double tau = 0.0;
    for (size_t k = 0; k < nrows_h; k++) tau = tau + hv[k] * x[rows[k]];
    tau = tau * beta;
    for (size_t k = 0; k < nrows_h; k++) x[rows[k]] = x[rows[k]] - hv[k] * tau;
}

// qr.rs:244-275
inline void calculate_householder(double* v, size_t len, double* norm_out, double* beta_out) {
    double beta, norm;
    double sigma = 0.0;
    for (size_t k = 1; k < len; k++) sigma = sigma + v[k] * v[k];
    if (sigma == 0.0) {
        norm = std::fabs(v[0]);
        beta = v[0] >= 0.0 ? 0.0 : 2.0;
        v[0] = 1.0;
    } else {
        norm = std::sqrt(sigma + v[0] * v[0]);
        if (v[0] <= 0.0) v[0] = v[0] - norm;
        else v[0] = -sigma / (v[0] + norm);
        beta = -(1.0 / (norm * v[0]));
    }
    *norm_out = norm;
    *beta_out = beta;
}

// Process-wide switch of Qr::factorize's scratch handling (0: the reference's per-column fill).
inline std::atomic<int>& qr_fast_mode() {
    static std::atomic<int> mode{0};
    return mode;
}

// qr.rs:91-104,209-223,281-356
struct Qr {
    const SymbolicQr* s;
    SparseColMat r;
    std::vector<double> h_values, h_betas, x;

    explicit Qr(const SymbolicQr& sym) : s(&sym) {
        r.structure = sym.r_structure;
        r.values.assign(sym.r_structure.row_indices.size(), 0.0);
        h_values.assign(sym.h_structure.row_indices.size(), 0.0);
        h_betas.assign(sym.h_structure.ncols, 0.0);
        x.assign(sym.h_structure.nrows, 0.0);
    }

    // `fast` (qr_fast_mode(), test infrastructure for sizes the O(m*n) fill cannot reach): instead of
    // clearing all of x before every column (qr.rs:287) only the positions this column WROTE are reset
    // to +0.0 when the column is done.  x is all +0.0 when a column starts in either mode (constructor /
    // full fill / reset of every written position), every arithmetic operation and its order are the
    // same, so R, H, beta and every later result are bit-identical by construction; the test-suite
    // also checks that on every size the slow mode reaches.
    void factorize(const SparseColMat& a) {
        if (qr_fast_mode().load(std::memory_order_relaxed) != 0) {
            factorize_fast(a);
            return;
        }
        std::fill(r.values.begin(), r.values.end(), 0.0);
        std::fill(h_values.begin(), h_values.end(), 0.0);
        size_t n = a.structure.ncols;
        const Structure& hs = s->h_structure;
        for (size_t j = 0; j < n; j++) {
            std::fill(x.begin(), x.end(), 0.0);  // qr.rs:287 — O(m) per column, kept
            size_t src = s->col_permutation[j];
            for (size_t k = a.structure.column_pointers[src]; k < a.structure.column_pointers[src + 1]; k++)
                x[s->row_permutation[a.structure.row_indices[k]]] = a.values[k];
            size_t rlo = r.structure.column_pointers[j], rhi = r.structure.column_pointers[j + 1];
            for (size_t k = rlo; k < rhi; k++) {
                size_t r_row = r.structure.row_indices[k];
                if (r_row == j) continue;
                size_t hlo = hs.column_pointers[r_row], hhi = hs.column_pointers[r_row + 1];
                apply_householder(x.data(), h_betas[r_row], hs.row_indices.data() + hlo, hhi - hlo,
                                  h_values.data() + hlo);
                r.values[k] = x[r_row];
                x[r_row] = 0.0;
            }
            size_t hlo = hs.column_pointers[j], hhi = hs.column_pointers[j + 1];
            for (size_t k = hlo; k < hhi; k++) {
                h_values[k] = x[hs.row_indices[k]];
                x[hs.row_indices[k]] = 0.0;
            }
            double norm, beta;
            calculate_householder(h_values.data() + hlo, hhi - hlo, &norm, &beta);
            h_betas[j] = beta;
            r.values[rhi - 1] = norm;
        }
    }

    // Fast mode (qr_fast_mode(); test infrastructure for the sizes the O(m*n) fill cannot reach).  The
    // same column loop with two changes that cannot alter a bit of the result:
    //  * instead of clearing all of x before every column (qr.rs:287), the positions this column WROTE
    //    (its scattered entries and the rows of every reflector applied to it) are reset to +0.0 when
    //    the column is done; x is all +0.0 at the start of a column in both modes, and only written
    //    positions can differ from +0.0;
    //  * the reflector rows are read from a 32-bit copy of h_structure.row_indices.
    // Every floating-point operation and its order are those of factorize();
    // tests/test_oracle_fast_mode.py compares R, H-derived solutions and LM traces on every size the
    // slow mode reaches.
    std::vector<uint32_t> h_rows32;
    void factorize_fast(const SparseColMat& a) {
        std::fill(r.values.begin(), r.values.end(), 0.0);
        std::fill(h_values.begin(), h_values.end(), 0.0);
        size_t n = a.structure.ncols;
        const Structure& hs = s->h_structure;
        if (h_rows32.size() != hs.row_indices.size()) {
            if (hs.nrows > 0xFFFFFFFFull) throw std::runtime_error("fast mode: more than 2^32 rows");
            h_rows32.resize(hs.row_indices.size());
            for (size_t k = 0; k < h_rows32.size(); k++) h_rows32[k] = (uint32_t)hs.row_indices[k];
        }
        std::fill(x.begin(), x.end(), 0.0);  // once per factorisation
        for (size_t j = 0; j < n; j++) {
            size_t src = s->col_permutation[j];
            for (size_t k = a.structure.column_pointers[src]; k < a.structure.column_pointers[src + 1]; k++)
                x[s->row_permutation[a.structure.row_indices[k]]] = a.values[k];
            size_t rlo = r.structure.column_pointers[j], rhi = r.structure.column_pointers[j + 1];
            for (size_t k = rlo; k < rhi; k++) {
                size_t r_row = r.structure.row_indices[k];
                if (r_row == j) continue;
                size_t hlo = hs.column_pointers[r_row], hhi = hs.column_pointers[r_row + 1];
                apply_householder32(x.data(), h_betas[r_row], h_rows32.data() + hlo, hhi - hlo, h_values.data() + hlo);
                r.values[k] = x[r_row];
                x[r_row] = 0.0;
            }
            size_t hlo = hs.column_pointers[j], hhi = hs.column_pointers[j + 1];
            for (size_t k = hlo; k < hhi; k++) {
                h_values[k] = x[h_rows32[k]];
                x[h_rows32[k]] = 0.0;
            }
            double norm, beta;
            calculate_householder(h_values.data() + hlo, hhi - hlo, &norm, &beta);
            h_betas[j] = beta;
            r.values[rhi - 1] = norm;
            for (size_t k = a.structure.column_pointers[src]; k < a.structure.column_pointers[src + 1]; k++)
                x[s->row_permutation[a.structure.row_indices[k]]] = 0.0;
            for (size_t k = rlo; k < rhi; k++) {
                size_t r_row = r.structure.row_indices[k];
                if (r_row == j) continue;
                const uint32_t* rw = h_rows32.data() + hs.column_pointers[r_row];
                const size_t cnt = hs.column_pointers[r_row + 1] - hs.column_pointers[r_row];
                for (size_t q = 0; q < cnt; q++) x[rw[q]] = 0.0;
            }
        }
    }

    void q_tr_mul_mut(double* b) const {
        const Structure& hs = s->h_structure;
        std::vector<double> y(hs.nrows, 0.0);
        for (size_t i = 0; i < hs.nrows; i++) y[s->row_permutation[i]] = b[i];
        for (size_t j = 0; j < hs.ncols; j++) {
            size_t hlo = hs.column_pointers[j], hhi = hs.column_pointers[j + 1];
            apply_householder(y.data(), h_betas[j], hs.row_indices.data() + hlo, hhi - hlo,
                              h_values.data() + hlo);
        }
        for (size_t i = 0; i < hs.nrows; i++) b[i] = y[i];
    }

    // b has length m (= h_structure.nrows); on return b[0..n) is the solution in original order.
    bool solve_mut(double* b) const {
        q_tr_mul_mut(b);
        bool solved = r.solve_upper_triangular_mut(b);
        gather_permute(s->inv_col_permutation, b);
        return solved;
    }
};

}  // namespace solvi
}  // namespace orc
