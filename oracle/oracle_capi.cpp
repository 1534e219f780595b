// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// extern "C" surface over the CPU restatement (colamd_ref.hpp, solvi_ref.hpp, fiksi_ref.hpp) so
// that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
// drive it through ctypes.  Nothing under fiksi_b200/ links or loads this library.
#include <stdexcept>
#include <chrono>
#include <cstring>
#include <thread>

#include "../include/fiksi_b200.h"  // POD layouts only (fk_problem, fk_report)
#include "fiksi_ref.hpp"

using namespace orc;
using solvi::idx_vec;

#define ORC_API extern "C" __attribute__((visibility("default")))

// ---- colamd --------------------------------------------------------------------------------
ORC_API int orc_colamd_recommended(int nnz, int n_row, int n_col, uint64_t* out) {
    size_t v = 0;
    if (!colamd::recommended(nnz, n_row, n_col, &v)) return 0;
    *out = v;
    return 1;
}
ORC_API int orc_colamd(int n_row, int n_col, uint64_t a_len, int* a, int* p, double dense_row,
                       double dense_col, int aggressive, int* stats) {
    colamd::Options o;
    o.dense_row_control = dense_row;
    o.dense_column_control = dense_col;
    o.aggressive_row_absorption = aggressive != 0;
    return colamd::colamd(n_row, n_col, (size_t)a_len, a, p, o, stats) ? 1 : 0;
}
ORC_API int orc_symamd(int n, const int* a, const int* p, int* perm, double dense_row,
                       double dense_col, int aggressive, int* stats) {
    colamd::Options o;
    o.dense_row_control = dense_row;
    o.dense_column_control = dense_col;
    o.aggressive_row_absorption = aggressive != 0;
    return colamd::symamd(n, a, p, perm, o, stats) ? 1 : 0;
}

// ---- solvi ---------------------------------------------------------------------------------
static solvi::Structure make_structure(uint64_t m, uint64_t n, const uint64_t* colptr, const uint64_t* rowidx) {
    solvi::Structure s;
    s.nrows = m;
    s.ncols = n;
    s.column_pointers.assign(colptr, colptr + n + 1);
    s.row_indices.assign(rowidx, rowidx + colptr[n]);
    return s;
}
ORC_API void orc_from_triplets(uint64_t m, uint64_t n, uint64_t nnz, const uint64_t* rows,
                               const uint64_t* cols, const double* vals, uint64_t* out_shape,
                               uint64_t* colptr, uint64_t* rowidx, double* out_vals) {
    solvi::TripletMat t(m, n);
    for (uint64_t k = 0; k < nnz; k++) t.push_triplet(rows[k], cols[k], vals[k]);
    solvi::SparseColMat c = solvi::SparseColMat::from_triplet_mat(t);
    out_shape[0] = c.structure.nrows;
    out_shape[1] = c.structure.ncols;
    out_shape[2] = c.values.size();
    for (size_t k = 0; k <= c.structure.ncols; k++) colptr[k] = c.structure.column_pointers[k];
    for (size_t k = 0; k < c.values.size(); k++) {
        rowidx[k] = c.structure.row_indices[k];
        out_vals[k] = c.values[k];
    }
}
ORC_API int orc_upper_solve(uint64_t n, const uint64_t* colptr, const uint64_t* rowidx,
                            const double* vals, double* b) {
    solvi::SparseColMat a;
    a.structure = make_structure(n, n, colptr, rowidx);
    a.values.assign(vals, vals + colptr[n]);
    return a.solve_upper_triangular_mut(b) ? 1 : 0;
}
ORC_API void orc_post_order(uint64_t n, const uint64_t* parents, uint64_t* out) {
    idx_vec p(parents, parents + n);
    idx_vec post = solvi::post_order(p);
    for (size_t k = 0; k < n; k++) out[k] = post[k];
}
ORC_API uint64_t orc_node_depth_levels(uint64_t n, const uint64_t* parents, uint64_t* out) {
    idx_vec p(parents, parents + n);
    size_t mx = 0;
    idx_vec lv = solvi::node_depth_levels(p, &mx);
    for (size_t k = 0; k < n; k++) out[k] = lv[k];
    return mx;
}
ORC_API void orc_permute_by_swaps(uint64_t n, const uint64_t* perm, double* data) {
    idx_vec p(perm, perm + n);
    for (auto sw : solvi::permutation_swaps(p)) std::swap(data[sw.first], data[sw.second]);
}
ORC_API void orc_permute_by_gather(uint64_t n, const uint64_t* perm, double* data) {
    idx_vec p(perm, perm + n);
    solvi::gather_permute(p, data);
}

struct SymHandle {
    solvi::Structure a;
    solvi::SymbolicQr sym;
    solvi::Qr* qr = nullptr;
    ~SymHandle() { delete qr; }
};
ORC_API void* orc_sym_build(uint64_t m, uint64_t n, const uint64_t* colptr, const uint64_t* rowidx, int ordering) {
    SymHandle* h = new SymHandle();
    h->a = make_structure(m, n, colptr, rowidx);
    try {
        h->sym = solvi::SymbolicQr::build(h->a, ordering ? solvi::QrOrdering::Colamd : solvi::QrOrdering::Natural);
    } catch (...) {
        delete h;
        return nullptr;
    }
    return h;
}
ORC_API void orc_sym_free(void* h) { delete (SymHandle*)h; }
// which: 0 parents 1 postorder 2 row_counts 3 col_counts 4 r_colptr 5 r_rowidx 6 h_colptr
//        7 h_rowidx 8 row_permutation 9 col_permutation 10 levels 11 first_columns
static const idx_vec* sym_vec(SymHandle* h, int which) {
    switch (which) {
        case 0: return &h->sym.parents;
        case 1: return &h->sym.postorder;
        case 2: return &h->sym.counts.row_counts;
        case 3: return &h->sym.counts.col_counts;
        case 4: return &h->sym.r_structure.column_pointers;
        case 5: return &h->sym.r_structure.row_indices;
        case 6: return &h->sym.h_structure.column_pointers;
        case 7: return &h->sym.h_structure.row_indices;
        case 8: return &h->sym.row_permutation;
        case 9: return &h->sym.col_permutation;
        case 10: return &h->sym.counts.levels;
        case 11: return &h->sym.counts.first_columns;
    }
    return nullptr;
}
ORC_API uint64_t orc_sym_len(void* h, int which) {
    const idx_vec* v = sym_vec((SymHandle*)h, which);
    return v ? v->size() : 0;
}
ORC_API void orc_sym_get(void* h, int which, uint64_t* out) {
    const idx_vec* v = sym_vec((SymHandle*)h, which);
    if (v) for (size_t k = 0; k < v->size(); k++) out[k] = (*v)[k];
}
ORC_API void orc_qr_factorize(void* hh, const double* values) {
    SymHandle* h = (SymHandle*)hh;
    if (!h->qr) h->qr = new solvi::Qr(h->sym);
    solvi::SparseColMat a;
    a.structure = h->a;
    a.values.assign(values, values + h->a.row_indices.size());
    h->qr->factorize(a);
}
// Scratch handling of Qr::factorize: 0 = the reference's per-column O(m) fill (the CPU-baseline cost),
// 1 = reset only the written positions (bit-identical results; reaches config 3).  Process-wide.
ORC_API void orc_set_qr_fast(int on) { solvi::qr_fast_mode().store(on ? 1 : 0); }
ORC_API int orc_get_qr_fast() { return solvi::qr_fast_mode().load(); }
ORC_API void orc_qr_r_values(void* hh, double* out) {
    SymHandle* h = (SymHandle*)hh;
    for (size_t k = 0; k < h->qr->r.values.size(); k++) out[k] = h->qr->r.values[k];
}
ORC_API int orc_qr_solve(void* hh, double* b) { return ((SymHandle*)hh)->qr->solve_mut(b) ? 1 : 0; }

// ---- fiksi: primitives ---------------------------------------------------------------------
ORC_API void orc_rng_u32(uint32_t seed, uint32_t n, uint32_t* out) {
    fiksi::Rng r(seed);
    for (uint32_t k = 0; k < n; k++) out[k] = r.next_u32();
}
ORC_API void orc_rng_f64(uint32_t seed, uint32_t n, double* out) {
    fiksi::Rng r(seed);
    for (uint32_t k = 0; k < n; k++) out[k] = r.next_f64();
}
ORC_API double orc_expr_eval(int kind, double param, const double* vars8, double* grad8) {
    fiksi::Expression e;
    e.kind = (uint8_t)kind;
    e.param = param;
    double v[8], g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 8; k++) v[k] = vars8[k];
    double r = fiksi::compute_residual_and_gradient(e, v, g);
    for (int k = 0; k < 8; k++) grad8[k] = g[k];
    return r;
}
ORC_API int orc_expr_slots(int kind) {
    fiksi::Expression e;
    e.kind = (uint8_t)kind;
    uint32_t vi[8];
    return fiksi::variable_indices(e, vi);
}

// ---- fiksi: flattened-problem LM == levenberg_marquardt(Subsystem) -------------------------
static void fill_report(const fiksi::LmReport& r, fk_report* out) {
    out->exit_reason = r.exit_reason;
    out->outer_iters = r.outer_iters;
    out->factorizations = r.factorizations;
    out->accepted = r.accepted;
    out->ssr = r.ssr;
    out->lambda = r.lambda;
    out->trace_hash = r.trace_hash;
}
struct FlatProblem {
    std::vector<fiksi::Expression> exprs;
    std::vector<uint32_t> free_vars, rows;
    FlatProblem(const fk_problem* p) {
        exprs.resize(p->n_expr);
        for (uint32_t k = 0; k < p->n_expr; k++) {
            exprs[k].kind = p->kind[k];
            for (int q = 0; q < 4; q++) exprs[k].idx[q] = p->idx[4 * k + q];
            exprs[k].param = p->param ? p->param[k] : 0.0;
        }
        free_vars.assign(p->free_vars, p->free_vars + p->n_free);
        rows.assign(p->rows, p->rows + p->n_rows);
    }
};
ORC_API int orc_lm_solve(const fk_problem* p, double* free_values, fk_report* report, char* trace,
                         uint32_t trace_cap) {
    FlatProblem fp(p);
    fiksi::Subsystem sub(p->vars, p->n_vars, fp.exprs.data(), fp.free_vars, fp.rows);
    fiksi::LmReport rep;
    fiksi::levenberg_marquardt(sub, free_values, rep);
    if (report) fill_report(rep, report);
    if (trace && trace_cap) {
        size_t n = std::min((size_t)trace_cap - 1, rep.trace.size());
        memcpy(trace, rep.trace.data(), n);
        trace[n] = 0;
    }
    return 0;
}

// Symbolic probes for a flattened problem: the augmented CSC pattern, the COLAMD permutation, the
// etree and the R pattern exactly as the LM call would build them (lm.rs:80-104).
ORC_API int orc_symbolic(const fk_problem* p, uint32_t* aug_colptr, uint32_t* aug_rowidx,
                         int32_t* colamd_perm, int32_t* etree_parent, uint32_t* r_colptr,
                         uint32_t* r_rowidx, uint32_t* sizes /* aug_nnz, r_nnz */) {
    FlatProblem fp(p);
    fiksi::Subsystem sub(p->vars, p->n_vars, fp.exprs.data(), fp.free_vars, fp.rows);
    size_t nrows = p->n_rows, ncols = p->n_free;
    std::vector<double> x(ncols), res(nrows);
    for (size_t k = 0; k < ncols; k++) x[k] = p->vars[p->free_vars[k]];
    solvi::TripletMat jac(nrows, ncols);
    sub.calculate_residuals_and_sparse_jacobian(x.data(), res.data(), jac);
    for (size_t k = 0; k < ncols; k++) jac.push_triplet(nrows + k, k, 0.);
    solvi::SparseColMat csc = solvi::SparseColMat::from_triplet_mat(jac);
    solvi::SymbolicQr sym = solvi::SymbolicQr::build(csc.structure, solvi::QrOrdering::Colamd);
    if (sizes) {
        sizes[0] = (uint32_t)csc.structure.row_indices.size();
        sizes[1] = (uint32_t)sym.r_structure.row_indices.size();
    }
    if (aug_colptr) for (size_t k = 0; k <= ncols; k++) aug_colptr[k] = (uint32_t)csc.structure.column_pointers[k];
    if (aug_rowidx) for (size_t k = 0; k < csc.structure.row_indices.size(); k++) aug_rowidx[k] = (uint32_t)csc.structure.row_indices[k];
    if (colamd_perm) for (size_t k = 0; k < ncols; k++) colamd_perm[k] = (int32_t)sym.col_permutation[k];
    if (etree_parent) for (size_t k = 0; k < ncols; k++) etree_parent[k] = sym.parents[k] == solvi::NONE ? -1 : (int32_t)sym.parents[k];
    if (r_colptr) for (size_t k = 0; k <= ncols; k++) r_colptr[k] = (uint32_t)sym.r_structure.column_pointers[k];
    if (r_rowidx) for (size_t k = 0; k < sym.r_structure.row_indices.size(); k++) r_rowidx[k] = (uint32_t)sym.r_structure.row_indices[k];
    return 0;
}

// Residuals + Jacobian (CSC values, duplicates merged, no damping rows) of a flattened problem at
// free_values: subsystem.rs:126-166 followed by from_triplet_mat.  out_j has jac_nnz entries in the
// order of the augmented pattern with the damping entries removed.
ORC_API int orc_eval(const fk_problem* p, const double* free_values, double* out_r, double* out_j) {
    FlatProblem fp(p);
    fiksi::Subsystem sub(p->vars, p->n_vars, fp.exprs.data(), fp.free_vars, fp.rows);
    size_t nrows = p->n_rows, ncols = p->n_free;
    solvi::TripletMat jac(nrows, ncols);
    sub.calculate_residuals_and_sparse_jacobian(free_values, out_r, jac);
    if (out_j) {
        for (size_t k = 0; k < ncols; k++) jac.push_triplet(nrows + k, k, 0.);
        solvi::SparseColMat csc = solvi::SparseColMat::from_triplet_mat(jac);
        size_t w = 0;
        for (size_t c = 0; c < ncols; c++)
            for (size_t k = csc.structure.column_pointers[c]; k + 1 < csc.structure.column_pointers[c + 1]; k++)
                out_j[w++] = csc.values[k];
    }
    return 0;
}
ORC_API int orc_residuals(const fk_problem* p, const double* free_values, double* out_r) {
    FlatProblem fp(p);
    fiksi::Subsystem sub(p->vars, p->n_vars, fp.exprs.data(), fp.free_vars, fp.rows);
    sub.calculate_residuals(free_values, out_r);
    return 0;
}

// Uniform batch on `threads` host threads (one sketch per task): the CPU baseline of the bench.
// vars[n][n_vars], param[n][n_expr]; the topology arrays come from `topo` (its vars/param ignored).
// Returns wall seconds.
ORC_API double orc_lm_solve_batch_uniform(const fk_problem* topo, uint32_t n, const double* vars,
                                          const double* param, double* free_out, fk_report* reports,
                                          int threads) {
    if (threads < 1) threads = 1;
    auto work = [&](uint32_t lo, uint32_t hi) {
        FlatProblem fp(topo);
        for (uint32_t s = lo; s < hi; s++) {
            for (uint32_t k = 0; k < topo->n_expr; k++) fp.exprs[k].param = param[(size_t)s * topo->n_expr + k];
            const double* v = vars + (size_t)s * topo->n_vars;
            // Everything the reference does per LM call (Subsystem::new, symbolic analysis, ...)
            // is redone per sketch, as in the reference.
            fiksi::Subsystem sub(v, topo->n_vars, fp.exprs.data(), fp.free_vars, fp.rows);
            double* x = free_out + (size_t)s * topo->n_free;
            for (uint32_t k = 0; k < topo->n_free; k++) x[k] = v[fp.free_vars[k]];
            fiksi::LmReport rep;
            fiksi::levenberg_marquardt(sub, x, rep);
            if (reports) fill_report(rep, &reports[s]);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) {
            uint32_t lo = (uint32_t)((uint64_t)n * t / threads), hi = (uint32_t)((uint64_t)n * (t + 1) / threads);
            pool.emplace_back(work, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- L-BFGS -----------------------------------------------------------------------------------
// == lbfgs(problem, variables) (fiksi/src/solve/lbfgs.rs:20).  report: exit_reason = LbfgsExit,
// outer_iters = line searches, factorizations = function evaluations, ssr, lambda = last step size,
// trace_hash = rolling hash of the evaluations per line search.
ORC_API int orc_lbfgs_solve(const fk_problem* p, double* free_values, fk_report* report) {
    std::vector<fiksi::Expression> ex(p->n_expr);
    for (uint32_t e = 0; e < p->n_expr; e++) {
        ex[e].kind = p->kind[e];
        for (int k = 0; k < 4; k++) ex[e].idx[k] = p->idx[4 * e + k];
        ex[e].param = p->param[e];
    }
    std::vector<uint32_t> fv(p->free_vars, p->free_vars + p->n_free), rows(p->rows, p->rows + p->n_rows);
    fiksi::Subsystem sub(p->vars, p->n_vars, ex.data(), fv, rows);
    fiksi::LbfgsReport rep;
    fiksi::lbfgs(sub, free_values, rep);
    if (report) {
        report->exit_reason = rep.exit_reason; report->outer_iters = rep.iterations; report->factorizations = rep.evaluations;
        report->accepted = rep.iterations; report->ssr = rep.ssr; report->lambda = rep.step; report->trace_hash = rep.trace_hash;
    }
    return 0;
}
// Uniform batch on `threads` host threads; returns seconds.
ORC_API double orc_lbfgs_solve_batch_uniform(const fk_problem* topo, uint32_t n, const double* vars, const double* param,
                                             double* free_out, fk_report* reports, int threads) {
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t s = lo; s < hi; s++) {
            fk_problem p = *topo;
            p.vars = vars + (size_t)s * topo->n_vars;
            p.param = param + (size_t)s * topo->n_expr;
            double* x = free_out + (size_t)s * topo->n_free;
            for (uint32_t k = 0; k < topo->n_free; k++) x[k] = p.vars[topo->free_vars[k]];
            orc_lbfgs_solve(&p, x, reports + s);
        }
    };
    if (threads <= 1) work(0, n);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back(work, (uint32_t)((uint64_t)n * t / threads), (uint32_t)((uint64_t)n * (t + 1) / threads));
        for (auto& th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- Decomposer::SinglePass on a flattened problem ---------------------------------------------------
// The loop of assemble/mod.rs:169-210 for ONE component given as fk_problem (all n_expr expressions form
// the equation graph, free_vars is the component's free set): vars_io[n_vars] in/out (scaled,
// perturbed by the caller), reports: one per strongly connected set, up to cap; returns their number.
ORC_API uint32_t orc_single_pass_problem(const fk_problem* p, double* vars_io, fk_report* reports, uint32_t cap) {
    std::vector<fiksi::Expression> ex(p->n_expr);
    fiksi::ExpressionGraph g;
    g.insert_variables((int)p->n_vars);
    for (uint32_t e = 0; e < p->n_expr; e++) {
        ex[e].kind = p->kind[e];
        for (int k = 0; k < 4; k++) ex[e].idx[k] = p->idx[4 * e + k];
        ex[e].param = p->param[e];
        uint32_t vi[8];
        int a = fiksi::variable_indices(ex[e], vi);
        g.insert_expression(vi, a);
    }
    std::set<uint32_t> free_set(p->free_vars, p->free_vars + p->n_free);
    fiksi::SinglePassPlanner planner(g, free_set);
    uint32_t count = 0;
    for (const fiksi::StronglyConnectedExpressions& scc : planner.plan()) {
        std::vector<double> x;
        for (uint32_t fv : scc.free_variables) x.push_back(vars_io[fv]);
        fiksi::Subsystem sub(vars_io, p->n_vars, ex.data(), scc.free_variables, scc.expressions);
        fiksi::LmReport rep;
        fiksi::levenberg_marquardt(sub, x.data(), rep, false);
        for (size_t k = 0; k < scc.free_variables.size(); k++) vars_io[scc.free_variables[k]] = x[k];
        if (reports && count < cap) {
            fk_report& r = reports[count];
            r.exit_reason = rep.exit_reason; r.outer_iters = rep.outer_iters; r.factorizations = rep.factorizations;
            r.accepted = rep.accepted; r.ssr = rep.ssr; r.lambda = rep.lambda; r.trace_hash = rep.trace_hash;
        }
        count++;
    }
    return count;
}

// ---- System::analyze ------------------------------------------------------------------------
// analyze/numerical/mod.rs:123-147 on a flattened problem: all n_vars variables are columns, all
// n_expr expressions are rows (in order).  out_independent[n_expr].
ORC_API void orc_analyze(uint32_t n_vars, const double* vars, uint32_t n_expr, const uint8_t* kind, const uint32_t* idx,
                         const double* param, uint8_t* out_independent) {
    std::vector<double> v(vars, vars + n_vars);
    std::vector<fiksi::Expression> ex(n_expr);
    for (uint32_t e = 0; e < n_expr; e++) {
        ex[e].kind = kind[e];
        for (int k = 0; k < 4; k++) ex[e].idx[k] = idx[4 * e + k];
        ex[e].param = param[e];
    }
    const std::vector<uint8_t> r = fiksi::analyze_expressions(v, ex);
    memcpy(out_independent, r.data(), r.size());
}
// Gauss-Jordan on its own (row-major matrix in/out, column order in/out).
ORC_API void orc_gauss_jordan(double* matrix, uint64_t nrows, uint64_t ncols, uint64_t* column_indices, uint8_t* out_increases_rank) {
    std::vector<double> m(matrix, matrix + nrows * ncols);
    std::vector<size_t> ci(column_indices, column_indices + ncols);
    const std::vector<uint8_t> r = fiksi::incremental_gauss_jordan_elimination(m, nrows, ncols, ci);
    memcpy(matrix, m.data(), m.size() * sizeof(double));
    for (uint64_t k = 0; k < ncols; k++) column_indices[k] = ci[k];
    memcpy(out_increases_rank, r.data(), r.size());
}

// ---- fiksi: System mirror ------------------------------------------------------------------
ORC_API void* orc_system_new() { return new fiksi::System(); }
ORC_API void orc_system_free(void* s) { delete (fiksi::System*)s; }
#define SYS ((fiksi::System*)s)
ORC_API uint32_t orc_add_point(void* s, double x, double y) { return SYS->add_point(x, y); }
ORC_API uint32_t orc_add_length(void* s, double l) { return SYS->add_length(l); }
ORC_API uint32_t orc_add_line(void* s, uint32_t p1, uint32_t p2) { return SYS->add_line(p1, p2); }
ORC_API uint32_t orc_add_circle(void* s, uint32_t c, uint32_t r) { return SYS->add_circle(c, r); }
ORC_API void orc_fix(void* s, uint32_t e) { SYS->fix(e); }
ORC_API void orc_unfix(void* s, uint32_t e) { SYS->unfix(e); }
ORC_API uint32_t orc_point_point_coincidence(void* s, uint32_t a, uint32_t b) { return SYS->point_point_coincidence(a, b); }
ORC_API uint32_t orc_point_point_distance(void* s, uint32_t a, uint32_t b, double d) { return SYS->point_point_distance(a, b, d); }
ORC_API uint32_t orc_point_point_point_angle(void* s, uint32_t a, uint32_t b, uint32_t c, double ang) { return SYS->point_point_point_angle(a, b, c, ang); }
ORC_API uint32_t orc_point_line_incidence(void* s, uint32_t p, uint32_t l) { return SYS->point_line_incidence(p, l); }
ORC_API uint32_t orc_point_line_distance(void* s, uint32_t p, uint32_t l, double d) { return SYS->point_line_distance(p, l, d); }
ORC_API uint32_t orc_point_circle_incidence(void* s, uint32_t p, uint32_t c) { return SYS->point_circle_incidence(p, c); }
ORC_API uint32_t orc_segment_segment_length_equality(void* s, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return SYS->segment_segment_length_equality(a, b, c, d); }
ORC_API uint32_t orc_line_line_angle(void* s, uint32_t a, uint32_t b, double ang) { return SYS->line_line_angle(a, b, ang); }
ORC_API uint32_t orc_line_line_parallelism(void* s, uint32_t a, uint32_t b) { return SYS->line_line_parallelism(a, b); }
ORC_API uint32_t orc_line_line_perpendicularity(void* s, uint32_t a, uint32_t b) { return SYS->line_line_perpendicularity(a, b); }
ORC_API uint32_t orc_line_circle_tangency(void* s, uint32_t l, uint32_t c) { return SYS->line_circle_tangency(l, c); }
ORC_API uint32_t orc_num_variables(void* s) { return (uint32_t)SYS->variables.size(); }
ORC_API uint32_t orc_num_expressions(void* s) { return (uint32_t)SYS->expressions.size(); }
ORC_API uint32_t orc_num_constraints(void* s) { return (uint32_t)SYS->constraints.size(); }
ORC_API void orc_get_variables(void* s, double* out) { memcpy(out, SYS->variables.data(), SYS->variables.size() * sizeof(double)); }
ORC_API void orc_set_variable(void* s, uint32_t i, double v) { SYS->variables[i] = v; }
ORC_API uint32_t orc_element_variable(void* s, uint32_t e) { return SYS->elements[e].a; }
ORC_API void orc_set_parameter(void* s, uint32_t constraint, double v) { SYS->expressions[SYS->constraints[constraint].expressions_idx].param = v; }
ORC_API double orc_calculate_residual(void* s, uint32_t c) { return SYS->calculate_residual(c); }
ORC_API double orc_system_scale(void* s) { return SYS->calculate_system_scale(); }
ORC_API uint32_t orc_system_analyze(void* s, uint32_t* out, uint32_t cap) {
    const std::vector<uint32_t> d = SYS->analyze();
    for (uint32_t k = 0; k < d.size() && k < cap; k++) out[k] = d[k];
    return (uint32_t)d.size();
}
ORC_API void orc_solve(void* s, int perturb, int keep_artifacts) {
    fiksi::SolvingOptions o;
    o.perturb = perturb != 0;
    SYS->solve(o, keep_artifacts != 0);
}
// Decomposer::SinglePass (assemble/mod.rs:169-210)
ORC_API void orc_solve_single_pass(void* s, int perturb) {
    fiksi::SolvingOptions o;
    o.perturb = perturb != 0;
    SYS->solve_single_pass(o);
}
// The sequence of strongly connected expression sets of all components: call with NULL arrays for the
// sizes (n_steps, total free variables, total expressions), then with arrays.
ORC_API void orc_single_pass_plan(void* s, uint32_t* sizes3, uint32_t* free_ptr, uint32_t* free_vars, uint32_t* expr_ptr,
                                  uint32_t* exprs) {
    std::vector<fiksi::StronglyConnectedExpressions> plan;
    fiksi::System copy = *SYS;  // the dry run still scales / perturbs: keep the caller's system untouched
    fiksi::SolvingOptions o;
    copy.solve_single_pass(o, &plan, true);
    uint32_t nf = 0, ne = 0;
    for (auto& p : plan) { nf += (uint32_t)p.free_variables.size(); ne += (uint32_t)p.expressions.size(); }
    if (sizes3) { sizes3[0] = (uint32_t)plan.size(); sizes3[1] = nf; sizes3[2] = ne; }
    if (!free_ptr || !free_vars || !expr_ptr || !exprs) return;
    uint32_t af = 0, ae = 0;
    for (size_t k = 0; k < plan.size(); k++) {
        free_ptr[k] = af; expr_ptr[k] = ae;
        for (uint32_t v : plan[k].free_variables) free_vars[af++] = v;
        for (uint32_t e : plan[k].expressions) exprs[ae++] = e;
    }
    free_ptr[plan.size()] = af; expr_ptr[plan.size()] = ae;
}
// Decomposer::RecursiveAssembly (assemble/mod.rs:212-277)
// Returns 0, or 1 where the reference would panic (an `unwrap()` on a missing map entry in
// recursive_assembly.rs:352-373 or assemble/mod.rs:374; the restatement uses map::at there).
ORC_API int orc_solve_recursive_assembly(void* s, int perturb) {
    fiksi::SolvingOptions o;
    o.perturb = perturb != 0;
    try {
        fiksi::solve_recursive_assembly(*SYS, o);
    } catch (const std::out_of_range&) {
        return 1;
    }
    return 0;
}
// The recombination plan of all components as one stream of 32-bit words (the same serialisation as
// fk_system_recursive_assembly_plan): per step  n_constraints, constraints..., n_elements, elements..., n_free,
// free elements..., then the three maps (on_frontiers, owned_elements, frontier_elements), each as n_keys and per key
// (ascending)  key, n, values....  Returns the number of words; writes at most `cap`.
ORC_API uint32_t orc_recursive_assembly_plan(void* s, uint32_t* out, uint32_t cap, uint32_t* n_steps) {
    std::vector<fiksi::RecombinationStep> plan;
    fiksi::System copy = *SYS;  // the dry run still scales / perturbs: keep the caller's system untouched
    fiksi::SolvingOptions o;
    try {
        fiksi::solve_recursive_assembly(copy, o, &plan, true);
    } catch (const std::out_of_range&) {  // the reference would panic (see orc_solve_recursive_assembly)
        if (n_steps) *n_steps = 0xFFFFFFFFu;
        return 0;
    }
    std::vector<uint32_t> w;
    auto list = [&](const std::vector<uint32_t>& v) {
        w.push_back((uint32_t)v.size());
        w.insert(w.end(), v.begin(), v.end());
    };
    auto map = [&](const std::map<uint32_t, std::vector<uint32_t>>& m) {
        w.push_back((uint32_t)m.size());
        for (const auto& kv : m) {
            w.push_back(kv.first);
            list(kv.second);
        }
    };
    for (const fiksi::RecombinationStep& st : plan) {
        list(st.constraints); list(st.elements); list(st.free_elements);
        map(st.on_frontiers); map(st.owned_elements); map(st.frontier_elements);
    }
    if (n_steps) *n_steps = (uint32_t)plan.size();
    for (size_t k = 0; k < w.size() && k < cap && out; k++) out[k] = w[k];
    return (uint32_t)w.size();
}
ORC_API uint32_t orc_num_reports(void* s) { return (uint32_t)SYS->last_reports.size(); }
ORC_API void orc_get_report(void* s, uint32_t i, fk_report* out, char* trace, uint32_t cap) {
    const fiksi::LmReport& r = SYS->last_reports[i];
    fill_report(r, out);
    if (trace && cap) {
        size_t n = std::min((size_t)cap - 1, r.trace.size());
        memcpy(trace, r.trace.data(), n);
        trace[n] = 0;
    }
}
ORC_API uint32_t orc_num_components(void* s) { return (uint32_t)SYS->graph.components.size(); }
ORC_API uint32_t orc_component_sizes(void* s, uint32_t ci, uint32_t* n_constraints) {
    *n_constraints = (uint32_t)SYS->graph.components[ci].constraints.size();
    return (uint32_t)SYS->graph.components[ci].elements.size();
}
ORC_API void orc_component_get(void* s, uint32_t ci, uint32_t* elements, uint32_t* constraints) {
    size_t k = 0;
    for (uint32_t e : SYS->graph.components[ci].elements) elements[k++] = e;
    k = 0;
    for (uint32_t c : SYS->graph.components[ci].constraints) constraints[k++] = c;
}

// The flattened problems exactly as `assemble::solve` hands them to `Subsystem::new`, one per
// non-empty component.  Call orc_prepare, then orc_prepared_* getters.  Each problem carries its
// own snapshot of variables_transformed (see for_each_component).
struct Prepared {
    double scale;
    std::vector<double> vars;
    std::vector<uint32_t> free_vars, rows;
};
struct PreparedSet {
    std::vector<Prepared> items;
    std::vector<uint8_t> kind;
    std::vector<uint32_t> idx;
    std::vector<double> param;
};
ORC_API void* orc_prepare(void* s, int perturb) {
    PreparedSet* ps = new PreparedSet();
    fiksi::SolvingOptions o;
    o.perturb = perturb != 0;
    SYS->for_each_component(o, [&](const fiksi::System::ComponentProblem& cp, double scale) {
        Prepared p;
        p.scale = scale;
        p.vars = SYS->variables_transformed;
        p.free_vars = cp.free_variables;
        p.rows = cp.rows;
        ps->items.push_back(std::move(p));
    });
    for (const fiksi::Expression& e : SYS->expressions_transformed) {
        ps->kind.push_back(e.kind);
        for (int q = 0; q < 4; q++) ps->idx.push_back(e.idx[q]);
        ps->param.push_back(e.param);
    }
    return ps;
}
ORC_API void orc_prepared_free(void* ps) { delete (PreparedSet*)ps; }
ORC_API uint32_t orc_prepared_count(void* ps) { return (uint32_t)((PreparedSet*)ps)->items.size(); }
ORC_API double orc_prepared_problem(void* pps, uint32_t i, fk_problem* out) {
    PreparedSet* ps = (PreparedSet*)pps;
    Prepared& p = ps->items[i];
    out->n_vars = (uint32_t)p.vars.size();
    out->vars = p.vars.data();
    out->n_expr = (uint32_t)ps->kind.size();
    out->kind = ps->kind.data();
    out->idx = ps->idx.data();
    out->param = ps->param.data();
    out->n_free = (uint32_t)p.free_vars.size();
    out->free_vars = p.free_vars.data();
    out->n_rows = (uint32_t)p.rows.size();
    out->rows = p.rows.data();
    return p.scale;
}
