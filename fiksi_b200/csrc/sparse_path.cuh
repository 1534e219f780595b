// Large single-system path (BASELINE config 3): the LM loop of fiksi/src/solve/lm.rs:21-193 for a
// problem whose state does not fit shared memory.  All state lives in HBM; a thin host loop makes
// the accept/reject decisions from three scalars per damping iteration.
//
//   K1/K2  sparse_eval_kernel      one thread per expression row (subsystem.rs:93-166)
//   K3     sparse_assemble_kernel  L := JᵀJ + lam2 I through contribution lists (no atomics),
//          sparse_gradient_kernel  g := -Jᵀ r (one thread per column of the CSC Jacobian)
//   K5     sparse_ldl_kernel       left-looking sparse LDLᵀ in the COLAMD order, one CTA per column,
//                                  scheduled by the elimination tree (a column starts when its
//                                  children are done) inside one persistent launch
//          sparse_forward_kernel / sparse_backward_kernel  sync-free triangular solves, one warp per
//                                  row / column, same tree-driven scheduling
// Replaces Qr::factorize / solve_mut (solvi/src/decomposition/sparse/qr.rs:281-356).
#pragma once
#include <cstdint>

#include "../../include/fiksi_b200.h"
#include "symbolic.hpp"

namespace fk {

class SparseSolver {
public:
    SparseSolver() = default;
    ~SparseSolver();
    SparseSolver(const SparseSolver&) = delete;
    SparseSolver& operator=(const SparseSolver&) = delete;

    // Uploads the symbolic structures of `t` (must outlive the solver) to `device`.
    int init(const Topology& t, int device, std::string* err);
    // == levenberg_marquardt(problem, variables): vars[n_vars], param[n_expr] host arrays,
    // free_values[n_free] in/out.
    int solve(const double* vars, const double* param, double* free_values, fk_report* report, std::string* err);
    // Timings of the last solve (ms, CUDA events): eval, assemble, factor, solve.
    struct Timing { float eval_ms = 0, assemble_ms = 0, factor_ms = 0, tri_ms = 0, transpose_ms = 0, fwd_ms = 0, bwd_ms = 0; uint32_t evals = 0, factors = 0, exact_sums = 0; } last;
    // One residual+Jacobian evaluation at `free_values` (parity probe / assembly-bandwidth metric).
    int eval_once(const double* vars, const double* param, const double* free_values, double* out_r, double* out_j,
                  int repeats, float* ms_per_eval, std::string* err);

private:
    struct Impl;
    Impl* impl_ = nullptr;
};

}  // namespace fk
