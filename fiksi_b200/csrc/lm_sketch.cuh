// K4, second generation: the batched Levenberg-Marquardt solve with ONE THREAD PER SKETCH.
//
// The tile kernel (lm_kernels.cu) gives a sketch 4-32 lanes and a private shared-memory block; ncu showed
// it bound by the SM's load/store pipe and by instruction issue (70 instructions per factor step of at most
// 16 useful FMAs: per-lane table loads, unpacking, masked warp barriers, scattered accesses with bank
// conflicts).  Here the 32 sketches of a warp are interleaved in shared memory (`entry * 32 + lane`):
//   * every access of a warp is 32 consecutive doubles - two conflict-free wavefronts, the minimum;
//   * the op tables are warp-uniform: one copy per CTA in shared memory, read with 16-byte broadcast loads
//     (four table words per instruction, one wavefront) and prefetched one record ahead;
//   * a pivot column's entries are loaded once into registers and reused by all of the column's updates
//     (1 load + 1 store per FMA instead of 3 + 1), and the targets of a column are fetched before the first
//     FMA, so one thread keeps up to 44 independent shared-memory accesses in flight;
//   * the Jacobian is never stored: a row's gradient stays in registers and goes straight into g = -J^T r and
//     H = J^T J; H is factorised in place and an accepted trial point leaves its own H and g behind, so the
//     state of a sketch is 3 n + nnz(L) doubles (a rejected step re-evaluates the accepted point instead of
//     keeping a copy of H);
//   * there is no barrier anywhere: a sketch never leaves its thread.
// Replaces fiksi/src/solve/lm.rs:21-193 + solvi qr.rs:281-356 exactly as the tile kernel does (same normal
// equations, same LDLt operations in the same order, same LM control flow).
#pragma once
#include <cstdint>

#include "../../include/fiksi_b200.h"

namespace fk {

// Launch descriptor (host struct, passed by value).  Field meanings: Topology::SketchTables (symbolic.hpp).
struct SkProgram {
    uint32_t n_vars, n_expr, n, m, lnnz, entries;
    uint32_t xa, xb, w, f, fx, pr, nfix, npar;
    uint32_t off_free, off_fix, off_par, off_eval, off_factor, off_back;
    uint32_t tab_words;
    const uint32_t* tab;  // DEVICE memory, 16-byte aligned; copied into shared memory by every CTA
};

// Raw mode (fk_batch_system_solve): the kernel takes UNSCALED variables and parameters and does assemble::solve's
// pre- and post-processing itself (fiksi/src/assemble/mod.rs:32-44,58-79,113-124,161-166): RMS scale of the sketch
// (sequential sums in the reference's order), variables * (1 / scale), distances * (1 / scale), the seeded perturbation
// of the listed variables, and scale * x on the way out.  All pointers are device memory; raw_vars == nullptr: plain mode.
struct SkRaw {
    const double* raw_vars;     // [n][n_vars]
    const double* raw_param;    // [n][n_expr], or one row when shared_param
    uint32_t shared_param, pad;
    const uint8_t* kinds;       // [n_expr]
    const uint32_t* free_draw;  // [n]: position of free column c in the perturbation list, 0xFFFFFFFF if it draws nothing
    const uint32_t* fix_draw;   // [nfix]: the same for the fixed variables the rows read
    const double* draws;        // [2 x perturbed variables] (fiksi/src/rand.rs:24-39)
    double* scales;             // [n] out, may be null
};
// Raw mode stages a sketch's variables (and its parameters unless shared) in the g and H regions before the first evaluation.
inline bool sk_raw_fits(const SkProgram& p, bool shared_param) {
    return (uint64_t)p.n_vars + (shared_param ? 0u : p.n_expr) <= (uint64_t)p.n + p.lnnz;
}

// Limits: one warp (32 sketches x `entries` doubles) plus the tables must fit an SM's shared memory.
constexpr uint32_t kSkMaxEntries = 840;
constexpr uint32_t kSkMaxTabWords = 16384;
inline bool sk_fits(uint32_t entries, size_t tab_words) {
    return entries <= kSkMaxEntries && tab_words <= kSkMaxTabWords && (size_t)entries * 256 + tab_words * 4 <= 224 * 1024;
}

int launch_batch_lm_sketch(const SkProgram& prog, uint32_t n_sketches, const double* vars, const double* params,
                           double* free_out, fk_report* reports, void* stream, const SkRaw* raw = nullptr);
// Sketches one full wave of the kernel holds on a device with sm_count SMs (callers that cut a batch into chunks
// make the chunks whole waves: with one or two CTAs per SM a partial wave leaves SMs idle for a whole solve).
uint32_t sketch_kernel_wave(const SkProgram& prog, int sm_count);
// 1: the warp-pair shape (two warps share 32 sketches) is the one launched for a batch of this size, 0: one warp per 32 sketches.
int sketch_kernel_is_pair(const SkProgram& prog, uint32_t n_sketches);

}  // namespace fk
