// Batched Levenberg–Marquardt kernels (K1, K2, K4 of SURVEY §2): one tile of TILE lanes (a
// sub-warp, a warp, or a whole CTA) owns one sketch / connected component and runs the complete
// LM loop of fiksi/src/solve/lm.rs:21-193 on it out of shared memory.
#pragma once
#include <cstdint>

#include "../../include/fiksi_b200.h"

namespace fk {

// Device view of a Topology's tables (all pointers device memory, shared by every sketch).
struct DevProgram {
    uint32_t n_vars, n_expr, n, m, jnnz, lnnz, nlevels, pad_;
    const uint8_t* row_kind;    // [m]
    const uint32_t* row_expr;   // [m]
    const uint32_t* slot_var;   // [m][8]
    const int32_t* slot_col;    // [m][8]
    const int32_t* slot_pos;    // [m][8]
    const uint8_t* slot_dup;    // [m][8]
    const uint32_t* free_vars;  // [n]
    const int32_t* perm;        // [n]
    const uint32_t* l_colptr;   // [n+1]
    const uint32_t* l_rowidx;   // [lnnz]
    const uint32_t* h_ptr;      // [lnnz+1]
    const uint32_t* h_pairs;
    const uint32_t* g_ptr;      // [n+1]
    const uint32_t* g_pairs;
    const uint32_t* u_ptr;      // [n+1]
    const uint32_t* u_trip;
    const uint32_t* r_colptr;   // [n+1]
    const uint32_t* r_rowidx;
    const uint32_t* r_lpos;
};

// Shared-memory doubles one sketch needs: x, xs, g, w, invd (5n) + rneg, rs (2m) + H0 (lnnz) +
// work (max(jnnz, lnnz)).
__host__ __device__ inline uint32_t lm_smem_doubles(uint32_t n, uint32_t m, uint32_t jnnz, uint32_t lnnz) {
    return 5u * n + 2u * m + lnnz + (jnnz > lnnz ? jnnz : lnnz);
}

// Launchers (defined in lm_kernels.cu).  `stream` is a cudaStream_t.
int launch_batch_lm(const DevProgram& prog, uint32_t tile, uint32_t n_sketches, const double* vars,
                    const double* params, double* free_out, fk_report* reports, void* stream);
int launch_batch_eval(const DevProgram& prog, uint32_t n_sketches, const double* vars,
                      const double* params, double* out_r, double* out_j, int mode, void* stream);
const char* lm_kernel_name();
// DFMA throughput microbenchmark on the current device (TFLOP/s, 2 flops per DFMA).
int measure_fp64_peak(double* tflops);

}  // namespace fk
