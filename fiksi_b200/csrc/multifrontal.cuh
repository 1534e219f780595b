// K5, second generation: supernodal multifrontal LDLᵀ of H = JᵀJ + lambda I in the COLAMD order,
// and the two triangular solves, for the large single-system path (BASELINE config 3).
//
// Replaces Qr::factorize / Qr::solve_mut (solvi/src/decomposition/sparse/qr.rs:281-356) on the
// pattern that SymbolicQr::build produces (qr.rs:118-206; R = Lᵀ, cholesky.rs:359-595).
//
// Symbolic side (host, once per topology): columns of L with nested patterns are merged into
// supernodes; each supernode s owns a dense frontal matrix of order f_s = |struct(first column)|
// whose first ns_s columns (the "panel", stored column-major with leading dimension f_s) become
// columns of L and whose trailing (f_s-ns_s)² block U_s is the update ("Schur complement") that
// its parent in the supernodal elimination tree assembles through a relative-index list.
//
// Numeric side: complete subtrees of small fronts (f <= 32) are factorised by ONE WARP each with the
// front in shared memory (one launch for all of them); the remaining "big" part of the tree — for
// the 400x250 lattice 6.6 k supernodes in 56 levels holding 99 % of the flops — runs level by level:
// an assembly kernel (extend-add of the children, parallel over disjoint target column blocks, so
// no atomics and a fixed summation order), then per 64-column pivot block a diagonal-tile kernel
// (left-looking update, LDLᵀ of the 64x64 tile in shared memory, inverse of its unit factor) and a
// panel kernel (left-looking update + multiplication with that inverse), then one kernel that
// applies the rank-ns update to all 64x64 tiles of U.  All tile kernels share one register-blocked
// FP64 micro-kernel (4x4 per thread, operands staged through shared memory).  The launch sequence is
// static per topology and is replayed as a CUDA graph.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

namespace fk {

struct Topology;

struct MfDev {
    uint32_t S = 0;               // supernodes
    const uint32_t* c0 = nullptr;        // [S] first column (permuted index)
    const uint32_t* ns = nullptr;        // [S] pivot columns
    const uint32_t* f = nullptr;         // [S] front order
    const uint32_t* rows_off = nullptr;  // [S+1]
    const uint32_t* rows = nullptr;      // front row lists (ascending permuted column indices)
    const uint64_t* pan_off = nullptr;   // [S] panel offset (doubles)
    const uint64_t* upd_off = nullptr;   // [S] update-matrix offset (doubles)
    const uint32_t* rel_off = nullptr;   // [S+1] offsets into rel and into ubuf
    const uint32_t* rel = nullptr;       // position of every update row in the parent's front
    const uint32_t* child_ptr = nullptr; // [S+1]
    const uint32_t* child = nullptr;     // children in ascending order
    const uint32_t* winv_blk = nullptr;  // [S] index of the supernode's first 64x64 inverse block
    double* pan = nullptr;        // panels: L below the diagonal (unit lower), D on the diagonal
    double* upd = nullptr;        // update matrices (lower triangles used)
    double* winv = nullptr;       // D^-1 L^-1 of every 64x64 diagonal tile of the big fronts
    double* ubuf = nullptr;       // update vectors of the forward solve
    int* status = nullptr;        // 0 ok, 1 non-positive pivot, 2 NaN pivot, 3 a polling loop gave up (internal error)
};

// Host-side symbolic structures of the multifrontal factorisation (everything init() uploads).
struct MfSymbolic {
    uint32_t n = 0, S = 0, nsub = 0, nlevels = 0, nbig = 0, max_front = 1, winv_blocks = 0;
    uint64_t pan_total = 0, upd_total = 0;
    std::vector<uint32_t> c0, ns, f, rows_off, rows, rel_off, rel, child_ptr, child, winv_blk, sub_ptr, sub_list, level_list, level, tasks;
    std::vector<uint64_t> pan_off, upd_off;
    std::vector<int32_t> sparent;
    std::vector<uint8_t> big;  // 1: level-scheduled tiled path, 0: member of a one-warp subtree
};

class Multifrontal {
public:
    struct Launch { int kind; uint32_t first, count; };  // kind: 0 asm, 1 diag, 2 col, 3 right-looking update
    MfSymbolic sym;
    // Host-only symbolic analysis (no CUDA calls).
    cudaError_t build_symbolic(const Topology& t, std::string* err);
    struct LaunchInfo { int kind; uint32_t first, count; };
    std::vector<LaunchInfo> factor_launches() const {
        std::vector<LaunchInfo> v;
        for (const Launch& l : factor_seq_) v.push_back({l.kind, l.first, l.count});
        return v;
    }
    ~Multifrontal();
    // Builds the supernodal structures from the topology's L pattern and uploads them.
    cudaError_t init(const Topology& t, cudaStream_t stream, std::string* err);
    // Entry k of L column j (Topology::l_rowidx order: diagonal first) lives at diag_panel()[j] + k in the panel storage.
    const std::vector<uint64_t>& diag_panel() const { return diag_map_; }  // [n] panel offset of d_k
    size_t panel_doubles() const { return pan_total_; }
    double* panels() const { return dev_.pan; }
    int* status() const { return dev_.status; }
    // Numeric factorisation of the assembled panels (H entries in place, zeros elsewhere).
    cudaError_t factor(cudaStream_t stream);
    // w (permuted right-hand side, overwritten with the permuted solution); delta[perm[k]] = z[k].
    cudaError_t solve(double* w, double* delta, const int32_t* d_perm, cudaStream_t stream);
    uint64_t flops() const { return flops_; }
    uint32_t launches_per_factor() const { return (uint32_t)factor_launches_; }

    struct Stats { uint32_t supernodes = 0, small_subtrees = 0, big = 0, levels = 0, max_front = 0; uint64_t upd_doubles = 0; } stats;

private:
    MfDev dev_;
    std::vector<void*> owned_;
    std::vector<uint64_t> diag_map_;
    size_t pan_total_ = 0;
    uint64_t flops_ = 0;
    uint32_t n_ = 0, nsub_ = 0;
    const uint32_t *d_sub_ptr_ = nullptr, *d_sub_list_ = nullptr;
    const uint4* d_tasks_ = nullptr;
    std::vector<Launch> factor_seq_;
    size_t factor_launches_ = 0;
    // big supernodes by level (solve kernels)
    const uint32_t* d_level_list_ = nullptr;
    std::vector<uint32_t> level_ptr_;
    std::vector<std::pair<uint32_t, uint32_t>> fwd_tasks_, bwd_tasks_;  // per level: {first task, count}
    double* d_tmp_ = nullptr;
    std::vector<bool> level_wide_;  // a supernode of the level has >= 64 pivot columns
    std::vector<bool> level_narrow_;  // every front of the level has order <= 64
    // chained solves of the wide levels near the root (64-row chunks, one CTA each, values polled in place)
    struct ChainLevel { bool on = false; uint32_t fwd_first = 0, fwd_count = 0, dot_first = 0, dot_count = 0, bwd_first = 0, bwd_count = 0, flow_first = 0, flow_count = 0, asm_first = 0; };
    std::vector<ChainLevel> chain_;
    std::vector<uint32_t> chain_tasks_;
    const uint4* d_chain_tasks_ = nullptr;
    double* d_chain_pub_ = nullptr;  // [2][n] published y (forward) and z (backward), polled in place
    uint32_t n_chain_flags_ = 0, inv_first_ = 0, inv_count_ = 0;
    // tile dataflow factorisation of the chained levels: published panel tiles [flow_pub_lo_, flow_pub_hi_) of the panel storage
    std::vector<uint32_t> level_seq_ptr_;
    double* d_pan_pub_ = nullptr;
    uint32_t* d_tickets_ = nullptr;  // task tickets of the polling launches (see take_ticket)
    uint32_t n_tickets_ = 0;
    // levels whose fronts are all small (<= 72 rows): one launch of mf_mid_factor_kernel per level, a CTA per front
    struct MidLevel { uint32_t first = 0, count = 0, ld = 0; };
    std::vector<MidLevel> mid_;
    std::vector<uint32_t> mid_list_;
    const uint32_t* d_mid_list_ = nullptr;
    std::vector<uint32_t> flow_asm_ptr_, flow_asm_;  // per dataflow tile: extend-add ranges of its children (uint4 entries)
    const uint32_t* d_flow_asm_ptr_ = nullptr;
    const uint4* d_flow_asm_ = nullptr;
    uint64_t flow_pub_lo_ = ~0ull, flow_pub_hi_ = 0;
    cudaGraphExec_t factor_graph_ = nullptr, solve_graph_ = nullptr;
    double* solve_w_ = nullptr; double* solve_delta_ = nullptr; const int32_t* solve_perm_ = nullptr;
    cudaError_t enqueue_factor(cudaStream_t stream);
    cudaError_t enqueue_solve(double* w, double* delta, const int32_t* d_perm, cudaStream_t stream);
};

}  // namespace fk
