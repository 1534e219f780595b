// See sparse_path.cuh.  Compiled with -fmad=false (the evaluators must not contract); the linear
// algebra uses explicit fma().
#include "sparse_path.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "expressions.cuh"

namespace fk {

namespace {

constexpr uint32_t kNop = 0xFFFFFFFFu;
constexpr int kFactorThreads = 128;
constexpr int kMaxTeamDefault = 16;          // CTAs that may share one column of the factorisation
constexpr uint32_t kTeamWorkDefault = 16384;  // multiply-adds per CTA above which a column is split
constexpr int kFactorWarps = kFactorThreads / 32;

struct SparseDev {
    uint32_t n, m, jnnz, lnnz, n_vars, n_expr, n_hent, acc_cap;
    // evaluation tables (Topology::Tables layout with tile == 1)
    const uint32_t* row_hdr;
    const uint2* row_slots;
    // Jacobian CSC without damping rows (g = -J^T r): column c = entries jcolptr[c]..jcolptr[c+1]
    const uint32_t* jcolptr;
    const uint32_t* jrow;
    const int32_t* perm;      // perm[k] = original column at position k
    // H = JtJ contribution lists, only entries that receive contributions
    const uint32_t* he_pos;   // [n_hent] L position
    const uint32_t* he_ptr;   // [n_hent+1]
    const uint32_t* he_pairs; // 2 per contribution
    // L (CSC, diagonal first) and R = L^T (CSC, diagonal last) patterns
    const uint32_t* l_colptr;
    const uint32_t* l_rowidx;
    const uint32_t* r_colptr;
    const uint32_t* r_rowidx;
    const uint32_t* r_lpos;
    const int32_t* parent;
    const uint32_t* nchildren;
    const uint32_t* order_up;    // columns by ascending height above the leaves
    const uint32_t* order_down;  // reverse
    const uint2* tasks;          // factor tasks {column, part | parts << 16}, by ascending height
    uint32_t n_tasks, pad0;
    const uint64_t* team_off;    // [n] offset of a split column's partial accumulators (doubles), ~0 if not split
    double* team_acc;            // partial accumulators of split columns
    int* arrive;                 // [n] parts of a split column that have delivered
    double* Rval;                // [lnnz] (L D) D^-1 in row-major (R = L^T CSC) order for the forward solve
    // mutable state
    double* Lval;
    double* invd;
    int* done_f;       // [n] column factorised
    int* done_s;       // [n] forward substitution finished for this row
    int* done;         // [n] backward substitution finished for this column
    int* counters;     // [4] task counters + fail flag
    int* rowmap;       // [grid][n]
};

// Release/acquire flags at GPU scope: a producer publishes a finished column / row with
// st.release after its data stores (a barrier first when other threads of the CTA wrote them);
// consumers spin with ld.acquire and then read the data past L1 (__ldcg).
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int* p, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void wait_flag(const int* p) {
    while (ld_acquire(p) == 0) __nanosleep(32);
}

// ---- K1 / K2 -----------------------------------------------------------------------------------
template <bool WITH_JACOBIAN>
__global__ void __launch_bounds__(256)
sparse_eval_kernel(SparseDev S, const double* __restrict__ x, const double* __restrict__ vars,
                   const double* __restrict__ params, double* __restrict__ r, double* __restrict__ J) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < S.m; row += gridDim.x * blockDim.x) {
        const uint32_t hdr = __ldg(S.row_hdr + row);
        const int kind = (int)(hdr & 0xFFu);
        const int a = dev::arity_of(kind);
        uint2 sl[8];
        double v[8], g[8];
#pragma unroll
        for (int s = 0; s < 8; s++) {
            v[s] = 0.0;
            sl[s] = make_uint2(kNop, kNop);
            if (s < a) {
                sl[s] = __ldg(S.row_slots + (size_t)row * 8 + s);
                v[s] = (int32_t)sl[s].x >= 0 ? x[sl[s].x] : __ldg(vars + (sl[s].x & 0x7FFFFFFFu));
            }
        }
        r[row] = dev::eval_expression(kind, v, __ldg(params + (hdr >> 8)), g);
        if (WITH_JACOBIAN) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
                if (s < a && sl[s].y != kNop) {
                    const uint32_t pos = sl[s].y & 0xFFFFFFu;
                    if (sl[s].y & 0x40000000u) J[pos] += g[s];
                    else J[pos] = g[s];
                }
            }
        }
    }
}

// ---- K3 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sparse_assemble_kernel(SparseDev S, const double* __restrict__ J) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < S.n_hent; e += gridDim.x * blockDim.x) {
        const uint32_t b = __ldg(S.he_ptr + e), en = __ldg(S.he_ptr + e + 1);
        double s = 0.0;
        for (uint32_t q = b; q < en; q++) s = fma(J[__ldg(S.he_pairs + 2 * q)], J[__ldg(S.he_pairs + 2 * q + 1)], s);
        S.Lval[__ldg(S.he_pos + e)] = s;
    }
}
// Damping (sqrt(lambda))^2 on the diagonal (lm.rs:119-125) and re-arming of the tree counters for
// the factorisation and the two triangular solves that follow.
__global__ void __launch_bounds__(256) sparse_arm_kernel(SparseDev S, double lam2) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < S.n; k += gridDim.x * blockDim.x) {
        S.Lval[__ldg(S.l_colptr + k)] += lam2;
        S.done_f[k] = 0;
        S.done_s[k] = 0;
        S.done[k] = 0;
        S.arrive[k] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) S.counters[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
sparse_gradient_kernel(SparseDev S, const double* __restrict__ J, const double* __restrict__ r, double* __restrict__ g) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < S.n; k += gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)__ldg(S.perm + k);
        double s = 0.0;
        for (uint32_t q = __ldg(S.jcolptr + c); q < __ldg(S.jcolptr + c + 1); q++) s = fma(J[q], r[__ldg(S.jrow + q)], s);
        g[k] = -s;  // r := -r of lm.rs:86-88 folded into the sign
    }
}

// ---- K5: left-looking LDLt, elimination-tree driven ------------------------------------------------
// Column j of (L D): u(i,j) = H(i,j) - sum_{k in row j of L} u(i,k) * u(j,k) / d_k, i >= j.
// A CTA claims tasks in order of height above the leaves and starts a column once all of its
// children (hence all descendants) are finished.  Heavy columns (the dense chain at the top of the
// tree) are split into up to kMaxTeam tasks that take the k's round-robin; every task sums its
// warps' private shared-memory accumulators in warp order, split columns park the partial sums in
// HBM and the last task to arrive adds them in part order — the result never depends on timing.
__device__ __forceinline__ void finalize_column(const SparseDev& S, uint32_t j, uint32_t p0, uint32_t i, double v) {
    S.Lval[p0 + i] = v;
    if (i == 0) {
        if (v != v) atomicMax(S.counters + 3, 2);
        else if (!(v > 0.0) || v == INFINITY) atomicMax(S.counters + 3, 1);
        S.invd[j] = 1.0 / v;
    }
}

constexpr int kStage = 1024;  // k's of a column staged in shared memory at a time

__global__ void __launch_bounds__(kFactorThreads)
sparse_ldl_kernel(SparseDev S) {
    extern __shared__ double acc[];  // [kFactorWarps][acc_cap]
    __shared__ int sh_task, sh_last;
    __shared__ uint32_t st_k[kStage], st_pos[kStage], st_end[kStage];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int* rowmap = S.rowmap + (size_t)blockIdx.x * S.n;
    double* my_acc = acc + (size_t)warp * S.acc_cap;
    for (;;) {
        if (tid == 0) sh_task = atomicAdd(S.counters + 0, 1);
        __syncthreads();
        const int t = sh_task;
        if (t >= (int)S.n_tasks) break;
        const uint2 task = __ldg(S.tasks + t);
        const uint32_t j = task.x, part = task.y & 0xFFFFu, parts = task.y >> 16;
        const uint32_t p0 = __ldg(S.l_colptr + j), c = __ldg(S.l_colptr + j + 1) - p0;
        for (uint32_t i = tid; i < c; i += kFactorThreads) rowmap[__ldg(S.l_rowidx + p0 + i)] = (int)i;
        for (uint32_t w = 0; w < kFactorWarps; w++)
            for (uint32_t i = tid; i < c; i += kFactorThreads) acc[(size_t)w * S.acc_cap + i] = 0.0;
        const uint32_t r0 = __ldg(S.r_colptr + j), r1 = __ldg(S.r_colptr + j + 1) - 1;  // diagonal is last
        // Work items are 128-entry chunks of the sub-columns (k, rows >= j); chunk ch of the q-th k
        // belongs to warp slot (q + ch) mod slots, so short k's go round-robin and a long k (the
        // dense chain right below j) is spread over every warp of every part.  A warp waits for a
        // column k only when it is about to read it: everything older than the last few columns is
        // long finished, so column j overlaps with the tail of its predecessors (look-ahead).
        const uint32_t slots = parts * kFactorWarps, slot = part * kFactorWarps + warp;
        for (uint32_t base = r0; base < r1; base += kStage) {
            const uint32_t cnt = min((uint32_t)kStage, r1 - base);
            __syncthreads();
            for (uint32_t i = tid; i < cnt; i += kFactorThreads) {
                const uint32_t k = __ldg(S.r_rowidx + base + i);
                st_k[i] = k;
                st_pos[i] = __ldg(S.r_lpos + base + i);
                st_end[i] = __ldg(S.l_colptr + k + 1);
            }
            __syncthreads();
            // 32 k's at a time: every lane tests one k for a chunk owned by this warp (slots is a
            // power of two), then the warp walks the owned ones together
            for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
                const uint32_t i = i0 + lane;
                uint32_t ch0 = 0, nch = 0;
                if (i < cnt) {
                    nch = (st_end[i] - st_pos[i] + 127) >> 7;
                    ch0 = (slot - (base - r0 + i)) & (slots - 1);
                }
                unsigned owned = __ballot_sync(0xFFFFFFFFu, ch0 < nch);
                while (owned) {
                    const int src = __ffs(owned) - 1;
                    owned &= owned - 1;
                    const uint32_t k = st_k[i0 + src], pos = st_pos[i0 + src], end = st_end[i0 + src];
                    const uint32_t nch_k = __shfl_sync(0xFFFFFFFFu, nch, src);
                    uint32_t ch = __shfl_sync(0xFFFFFFFFu, ch0, src);
                    if (lane == 0) wait_flag(S.done_f + k);
                    __syncwarp();
                    const double f = __ldcg(S.Lval + pos) * __ldcg(S.invd + k);
                    for (; ch < nch_k; ch += slots) {
                        const uint32_t e = pos + (ch << 7) + lane;
                        uint32_t row[4];
                        double v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t ee = e + 32 * u;
                            row[u] = ee < end ? __ldg(S.l_rowidx + ee) : kNop;
                            v[u] = ee < end ? __ldcg(S.Lval + ee) : 0.0;
                        }
                        int sl[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) sl[u] = row[u] != kNop ? rowmap[row[u]] : -1;
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (sl[u] >= 0) my_acc[sl[u]] = fma(-v[u], f, my_acc[sl[u]]);
                    }
                }
            }
        }
        __syncthreads();
        bool finished = true;
        if (parts == 1) {
            for (uint32_t i = tid; i < c; i += kFactorThreads) {
                double v = S.Lval[p0 + i];
#pragma unroll
                for (int w = 0; w < kFactorWarps; w++) v += acc[(size_t)w * S.acc_cap + i];
                finalize_column(S, j, p0, i, v);
            }
        } else {
            double* mine = S.team_acc + __ldg(S.team_off + j) + (size_t)part * c;
            for (uint32_t i = tid; i < c; i += kFactorThreads) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < kFactorWarps; w++) v += acc[(size_t)w * S.acc_cap + i];
                mine[i] = v;
            }
            __syncthreads();
            if (tid == 0) sh_last = atom_add_acq_rel(S.arrive + j, 1) == (int)parts - 1;
            __syncthreads();
            finished = sh_last != 0;  // uniform: otherwise another task of the team finishes the column
            if (finished) {
                const double* all = S.team_acc + __ldg(S.team_off + j);
                for (uint32_t i = tid; i < c; i += kFactorThreads) {
                    double v = S.Lval[p0 + i];
                    for (uint32_t p = 0; p < parts; p++) v += __ldcg(all + (size_t)p * c + i);
                    finalize_column(S, j, p0, i, v);
                }
            }
        }
        __syncthreads();
        if (finished && tid == 0) st_release(S.done_f + j, 1);
    }
}

// Row-major copy of the unit factor for the forward solve: Rval[q] = (L D)(i,k) / d_k for the q-th
// entry (k, i) of R = L^T.
__global__ void __launch_bounds__(256) sparse_transpose_kernel(SparseDev S) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        const uint32_t r0 = __ldg(S.r_colptr + i), r1 = __ldg(S.r_colptr + i + 1) - 1;
        for (uint32_t q = r0; q < r1; q++) S.Rval[q] = S.Lval[__ldg(S.r_lpos + q)] * S.invd[__ldg(S.r_rowidx + q)];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Forward substitution with the unit lower factor L' = (L D) D^-1, row oriented, one warp per row:
// y_i = g_i - sum_k L'(i,k) y_k.  Every lane waits for exactly the y_k it is about to read (its k's
// are descendants of i, claimed earlier); the factor values and indices are fetched before the wait.
__global__ void __launch_bounds__(128)
sparse_forward_kernel(SparseDev S, double* __restrict__ w) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(S.counters + 1, 1);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= (int)S.n) break;
        const uint32_t i = __ldg(S.order_up + t);
        const uint32_t r0 = __ldg(S.r_colptr + i), r1 = __ldg(S.r_colptr + i + 1) - 1;
        double s = 0.0;
        for (uint32_t q = r0 + lane; q < r1; q += 128) {
            double a[4], b[4];
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t qq = q + 32 * u;
                a[u] = qq < r1 ? S.Rval[qq] : 0.0;
                k[u] = qq < r1 ? __ldg(S.r_rowidx + qq) : kNop;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                b[u] = 0.0;
                if (k[u] != kNop) {
                    wait_flag(S.done_s + k[u]);
                    b[u] = __ldcg(w + k[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) s = fma(a[u], b[u], s);
        }
        s = warp_sum(s);
        if (lane == 0) {
            w[i] = w[i] - s;
            st_release(S.done_s + i, 1);
        }
    }
}

// Backward substitution D L^T z = y, column oriented gather, one warp per column, parents first.
// The column is walked from its last row (an early ancestor) to its first (the parent, the last
// one to finish), each lane waiting only for the z it reads.
__global__ void __launch_bounds__(128)
sparse_backward_kernel(SparseDev S, double* __restrict__ w, double* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(S.counters + 2, 1);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= (int)S.n) break;
        const uint32_t k = __ldg(S.order_down + t);
        const uint32_t p0 = __ldg(S.l_colptr + k) + 1, p1 = __ldg(S.l_colptr + k + 1);
        double s = 0.0;
        for (uint32_t off = lane; p0 + off < p1; off += 128) {
            double a[4], b[4];
            uint32_t row[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t o = off + 32 * u;
                const bool ok = p0 + o < p1;
                const uint32_t e = p1 - 1 - o;
                a[u] = ok ? S.Lval[e] : 0.0;
                row[u] = ok ? __ldg(S.l_rowidx + e) : kNop;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                b[u] = 0.0;
                if (row[u] != kNop) {
                    wait_flag(S.done + row[u]);
                    b[u] = __ldcg(w + row[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) s = fma(a[u], b[u], s);
        }
        s = warp_sum(s);
        if (lane == 0) {
            const double z = (w[k] - s) * S.invd[k];
            w[k] = z;
            delta[__ldg(S.perm + k)] = z;
            st_release(S.done + k, 1);
        }
    }
}

// ---- deterministic reductions and vector helpers -----------------------------------------------------
constexpr int kRedBlocks = 592, kRedThreads = 256;
__global__ void __launch_bounds__(kRedThreads)
sumsq_partial_kernel(const double* __restrict__ v, uint32_t n, double* __restrict__ partial) {
    __shared__ double sh[kRedThreads];
    double s = 0.0;
    for (uint32_t i = blockIdx.x * kRedThreads + threadIdx.x; i < n; i += kRedBlocks * kRedThreads) s = fma(v[i], v[i], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(1024)
sumsq_final_kernel(const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x < kRedBlocks ? partial[threadIdx.x] : 0.0;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
__global__ void __launch_bounds__(256)
add_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}
__global__ void fetch_flag_kernel(const int* counters, double* scalars) { scalars[2] = (double)counters[3]; }

template <class T>
cudaError_t upload(const std::vector<T>& v, const T** out, std::vector<void*>& owned) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    owned.push_back(d);
    if (!v.empty()) e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T*)d;
    return e;
}

}  // namespace

struct SparseSolver::Impl {
    const Topology* t = nullptr;
    int device = 0;
    SparseDev S{};
    std::vector<void*> owned;
    double *d_x = nullptr, *d_xs = nullptr, *d_vars = nullptr, *d_params = nullptr;
    double *d_r = nullptr, *d_rs = nullptr, *d_J = nullptr, *d_Jt = nullptr, *d_g = nullptr, *d_w = nullptr;
    double *d_delta = nullptr, *d_partial = nullptr, *d_scalars = nullptr;
    double* h_scalars = nullptr;  // pinned: [0] dn, [1] ssr, [2] factor status
    int ldl_grid = 0, sm_count = 0, solve_grid = 148;
    size_t ldl_smem = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    template <class T>
    cudaError_t alloc(T** p, size_t count) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, count) * sizeof(T));
        if (e == cudaSuccess) {
            owned.push_back(d);
            *p = (T*)d;
        }
        return e;
    }
    int grid_for(uint32_t items) const {
        uint32_t b = (items + 255) / 256;
        return (int)std::max<uint32_t>(1, std::min<uint32_t>(b, (uint32_t)sm_count * 8));
    }
    cudaError_t sumsq(const double* v, uint32_t n, double* out) {
        sumsq_partial_kernel<<<kRedBlocks, kRedThreads, 0, stream>>>(v, n, d_partial);
        sumsq_final_kernel<<<1, 1024, 0, stream>>>(d_partial, out);
        return cudaGetLastError();
    }
    void eval(const double* x, double* r, double* J) {
        sparse_eval_kernel<true><<<grid_for(S.m), 256, 0, stream>>>(S, x, d_vars, d_params, r, J);
    }
};

SparseSolver::~SparseSolver() {
    if (!impl_) return;
    cudaSetDevice(impl_->device);
    for (void* p : impl_->owned) cudaFree(p);
    if (impl_->h_scalars) cudaFreeHost(impl_->h_scalars);
    if (impl_->stream) cudaStreamDestroy(impl_->stream);
    for (auto& e : impl_->ev)
        if (e) cudaEventDestroy(e);
    delete impl_;
}

#define SP_CU(call)                                                              \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);  \
            return e_ == cudaErrorMemoryAllocation ? FK_ERR_OOM : FK_ERR_CUDA;   \
        }                                                                        \
    } while (0)

int SparseSolver::init(const Topology& t, int device, std::string* err) {
    impl_ = new Impl();
    Impl& I = *impl_;
    I.t = &t;
    I.device = device;
    SP_CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    SP_CU(cudaGetDeviceProperties(&prop, device));
    I.sm_count = prop.multiProcessorCount;
    I.solve_grid = I.sm_count;
    if (const char* e = std::getenv("FK_SOLVE_GRID")) I.solve_grid = std::max(1, std::atoi(e));
    const uint32_t n = t.n_free, m = t.n_rows, lnnz = (uint32_t)t.l_rowidx.size();
    if (t.jac_nnz >= (1u << 24) || t.n_expr >= (1u << 24)) {
        if (err) *err = "problem exceeds the 24-bit position fields of the evaluation tables";
        return FK_ERR_TOO_LARGE;
    }
    SparseDev& S = I.S;
    S.n = n; S.m = m; S.jnnz = t.jac_nnz; S.lnnz = lnnz; S.n_vars = t.n_vars; S.n_expr = t.n_expr;

    // evaluation tables
    SP_CU(upload(t.tab.row_hdr, &S.row_hdr, I.owned));
    {
        const uint32_t* p = nullptr;
        SP_CU(upload(t.tab.row_slots, &p, I.owned));
        S.row_slots = (const uint2*)p;
    }
    // CSC Jacobian without the damping rows
    std::vector<uint32_t> jcolptr(n + 1), jrow(t.jac_nnz);
    for (uint32_t c = 0; c <= n; c++) jcolptr[c] = t.aug_colptr[c] - c;
    for (uint32_t c = 0; c < n; c++)
        for (uint32_t q = t.aug_colptr[c]; q + 1 < t.aug_colptr[c + 1]; q++) jrow[q - c] = t.aug_rowidx[q];
    SP_CU(upload(jcolptr, &S.jcolptr, I.owned));
    SP_CU(upload(jrow, &S.jrow, I.owned));
    SP_CU(upload(t.perm, &S.perm, I.owned));
    // compacted H contribution lists
    std::vector<uint32_t> he_pos, he_ptr;
    for (uint32_t p = 0; p < lnnz; p++)
        if (t.h_ptr[p + 1] > t.h_ptr[p]) {
            he_pos.push_back(p);
            he_ptr.push_back(t.h_ptr[p]);  // h_pairs is ordered by L position: lists stay contiguous
        }
    he_ptr.push_back(t.h_ptr[lnnz]);
    S.n_hent = (uint32_t)he_pos.size();
    SP_CU(upload(he_pos, &S.he_pos, I.owned));
    SP_CU(upload(he_ptr, &S.he_ptr, I.owned));
    SP_CU(upload(t.h_pairs, &S.he_pairs, I.owned));
    SP_CU(upload(t.l_colptr, &S.l_colptr, I.owned));
    SP_CU(upload(t.l_rowidx, &S.l_rowidx, I.owned));
    SP_CU(upload(t.r_colptr, &S.r_colptr, I.owned));
    SP_CU(upload(t.r_rowidx, &S.r_rowidx, I.owned));
    SP_CU(upload(t.r_lpos, &S.r_lpos, I.owned));
    SP_CU(upload(t.parent, &S.parent, I.owned));
    // tree scheduling data
    std::vector<uint32_t> nchildren(n, 0), height(n, 0), order(n);
    uint32_t max_c = 1;
    for (uint32_t j = 0; j < n; j++) {
        if (t.parent[j] >= 0) {
            nchildren[t.parent[j]]++;
            height[t.parent[j]] = std::max(height[t.parent[j]], height[j] + 1);  // parents have larger indices
        }
        order[j] = j;
        max_c = std::max(max_c, t.l_colptr[j + 1] - t.l_colptr[j]);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return height[a] < height[b]; });
    std::vector<uint32_t> order_down(order.rbegin(), order.rend());
    SP_CU(upload(nchildren, &S.nchildren, I.owned));
    SP_CU(upload(order, &S.order_up, I.owned));
    SP_CU(upload(order_down, &S.order_down, I.owned));
    // work per column = multiply-adds of its left-looking update = sum over the k's of row j of the
    // length of column k from row j down; heavy columns are split into team tasks
    std::vector<uint64_t> work(n, 0);
    for (uint32_t j = 0; j < n; j++)
        for (uint32_t q = t.r_colptr[j]; q + 1 < t.r_colptr[j + 1]; q++)
            work[j] += t.l_colptr[t.r_rowidx[q] + 1] - t.r_lpos[q];
    uint64_t kMaxTeam = kMaxTeamDefault, kTeamWork = kTeamWorkDefault;
    if (const char* e = std::getenv("FK_TEAM_MAX")) kMaxTeam = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("FK_TEAM_WORK")) kTeamWork = std::max(256, std::atoi(e));
    std::vector<uint32_t> tasks;
    std::vector<uint64_t> team_off(n, ~0ull);
    uint64_t team_doubles = 0;
    for (uint32_t j : order) {
        uint32_t parts = (uint32_t)std::min<uint64_t>(kMaxTeam, std::max<uint64_t>(1, (work[j] + kTeamWork - 1) / kTeamWork));
        uint32_t nk = t.r_colptr[j + 1] - t.r_colptr[j] - 1;
        parts = std::max(1u, std::min(parts, std::max(1u, nk / kFactorWarps)));
        while (parts & (parts - 1)) parts &= parts - 1;  // power of two: chunk ownership uses a mask
        if (parts > 1) {
            team_off[j] = team_doubles;
            team_doubles += (uint64_t)parts * (t.l_colptr[j + 1] - t.l_colptr[j]);
        }
        for (uint32_t p = 0; p < parts; p++) {
            tasks.push_back(j);
            tasks.push_back(p | (parts << 16));
        }
    }
    S.n_tasks = (uint32_t)(tasks.size() / 2);
    {
        const uint32_t* p = nullptr;
        SP_CU(upload(tasks, &p, I.owned));
        S.tasks = (const uint2*)p;
    }
    SP_CU(upload(team_off, &S.team_off, I.owned));
    SP_CU(I.alloc(&S.team_acc, (size_t)team_doubles));
    SP_CU(I.alloc(&S.arrive, n));
    SP_CU(I.alloc(&S.Rval, lnnz));

    // factor launch geometry: per-warp accumulators of max column length
    S.acc_cap = max_c;
    I.ldl_smem = (size_t)kFactorWarps * max_c * sizeof(double);
    if (I.ldl_smem > 200 * 1024) {
        if (err) *err = "a column of L is too long for the shared-memory accumulators of the factor kernel";
        return FK_ERR_TOO_LARGE;
    }
    SP_CU(cudaFuncSetAttribute(sparse_ldl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I.ldl_smem));
    int occ = 0;
    SP_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sparse_ldl_kernel, kFactorThreads, I.ldl_smem));
    I.ldl_grid = std::max(1, occ) * I.sm_count;
    if ((size_t)I.ldl_grid * n * sizeof(int) > (size_t)8 << 30) I.ldl_grid = std::max<int>(1, (int)(((size_t)8 << 30) / ((size_t)n * sizeof(int))));

    SP_CU(I.alloc(&S.Lval, lnnz));
    SP_CU(I.alloc(&S.invd, n));
    SP_CU(I.alloc(&S.done_f, n));
    SP_CU(I.alloc(&S.done_s, n));
    SP_CU(I.alloc(&S.done, n));
    SP_CU(I.alloc(&S.counters, 4));
    SP_CU(I.alloc(&S.rowmap, (size_t)I.ldl_grid * n));
    SP_CU(I.alloc(&I.d_x, n));
    SP_CU(I.alloc(&I.d_xs, n));
    SP_CU(I.alloc(&I.d_vars, t.n_vars));
    SP_CU(I.alloc(&I.d_params, t.n_expr));
    SP_CU(I.alloc(&I.d_r, m));
    SP_CU(I.alloc(&I.d_rs, m));
    SP_CU(I.alloc(&I.d_J, t.jac_nnz));
    SP_CU(I.alloc(&I.d_Jt, t.jac_nnz));
    SP_CU(I.alloc(&I.d_g, n));
    SP_CU(I.alloc(&I.d_w, n));
    SP_CU(I.alloc(&I.d_delta, n));
    SP_CU(I.alloc(&I.d_partial, kRedBlocks));
    SP_CU(I.alloc(&I.d_scalars, 4));
    SP_CU(cudaMallocHost((void**)&I.h_scalars, 4 * sizeof(double)));
    SP_CU(cudaStreamCreateWithFlags(&I.stream, cudaStreamNonBlocking));
    SP_CU(cudaEventCreate(&I.ev[0]));
    SP_CU(cudaEventCreate(&I.ev[1]));
    return FK_OK;
}

int SparseSolver::eval_once(const double* vars, const double* param, const double* free_values, double* out_r,
                            double* out_j, int repeats, float* ms_per_eval, std::string* err) {
    Impl& I = *impl_;
    const Topology& t = *I.t;
    SP_CU(cudaSetDevice(I.device));
    SP_CU(cudaMemcpyAsync(I.d_vars, vars, sizeof(double) * t.n_vars, cudaMemcpyHostToDevice, I.stream));
    if (t.n_expr) SP_CU(cudaMemcpyAsync(I.d_params, param, sizeof(double) * t.n_expr, cudaMemcpyHostToDevice, I.stream));
    SP_CU(cudaMemcpyAsync(I.d_x, free_values, sizeof(double) * t.n_free, cudaMemcpyHostToDevice, I.stream));
    I.eval(I.d_x, I.d_r, I.d_J);
    SP_CU(cudaEventRecord(I.ev[0], I.stream));
    for (int k = 0; k < repeats; k++) I.eval(I.d_x, I.d_r, I.d_J);
    SP_CU(cudaEventRecord(I.ev[1], I.stream));
    if (out_r) SP_CU(cudaMemcpyAsync(out_r, I.d_r, sizeof(double) * t.n_rows, cudaMemcpyDeviceToHost, I.stream));
    if (out_j) SP_CU(cudaMemcpyAsync(out_j, I.d_J, sizeof(double) * t.jac_nnz, cudaMemcpyDeviceToHost, I.stream));
    SP_CU(cudaStreamSynchronize(I.stream));
    if (ms_per_eval && repeats > 0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, I.ev[0], I.ev[1]);
        *ms_per_eval = ms / repeats;
    }
    return FK_OK;
}

int SparseSolver::solve(const double* vars, const double* param, double* free_values, fk_report* report, std::string* err) {
    Impl& I = *impl_;
    const Topology& t = *I.t;
    SparseDev& S = I.S;
    const uint32_t n = t.n_free, m = t.n_rows;
    SP_CU(cudaSetDevice(I.device));
    cudaStream_t st = I.stream;
    last = Timing();
    SP_CU(cudaMemcpyAsync(I.d_vars, vars, sizeof(double) * t.n_vars, cudaMemcpyHostToDevice, st));
    if (t.n_expr) SP_CU(cudaMemcpyAsync(I.d_params, param, sizeof(double) * t.n_expr, cudaMemcpyHostToDevice, st));
    SP_CU(cudaMemcpyAsync(I.d_x, free_values, sizeof(double) * n, cudaMemcpyHostToDevice, st));

    double* x = I.d_x;
    double* xs = I.d_xs;
    double* r = I.d_r;
    double* rs = I.d_rs;
    double* J = I.d_J;
    double* Jt = I.d_Jt;

    auto phase = [&](float& acc, auto&& body) -> cudaError_t {
        cudaEventRecord(I.ev[0], st);
        body();
        cudaEventRecord(I.ev[1], st);
        cudaError_t e = cudaEventSynchronize(I.ev[1]);
        if (e != cudaSuccess) return e;
        float ms = 0;
        cudaEventElapsedTime(&ms, I.ev[0], I.ev[1]);
        acc += ms;
        return cudaGetLastError();
    };

    // lm.rs:80-106
    SP_CU(phase(last.eval_ms, [&] { I.eval(x, r, J); }));
    last.evals++;
    SP_CU(I.sumsq(r, m, I.d_scalars + 1));
    sparse_gradient_kernel<<<I.grid_for(n), 256, 0, st>>>(S, J, r, I.d_g);
    SP_CU(cudaMemcpyAsync(I.h_scalars, I.d_scalars, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SP_CU(cudaStreamSynchronize(st));
    double ssr = I.h_scalars[1];

    double lambda = 0.5;
    uint32_t exit_reason = FK_EXIT_MAX_OUTER, outer_iters = 0, factorizations = 0, accepted = 0;
    uint64_t trace = 0;
    auto push = [&](uint32_t code) { trace = trace * 3ull + code + 1ull; };
    bool active = true;
    if (ssr < 1e-8) {
        exit_reason = FK_EXIT_CONVERGED_RESIDUAL;
        active = false;
    } else {
        outer_iters = 1;
    }
    while (active) {
        if (!std::isfinite(lambda)) {
            exit_reason = FK_EXIT_LAMBDA_OVERFLOW;
            break;
        }
        const double sl = std::sqrt(lambda);
        const double lam2 = sl * sl;
        SP_CU(phase(last.assemble_ms, [&] {
            cudaMemsetAsync(S.Lval, 0, sizeof(double) * S.lnnz, st);
            sparse_assemble_kernel<<<I.grid_for(S.n_hent), 256, 0, st>>>(S, J);
            sparse_arm_kernel<<<I.grid_for(n), 256, 0, st>>>(S, lam2);
        }));
        SP_CU(phase(last.factor_ms, [&] { sparse_ldl_kernel<<<I.ldl_grid, kFactorThreads, I.ldl_smem, st>>>(S); }));
        factorizations++;
        last.factors++;
        SP_CU(phase(last.transpose_ms, [&] {
            cudaMemcpyAsync(I.d_w, I.d_g, sizeof(double) * n, cudaMemcpyDeviceToDevice, st);
            sparse_transpose_kernel<<<I.grid_for(n), 256, 0, st>>>(S);
        }));
        SP_CU(phase(last.fwd_ms, [&] { sparse_forward_kernel<<<I.solve_grid, 128, 0, st>>>(S, I.d_w); }));
        SP_CU(phase(last.bwd_ms, [&] { sparse_backward_kernel<<<I.solve_grid, 128, 0, st>>>(S, I.d_w, I.d_delta); }));
        last.tri_ms = last.transpose_ms + last.fwd_ms + last.bwd_ms;
        SP_CU(I.sumsq(I.d_delta, n, I.d_scalars + 0));
        add_kernel<<<I.grid_for(n), 256, 0, st>>>(x, I.d_delta, xs, n);
        SP_CU(phase(last.eval_ms, [&] { I.eval(xs, rs, Jt); }));
        last.evals++;
        SP_CU(I.sumsq(rs, m, I.d_scalars + 1));
        fetch_flag_kernel<<<1, 1, 0, st>>>(S.counters, I.d_scalars);
        SP_CU(cudaMemcpyAsync(I.h_scalars, I.d_scalars, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        SP_CU(cudaStreamSynchronize(st));
        const int fstat = (int)I.h_scalars[2];
        const double dn = I.h_scalars[0];
        double ssr_s = I.h_scalars[1];
        if (fstat == 1) {  // non-positive pivot == the reference's `!solved` (lm.rs:134-137)
            lambda *= 8.0;
            push(0);
            continue;
        }
        if (fstat == 2) ssr_s = NAN;
        if (fstat == 0 && dn < 1e-12) {  // lm.rs:139-142
            exit_reason = FK_EXIT_SMALL_STEP;
            break;
        }
        if (ssr_s < ssr) {
            lambda *= 0.125;
            if (lambda < 1e-50) lambda = 1e-50;
            accepted++;
            push(1);
            std::swap(x, xs);
            std::swap(r, rs);
            std::swap(J, Jt);
            const bool stalled = (ssr - ssr_s) / ssr <= 1e-6;
            ssr = ssr_s;
            if (stalled) {
                exit_reason = FK_EXIT_STALLED;
                break;
            }
            sparse_gradient_kernel<<<I.grid_for(n), 256, 0, st>>>(S, J, r, I.d_g);
            if (outer_iters == 100) break;
            if (ssr < 1e-8) {
                exit_reason = FK_EXIT_CONVERGED_RESIDUAL;
                break;
            }
            outer_iters++;
        } else {
            lambda *= 2.0;
            push(2);
        }
    }
    SP_CU(cudaMemcpyAsync(free_values, x, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    SP_CU(cudaStreamSynchronize(st));
    if (report) {
        report->exit_reason = exit_reason;
        report->outer_iters = outer_iters;
        report->factorizations = factorizations;
        report->accepted = accepted;
        report->ssr = ssr;
        report->lambda = lambda;
        report->trace_hash = trace;
    }
    return FK_OK;
}

}  // namespace fk
