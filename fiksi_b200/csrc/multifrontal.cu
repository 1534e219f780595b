// See multifrontal.cuh.  Compiled with -fmad=false like the rest of the library; the linear algebra
// uses explicit fma().
#include "multifrontal.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "symbolic.hpp"

namespace fk {

namespace {

constexpr int kSmallFront = 32;   // fronts up to this order are factorised by one warp in shared memory
constexpr int kSmallLd = kSmallFront + 1;
constexpr int kWarpsPerCta = 8;
constexpr int TB = 64;            // tile order of the big path
constexpr int KC = 16;            // pivot columns staged per step of the micro-kernel
constexpr int kTileThreads = 256;
constexpr uint32_t kChainMaxCtas = 1024;  // chained solve levels (128 while the chunks had to be co-resident; config 3: 5 solves 6.73 -> 6.39 ms)

__device__ __forceinline__ void flag_pivot(int* status, double d) {
    if (d != d) atomicMax(status, 2);
    else if (!(d > 0.0) || d == INFINITY) atomicMax(status, 1);
}

// 1/d: hardware seed (about 20 bits) + one third-order step r (1 + e + e^2), e = 1 - d r: relative error ~e^3 < 1e-17
// before rounding, three dependent FMAs (two Newton steps are four), no slow path, no branches.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// ---- small subtrees: one warp per subtree, supernodes in ascending (= topological) order ---------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
mf_small_factor_kernel(MfDev D, const uint32_t* __restrict__ sub_ptr, const uint32_t* __restrict__ sub_list, uint32_t nsub) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sub = blockIdx.x * kWarpsPerCta + warp;
    if (sub >= nsub) return;
    double* F = sm + warp * (kSmallFront * kSmallLd);  // F[i][j], j <= i
    const uint32_t i = lane;
    for (uint32_t q = __ldg(sub_ptr + sub); q < __ldg(sub_ptr + sub + 1); q++) {
        const uint32_t s = __ldg(sub_list + q);
        const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), r = f - ns;
        double* P = D.pan + __ldg(D.pan_off + s);
        if (i < f)
            for (uint32_t j = 0; j <= i; j++) F[i * kSmallLd + j] = j < ns ? P[(size_t)j * f + i] : 0.0;
        __syncwarp();
        // extend-add of the children's update matrices, children in ascending order
        for (uint32_t cq = __ldg(D.child_ptr + s); cq < __ldg(D.child_ptr + s + 1); cq++) {
            const uint32_t c = __ldg(D.child + cq);
            const uint32_t rc = __ldg(D.f + c) - __ldg(D.ns + c);
            const double* U = D.upd + __ldg(D.upd_off + c);
            const uint32_t ta = i < rc ? __ldg(D.rel + __ldg(D.rel_off + c) + i) : 0u;
            for (uint32_t b = 0; b < rc; b++) {
                const uint32_t tb = __shfl_sync(0xFFFFFFFFu, ta, b);
                if (i >= b && i < rc) F[ta * kSmallLd + tb] += U[(size_t)b * rc + i];
            }
            __syncwarp();
        }
        // LDLt of the pivot columns; the trailing block receives the Schur complement
        for (uint32_t k = 0; k < ns; k++) {
            const double d = F[k * kSmallLd + k];
            if (lane == 0) flag_pivot(D.status, d);
            const double inv = fast_rcp(d);
            double li = 0.0;
            if (i > k && i < f) {
                li = F[i * kSmallLd + k] * inv;
                double* row = F + i * kSmallLd;
                uint32_t j = k + 1;
                for (; j + 3 <= i; j += 4) {  // four entries per step, loads before stores (independent chains in flight)
                    double c[4], v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        c[u] = F[(j + u) * kSmallLd + k];
                        v[u] = row[j + u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) row[j + u] = fma(-li, c[u], v[u]);
                }
                for (; j <= i; j++) row[j] = fma(-li, F[j * kSmallLd + k], row[j]);
            }
            __syncwarp();
            if (i > k && i < f) F[i * kSmallLd + k] = li;
            __syncwarp();
        }
        if (i < f)
            for (uint32_t j = 0; j <= i && j < ns; j++) P[(size_t)j * f + i] = F[i * kSmallLd + j];
        if (i >= ns && i < f) {
            double* U = D.upd + __ldg(D.upd_off + s);
            for (uint32_t j = ns; j <= i; j++) U[(size_t)(j - ns) * r + (i - ns)] = F[i * kSmallLd + j];
        }
        __syncwarp();
    }
}

// ---- mid-size fronts: one CTA per supernode, the whole front in shared memory -----------------------------
// The first levels above the small subtrees hold thousands of supernodes with fronts of 33..72 rows and 5..14 pivot
// columns (400x250 lattice: 3,277 / 1,630 supernodes with fronts <= 48 / 72).  The tile kernels spend five
// launches and a 64x64-tile machinery per level on them (0.39 ms for 0.13 GFLOP); here one CTA assembles the front
// (children in ascending order, distinct targets within a child: no atomics, fixed summation order), eliminates its
// pivot columns right-looking (rows across warps, columns across lanes, one barrier per column: scaling column k
// touches nothing the update with column k + 1 reads) and writes the panel and the update matrix back.
constexpr int kMidFront = 72;  // (fronts up to 128 rows were tried: above ~72 rows the 64x64-tile kernels are faster, 245 us against 176 us on the third level of the lattice)
constexpr int kMidThreads = 256;
constexpr int kMidChildren = 64;  // children's descriptors staged per batch
__global__ void __launch_bounds__(kMidThreads)
mf_mid_factor_kernel(MfDev D, const uint32_t* __restrict__ list, uint32_t ld) {
    extern __shared__ __align__(16) double Fm[];  // F[i][j], j <= i, row stride ld (odd)
    const uint32_t s = __ldg(list + blockIdx.x);
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), r = f - ns;
    double* P = D.pan + __ldg(D.pan_off + s);
    double* U = D.upd + __ldg(D.upd_off + s);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t NW = kMidThreads / 32;
    for (uint32_t j = warp; j < f; j += NW) {
        if (j < ns) {
            const double* col = P + (size_t)j * f;
            for (uint32_t i = j + lane; i < f; i += 32) Fm[i * ld + j] = col[i];
        } else {
            for (uint32_t i = j + lane; i < f; i += 32) Fm[i * ld + j] = 0.0;
        }
    }
    // children's descriptors first (one round trip for all of them instead of three dependent ones per child)
    __shared__ uint32_t ch_rc[kMidChildren], ch_rel[kMidChildren];
    __shared__ uint64_t ch_upd[kMidChildren];
    const uint32_t cq0 = __ldg(D.child_ptr + s), nch = __ldg(D.child_ptr + s + 1) - cq0;
    for (uint32_t base = 0; base < nch; base += kMidChildren) {
        const uint32_t nb = min((uint32_t)kMidChildren, nch - base);
        __syncthreads();  // (the front is loaded / the previous batch's descriptors are no longer read)
        if (tid < nb) {
            const uint32_t c = __ldg(D.child + cq0 + base + tid);
            ch_rc[tid] = __ldg(D.f + c) - __ldg(D.ns + c);
            ch_rel[tid] = __ldg(D.rel_off + c);
            ch_upd[tid] = __ldg(D.upd_off + c);
        }
        __syncthreads();
        for (uint32_t q = 0; q < nb; q++) {
            const uint32_t rc = ch_rc[q];
            const double* Uc = D.upd + ch_upd[q];
            const uint32_t* relc = D.rel + ch_rel[q];
            for (uint32_t b = warp; b < rc; b += NW) {
                const uint32_t tb = __ldg(relc + b);
                const double* src = Uc + (size_t)b * rc;
                uint32_t a = b + lane;
                for (; a + 32 < rc; a += 64) {  // two independent gather / add chains in flight
                    const uint32_t t0 = __ldg(relc + a), t1 = __ldg(relc + a + 32);
                    const double v0 = src[a], v1 = src[a + 32];
                    const double f0 = Fm[t0 * ld + tb], f1 = Fm[t1 * ld + tb];
                    Fm[t0 * ld + tb] = f0 + v0;
                    Fm[t1 * ld + tb] = f1 + v1;
                }
                for (; a < rc; a += 32) Fm[__ldg(relc + a) * ld + tb] += src[a];
            }
            __syncthreads();
        }
    }
    for (uint32_t k = 0; k < ns; k++) {
        const double d = Fm[k * ld + k];
        if (tid == 0) flag_pivot(D.status, d);
        const double inv = fast_rcp(d);
        // four rows per warp and step: the column entry F[j][k] is read once for the four, and the four chains are independent
        for (uint32_t i0 = k + 1 + 4 * warp; i0 < f; i0 += 4 * NW) {
            double li[4];
            double* rowp[4];
            uint32_t last[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = min(i0 + (uint32_t)u, f - 1);
                rowp[u] = Fm + i * ld;
                li[u] = rowp[u][k] * inv;
                last[u] = i0 + (uint32_t)u < f ? i : 0u;  // (0: the clamped duplicate of the last row updates nothing)
            }
            const uint32_t jmax = min(i0 + 3u, f - 1);
            for (uint32_t j = k + 1 + lane; j <= jmax; j += 32) {
                const double fjk = Fm[j * ld + k];
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = rowp[u][j];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = fma(-li[u], fjk, v[u]);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (j <= last[u]) rowp[u][j] = v[u];
            }
        }
        __syncthreads();
        for (uint32_t i = k + 1 + tid; i < f; i += kMidThreads) Fm[i * ld + k] *= inv;
    }
    __syncthreads();
    for (uint32_t j = warp; j < f; j += NW) {
        if (j < ns) {
            double* col = P + (size_t)j * f;
            for (uint32_t i = j + lane; i < f; i += 32) col[i] = Fm[i * ld + j];
        } else {
            double* col = U + (size_t)(j - ns) * r - ns;
            for (uint32_t i = j + lane; i < f; i += 32) col[i] = Fm[i * ld + j];
        }
    }
}

// ---- big path -----------------------------------------------------------------------------------------
// Assembly of one block of front columns [lo, hi) of supernode s: zero the block's part of U_s,
// then add the children's update matrices in ascending child order.  Blocks of one front have
// disjoint targets, so the kernel needs no atomics and the summation order is fixed.
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
mf_asm_kernel(MfDev D, const uint4* __restrict__ tasks) {
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, lo = t.y, hi = t.z;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), r = f - ns;
    double* P = D.pan + __ldg(D.pan_off + s);
    double* U = D.upd + __ldg(D.upd_off + s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t col = max(lo, ns) + warp; col < hi; col += 8)
        for (uint32_t row = col + lane; row < f; row += 32) U[(size_t)(col - ns) * r + (row - ns)] = 0.0;
    __syncthreads();
    for (uint32_t cq = __ldg(D.child_ptr + s); cq < __ldg(D.child_ptr + s + 1); cq++) {
        const uint32_t c = __ldg(D.child + cq);
        const uint32_t rc = __ldg(D.f + c) - __ldg(D.ns + c);
        const double* Uc = D.upd + __ldg(D.upd_off + c);
        const uint32_t* relc = D.rel + __ldg(D.rel_off + c);
        // child columns whose target lies in [lo, hi): rel is ascending, so the range is
        // [#entries < lo, #entries < hi) -- counted by the whole CTA in parallel (a binary search
        // would be ten dependent L2 round trips per bound)
        uint32_t b_lo = 0, b_hi = 0;
        for (uint32_t base = 0; base < rc; base += 256) {
            const uint32_t v = base + threadIdx.x < rc ? __ldg(relc + base + threadIdx.x) : 0xFFFFFFFFu;
            b_lo += (uint32_t)__syncthreads_count(v < lo);
            b_hi += (uint32_t)__syncthreads_count(v < hi);
        }
        for (uint32_t b = b_lo + warp; b < b_hi; b += 8) {
            const uint32_t tb = __ldg(relc + b);
            double* dst = tb < ns ? P + (size_t)tb * f : U + (size_t)(tb - ns) * r - ns;
            const double* __restrict__ src = Uc + (size_t)b * rc;
            uint32_t a = b + lane;
            for (; a + 96 < rc; a += 128) {  // four independent gather / add / scatter chains in flight
                uint32_t ix[4];
                double sv[4], dv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    ix[u] = __ldg(relc + a + 32 * u);
                    sv[u] = src[a + 32 * u];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) dv[u] = dst[ix[u]];
#pragma unroll
                for (int u = 0; u < 4; u++) dst[ix[u]] = dv[u] + sv[u];
            }
            for (; a < rc; a += 32) dst[__ldg(relc + a)] += src[a];
        }
        __syncthreads();
    }
}

// Register-blocked FP64 micro-kernel: acc(4x4 per thread, 64x64 per CTA) -= A * diag(d) * B^T over
// K pivot columns of the panel P (column-major, leading dimension f), A = rows rowA0..rowA0+nrA,
// B = rows rowB0..rowB0+nrB.  Thread tx = tid & 15 owns the INTERLEAVED rows tx, tx+16, tx+32, tx+48
// (a warp's A-operand loads are 16 consecutive doubles: one conflict-free wavefront each; with four
// consecutive rows per thread the 16-byte loads hit only half of the banks, 4 wavefronts each, and
// ncu showed the kernel stalled on shared memory), thread ty = tid >> 4 owns columns 4ty..4ty+3
// (two broadcast 16-byte loads).  Chunks of KC columns go through double-buffered shared memory;
// the next chunk is fetched into registers while the current one is multiplied, and the operands
// of step k+1 are loaded while step k is multiplied (explicit register double buffering).
// Published values are polled in place (see the chained solves / the tile dataflow factorisation below): a buffer is
// filled with an all-ones bit pattern (a NaN no computation produces), a producer stores each 8-byte value
// once, and a consumer re-reads its values at GPU scope until none of them is the sentinel.  A value is either
// absent or final, so no fence, flag or second round trip is needed.
__device__ __forceinline__ double ld_relaxed(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool is_unpublished(double v) { return __double_as_longlong(v) == -1ll; }
// Publishing store.  FP64 arithmetic passes NaN payloads through, so an all-ones NaN in the caller's
// variables or parameters could otherwise arrive here and look "absent" for ever: any value with the
// sentinel's bit pattern is replaced by the canonical quiet NaN before it is stored.
__device__ __forceinline__ void st_relaxed(double* p, double v) {
    if (is_unpublished(v)) v = __longlong_as_double(0x7ff8000000000000ll);
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v));
}

// Every polling loop is bounded: after kSpinBatch unsuccessful polls a thread looks at the status word
// and at the wall clock; once a waiter has spun for kSpinLimitNs (or another waiter has already given
// up: status >= kStatusStalled) it raises the status to kStatusStalled and leaves the loop with whatever
// it read.  The kernel then runs to completion on NaN-poisoned data (its own publications still happen,
// so nobody waits for it), and the host turns the status into an error instead of a hung device.
constexpr int kStatusStalled = 3;
constexpr uint32_t kSpinBatch = 1024;
constexpr long long kSpinLimitNs = 4000000000ll;
struct SpinGuard {
    int* status;
    uint32_t polls = 0;
    long long t0 = 0;
    __device__ __forceinline__ explicit SpinGuard(int* st) : status(st) {}
    // call after an unsuccessful poll; true: give up
    __device__ __forceinline__ bool expired() {
        if (++polls < kSpinBatch) return false;
        polls = 0;
        long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        const int stv = *reinterpret_cast<volatile int*>(status);
        if (stv >= kStatusStalled || now - t0 > kSpinLimitNs) {
            atomicMax(status, kStatusStalled);
            return true;
        }
        return false;
    }
};

// Work items of the polling kernels are handed out by an atomic ticket instead of blockIdx: CUDA does not
// promise that CTAs start in index order, but a CTA that holds ticket k knows that the CTAs holding every
// ticket < k have started, and the task lists are ordered so that a task only waits for tasks before it.
// Hence no residency condition and no dependence on dispatch order.
__device__ __forceinline__ uint32_t take_ticket(uint32_t* counter) {
    __shared__ uint32_t s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    return s_ticket;
}

#ifdef FK_CHAIN_PROFILE
__device__ __forceinline__ long long global_ns() {
    long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    return v;
}
#endif
struct TileBuf {
    double A[2][KC][TB];
    double B[2][KC][TB];
};

__device__ __forceinline__ uint32_t tile_row(uint32_t tx, int i) { return tx + 16u * (uint32_t)i; }

template <bool POLL = false>
__device__ __forceinline__ void tile_kloop(double (&acc)[4][4], const double* __restrict__ P, uint32_t f, uint32_t rowA0,
                                           uint32_t nrA, uint32_t rowB0, uint32_t nrB, uint32_t K, uint32_t diag0, TileBuf& buf,
                                           int* status = nullptr) {
    const uint32_t tid = threadIdx.x;
    const uint32_t lk = tid >> 4, lr = (tid & 15) * 4;   // staging: column lk of the chunk, rows lr..lr+3
    const uint32_t tx = tid & 15, c4 = (tid >> 4) * 4;   // compute mapping
    const uint32_t nchunks = (K + KC - 1) / KC;
    double ra[4], rb[4], rdk = 0.0;
    auto gload = [&](uint32_t ch) {
        const uint32_t k = ch * KC + lk;
#pragma unroll
        for (int u = 0; u < 4; u++) ra[u] = rb[u] = 0.0;
        rdk = 0.0;
        if (k < K) {
            const double* col = P + (size_t)k * f;
            if (POLL) {  // operands published by other CTAs of this launch: issued here, verified by gverify() after the
                         // multiplication of the current chunk (a value still missing is re-read there)
                rdk = ld_relaxed(col + diag0 + k);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (lr + u < nrA) ra[u] = ld_relaxed(col + rowA0 + lr + u);
                    if (lr + u < nrB) rb[u] = ld_relaxed(col + rowB0 + lr + u);
                }
            } else {
                rdk = col[diag0 + k];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (lr + u < nrA) ra[u] = col[rowA0 + lr + u];
                    if (lr + u < nrB) rb[u] = col[rowB0 + lr + u];
                }
            }
        }
    };
    auto gverify = [&](uint32_t ch) {
        if (!POLL) return;
        const uint32_t k = ch * KC + lk;
        if (k >= K) return;
        const double* col = P + (size_t)k * f;
        SpinGuard guard(status);
        for (;;) {
            bool missing = is_unpublished(rdk);
#pragma unroll
            for (int u = 0; u < 4; u++) missing = missing || (lr + u < nrA && is_unpublished(ra[u])) || (lr + u < nrB && is_unpublished(rb[u]));
            if (!missing || guard.expired()) break;
            rdk = ld_relaxed(col + diag0 + k);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (lr + u < nrA) ra[u] = ld_relaxed(col + rowA0 + lr + u);
                if (lr + u < nrB) rb[u] = ld_relaxed(col + rowB0 + lr + u);
            }
        }
    };
    auto lds = [&](uint32_t b, int k, double (&a)[4], double (&bb)[4]) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = buf.A[b][k][tile_row(tx, u)];
            bb[u] = buf.B[b][k][c4 + u];
        }
    };
    auto mac = [&](const double (&a)[4], const double (&bb)[4]) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fma(-a[i], bb[j], acc[i][j]);
    };
    if (nchunks) gload(0);
    for (uint32_t ch = 0; ch < nchunks; ch++) {
        const uint32_t b = ch & 1;
        gverify(ch);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            buf.A[b][lk][lr + u] = ra[u] * rdk;
            buf.B[b][lk][lr + u] = rb[u];
        }
        __syncthreads();
        if (ch + 1 < nchunks) gload(ch + 1);
        double a0[4], b0[4], a1[4], b1[4];
        lds(b, 0, a0, b0);
#pragma unroll
        for (int k = 0; k < KC; k += 2) {
            lds(b, k + 1, a1, b1);
            mac(a0, b0);
            if (k + 2 < KC) lds(b, k + 2, a0, b0);
            mac(a1, b1);
        }
    }
    __syncthreads();
}

#ifdef FK_DIAG_PROFILE
#define FK_STAMP(k) do { if (threadIdx.x == 0) ((long long*)D.ubuf)[k] = clock64(); } while (0)
#else
#define FK_STAMP(k) do { } while (0)
#endif
// ---- 8-column register micro-kernels ---------------------------------------------------------------------
// The pivot chain of a 64-column tile is latency bound (measured on B200: DFMA 8.6, LDS ~29,
// rcp+2 Newton 47 cycles, a CTA barrier ~100), so the tile kernels work on micro-panels of 8
// columns: EVERY thread factorises the 8x8 diagonal block redundantly in registers (operands are
// shared-memory broadcasts; static register indices, ~250 instructions, executed 8 times per tile,
// so the code stays instruction-cache resident), then solves its own row against it — no
// communication inside a micro-panel — and the rest of the tile receives a rank-8 update.
constexpr int MB = 8;

// LDLt of the 8x8 block at rows/cols k0.. of tile Cs (row stride ld): on return ll[a][b] (b < a)
// holds the unit-lower factor, inv[k] = 1 / d_k, piv[k] = d_k.  Columns >= kw are treated as identity.
__device__ __forceinline__ void micro_ldl(const double* Cs, uint32_t ld, uint32_t k0, uint32_t kw, double (&ll)[MB][MB],
                                          double (&inv)[MB], double (&piv)[MB]) {
#pragma unroll
    for (int a = 0; a < MB; a++)
#pragma unroll
        for (int b = 0; b <= a; b++)
            ll[a][b] = ((uint32_t)a < kw) ? Cs[(k0 + a) * ld + k0 + b] : (a == b ? 1.0 : 0.0);
#pragma unroll
    for (int k = 0; k < MB; k++) {
        piv[k] = ll[k][k];
        inv[k] = fast_rcp(piv[k]);
        // a_ab -= (a_ak a_bk) / d_k: the products do not wait for the reciprocal, so the chain from one pivot to the
        // next is reciprocal + one FMA
#pragma unroll
        for (int a = k + 1; a < MB; a++)
#pragma unroll
            for (int b = k + 1; b <= a; b++) ll[a][b] = fma(-(ll[a][k] * ll[b][k]), inv[k], ll[a][b]);
#pragma unroll
        for (int a = k + 1; a < MB; a++) ll[a][k] *= inv[k];
    }
}

// Diagonal tile (pivot block at col0, nc <= 64 columns) of supernode s, right-looking (the tile has
// already received the updates of all earlier pivot blocks).  256 threads = 64 rows x 4 column
// quarters.  Per micro-panel: barrier, register LDLt of the 8x8 block + substitution of the own
// row (all four threads of a row redundantly), barrier, rank-8 update of the trailing tile.
constexpr int kDiagThreads = 256;
constexpr int kTsLd = TB + 1;

// LDLt of the nc x nc tile whose lower part sits in Cs (row stride kTsLd).  Writes the unit factor (scaled)
// and D (diagonal) to T (column-major, leading dimension f) and, when PUB, to Tp for the CTAs polling it.
#ifdef FK_CHAIN_PROFILE
// harness only (status is a large buffer there): clock64 stamps of the first micro-panels of CTA 0
#define FK_DSTAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && k0 < 3 * MB) ((long long*)status)[8 + (k0 / MB) * 8 + (k)] = clock64(); } while (0)
#else
#define FK_DSTAMP(k) do { } while (0)
#endif
// Named CTA barriers of the diagonal-tile factorisation (barrier 0 is __syncthreads).
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
constexpr int kYfLd = TB + 2;          // row stride of the unscaled-multiplier tile: even, so an 8-column strip of a row is 16-byte aligned
constexpr uint32_t kDiagBar = 224;     // threads on the named barriers: the pivot warp + six helper warps
constexpr size_t kDiagSmem = (TB * kTsLd + TB * kYfLd) * sizeof(double);

// The pivot chain of the tile is a sequence of dependent steps (per column: reciprocal, one FMA into the next pivot), so
// the tile is NOT spread over the CTA: ONE warp (warp 0) walks the chain, lane l owning rows l and l + 32, with no CTA
// barrier on its path.  Per micro-panel p of 8 columns it (1) applies micro-panel p - 1 to the 8x8 diagonal block only
// (36 entries, one or two per lane), (2) factorises the block redundantly in registers, (3) meets the helpers, solves
// its two rows against the block and leaves the scaled multipliers in Cs, the unscaled ones (the pivot itself on the
// diagonal) in Ys.  Six helper warps apply every finished micro-panel to everything below the next diagonal block
// (thread = row x column quarter, columns in ascending order, so the next micro-panel's columns come first) while the
// pivot warp is busy with (1) and (2) of the next micro-panel.  Warp 4 -- it shares the scheduler with the pivot warp --
// only copies finished micro-panels to global memory (panel storage and, when PUB, the polled copy).  Every entry
// receives its updates in ascending micro-panel order whoever applies them, so the result does not depend on timing.
// Hand-over by named barriers (producer bar.arrive, consumer bar.sync; the ids alternate with the parity of the
// micro-panel): 1, 2 "micro-panel p is in shared memory" (all 256 threads), 3, 4 "the helpers have applied
// micro-panel p" (helpers -> pivot warp, 224 threads).  The upper triangle of Cs / Ys is never read.
template <bool PUB>
__device__ __forceinline__ void diag_tile_factor(double* Cs, double* Ys, uint32_t nc, double* T, double* Tp, uint32_t f, int* status) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t np = (nc + MB - 1) / MB;
    __syncthreads();  // the tile is complete in Cs
    if (warp == 0) {
        for (uint32_t p = 0; p < np; p++) {
            const uint32_t k0 = p * MB, kw = min((uint32_t)MB, nc - k0);
            FK_DSTAMP(0);
            if (p >= 1) {
                // micro-panel p - 1 -> the diagonal block: lane (a, q) takes the entries (a, q) and (a, q + 4) of the block
                // (computed for the whole 8x8 square: the upper triangle is never read)
                const uint32_t a = lane >> 2, q = lane & 3;
                const double* lr = Cs + (k0 + a) * kTsLd + (k0 - MB);
                const double2* y0 = reinterpret_cast<const double2*>(Ys + (k0 + q) * kYfLd + (k0 - MB));
                const double2* y1 = reinterpret_cast<const double2*>(Ys + (k0 + q + 4) * kYfLd + (k0 - MB));
                double* c0 = Cs + (k0 + a) * kTsLd + k0 + q;
                double s00 = c0[0], s01 = 0.0, s10 = c0[4], s11 = 0.0;  // two partial sums per entry
#pragma unroll
                for (int cp = 0; cp < MB; cp += 2) {
                    const double la = lr[cp], lb = lr[cp + 1];
                    const double2 ya = y0[cp >> 1], yb = y1[cp >> 1];
                    s00 = fma(-la, ya.x, s00);
                    s01 = fma(-lb, ya.y, s01);
                    s10 = fma(-la, yb.x, s10);
                    s11 = fma(-lb, yb.y, s11);
                }
                c0[0] = s00 + s01;
                c0[4] = s10 + s11;
                __syncwarp();
            }
            FK_DSTAMP(1);
            double ll[MB][MB], inv[MB], piv[MB];
            micro_ldl(Cs, kTsLd, k0, kw, ll, inv, piv);
            FK_DSTAMP(2);
            {   // columns beyond kw are the identity, so every piv[k] can be looked at
                bool isnan_ = false, bad = false;
#pragma unroll
                for (int k = 0; k < MB; k++) {
                    isnan_ = isnan_ || piv[k] != piv[k];
                    bad = bad || !(piv[k] > 0.0) || piv[k] == INFINITY;
                }
                if (lane == 0 && (isnan_ || bad)) atomicMax(status, isnan_ ? 2 : 1);
            }
            if (p >= 1) bar_sync(3 + ((p - 1) & 1), kDiagBar);  // the helpers have applied micro-panel p - 1
            // (no predicates below: the upper triangle of Cs / Ys, rows that are already finished and columns beyond nc hold
            // values nobody reads, so the warp loads, solves and stores all 2 x 8 entries of its rows)
            double v[2][MB];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int c = 0; c < MB; c++) v[h][c] = Cs[(lane + 32u * h) * kTsLd + k0 + c];
            // own rows against the block: y_c = v_c - sum_{c' < c} y_c' L88[c][c'] (column-oriented: the FMAs of one step are
            // independent); for a row inside the block y_c is its pivot at c == r - k0 and unused beyond
#pragma unroll
            for (int cp = 0; cp + 1 < MB; cp++)
#pragma unroll
                for (int c = cp + 1; c < MB; c++)
#pragma unroll
                    for (int h = 0; h < 2; h++) v[h][c] = fma(-v[h][cp], ll[c][cp], v[h][c]);
            __syncwarp();  // every lane has read its rows (the block's rows among them)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t r = lane + 32u * h;
#pragma unroll
                for (int c = 0; c < MB; c++) {
                    Ys[r * kYfLd + k0 + c] = v[h][c];            // unscaled; the pivot itself on the diagonal
                    Cs[r * kTsLd + k0 + c] = v[h][c] * inv[c];   // scaled multipliers
                }
            }
            __syncwarp();  // (the next diagonal-block update reads what the other lanes have just stored)
            bar_arrive(1 + (p & 1), kDiagThreads);
            FK_DSTAMP(3);
        }
    } else if (warp != 4) {
        const uint32_t hid = (warp < 4 ? warp - 1 : warp - 2) * 32 + lane;                     // 0..191
        const uint32_t i = 2 * MB + hid % (TB - 2 * MB), q = hid / (TB - 2 * MB);               // row 16..63, column quarter
        double* row = Cs + i * kTsLd;
        for (uint32_t p = 0; p < np; p++) {
            const uint32_t k0 = p * MB;
            bar_sync(1 + (p & 1), kDiagThreads);
            if (p + 1 >= np) continue;
            if (i >= k0 + 2 * MB && i < nc) {
                double l[MB];
#pragma unroll
                for (int c = 0; c < MB; c++) l[c] = row[k0 + c];
                // four columns per step, all loads before the first store (the compiler cannot tell that the store to row[j]
                // leaves the next column's operands alone and would otherwise run the columns one after the other)
                for (uint32_t j0 = k0 + MB + q; j0 <= i; j0 += 16) {
                    double v0[4], v1[4];
                    double2 y2[4][MB / 2];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t j = min(j0 + 4u * u, (uint32_t)TB - 1);  // (clamped: columns beyond i are computed, not stored)
                        const double2* yj = reinterpret_cast<const double2*>(Ys + j * kYfLd + k0);
                        v0[u] = row[j];
                        v1[u] = 0.0;  // two partial sums: half the dependent-FMA chain
#pragma unroll
                        for (int c = 0; c < MB / 2; c++) y2[u][c] = yj[c];
                    }
#pragma unroll
                    for (int c = 0; c < MB; c += 2)
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            v0[u] = fma(-l[c], y2[u][c >> 1].x, v0[u]);
                            v1[u] = fma(-l[c + 1], y2[u][c >> 1].y, v1[u]);
                        }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (j0 + 4u * u <= i) row[j0 + 4u * u] = v0[u] + v1[u];
                }
            }
            bar_arrive(3 + (p & 1), kDiagBar);
        }
    } else {
        // copy-out warp: the finished micro-panel (rows k0.., 8 columns) from shared memory to the panel storage
        for (uint32_t p = 0; p < np; p++) {
            const uint32_t k0 = p * MB, kw = min((uint32_t)MB, nc - k0);
            bar_sync(1 + (p & 1), kDiagThreads);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t r = lane + 32u * h;
                if (r >= k0 && r < nc) {
                    const uint32_t cm = min(r - k0, kw - 1);
                    double* tp = T + (size_t)k0 * f + r;
#pragma unroll
                    for (int c = 0; c < MB; c++) {
                        if ((uint32_t)c <= cm) {
                            const double out = (uint32_t)c == r - k0 ? Ys[r * kYfLd + k0 + c] : Cs[r * kTsLd + k0 + c];  // D on the diagonal
                            tp[(size_t)c * f] = out;
                            if (PUB) st_relaxed(Tp + (size_t)(k0 + c) * f + r, out);
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kDiagThreads)
mf_diag_kernel(MfDev D, const uint4* __restrict__ tasks) {
    extern __shared__ __align__(16) double smd_diag[];
    double* Cs = smd_diag;               // [TB][kTsLd] the tile (lower part)
    double* Ys = Cs + TB * kTsLd;        // [TB][kYfLd] unscaled multipliers
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, col0 = t.z, nc = t.w >> 16;
    const uint32_t f = __ldg(D.f + s);
    double* T = D.pan + __ldg(D.pan_off + s) + (size_t)col0 * f + col0;
    const uint32_t tid = threadIdx.x;
    for (uint32_t e = tid; e < nc * TB; e += kDiagThreads) {
        const uint32_t ii = e & 63, j = e >> 6;
        if (ii < nc && j <= ii) Cs[ii * kTsLd + j] = T[(size_t)j * f + ii];
    }
    diag_tile_factor<false>(Cs, Ys, nc, T, nullptr, f, D.status);
}

// Panel tile (rows row0..row0+nr) x (pivot block at col0, nc columns): L = C L_kk^-T D^-1 by blocked
// forward substitution.  256 threads = 64 rows x 4 column quarters, the four threads of a row in
// one warp (lanes 4r..4r+3): per micro-panel every thread solves its row's 8 entries in registers
// (redundantly x4), then the four share the rank-8 update of the rest of the row; __syncwarp only.
constexpr int kColThreads = 256;
constexpr int kYsLd = TB + 1;  // row stride of the substitution tile: odd, so lanes = rows hit distinct banks (with TB + 4, the stride
                               // of the four-threads-per-row mapping this tile used to have, a column access by 32 rows was an 8-way conflict)
constexpr int kLtLd = TB + 2;  // row stride of the TRANSPOSED unit-lower tile: even (16-byte aligned 8-column strips), and the four
                               // rows k, k+1, k+2, k+3 the quarters of a row read together start 4 banks apart
// Blocked forward substitution of the rows in Cs ([TB][kYsLd], zero padded) against the unit-lower tile, held transposed in
// Lt ([TB][kLtLd]: Lt[k][j] = L[j][k] for j > k, zero elsewhere): on return Cs holds Y = C L_kk^-T (unscaled).
// One strip of 8 columns at a time, LEFT-looking: y[r][k0+c] = C[r][k0+c] - sum_{k < k0+c} y[r][k] L[k0+c][k], in two steps
// separated by a CTA barrier (the caller's):
//  (1) col_strip_partial, all 256 threads: warp w takes the rows 32 (w & 1) + lane and the columns k = (w >> 1) mod 4 of the
//      sum over k < k0: per k one value of the lane's row (a conflict-free load) and one 64-byte strip of Lt that the whole
//      warp reads from the SAME address (four broadcast loads) feed eight FMAs; the eight partial sums go to Ps;
//  (2) col_strip_finish, threads 0..63 = rows: the four partial sums are added in a fixed order, the eight columns of the
//      strip are solved in registers, y goes back to Cs.
// What this replaces read the strip of Lt from four addresses per warp (a 16-byte load from four addresses costs four
// passes of the shared-memory pipe, not one) and, before that, nine values per FMA pair of every LATER column in every
// strip (right-looking): the tile took 1.9 us per strip behind a diagonal tile that publishes a strip every microsecond.
constexpr int kPsLd = MB + 1;  // row stride of the partial sums: 32 rows -> 32 different banks
constexpr size_t kPsDoubles = 4 * (size_t)TB * kPsLd;

__device__ __forceinline__ void col_strip_partial(const double* Cs, const double* Lt, double* Ps, uint32_t k0) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = 32u * (warp & 1u) + lane, ks = warp >> 1;
    const double* row = Cs + r * kYsLd;
    double s[MB];
#pragma unroll
    for (int c = 0; c < MB; c++) s[c] = 0.0;
#pragma unroll 2
    for (uint32_t k = ks; k < k0; k += 4) {
        const double yk = row[k];
        const double2* l2 = reinterpret_cast<const double2*>(Lt + k * kLtLd + k0);
#pragma unroll
        for (int c = 0; c < MB; c += 2) {
            const double2 t = l2[c >> 1];
            s[c] = fma(yk, t.x, s[c]);
            s[c + 1] = fma(yk, t.y, s[c + 1]);
        }
    }
    double* ps = Ps + (ks * TB + r) * kPsLd;
#pragma unroll
    for (int c = 0; c < MB; c++) ps[c] = s[c];
}

// threads 0..63 only (r = threadIdx.x); y: the row's eight values of the strip (unscaled)
__device__ __forceinline__ void col_strip_finish(double* Cs, const double* Lt, const double* Ps, uint32_t k0, uint32_t r, double (&y)[MB]) {
    double* row = Cs + r * kYsLd;
#pragma unroll
    for (int c = 0; c < MB; c++) {
        const double a = Ps[(0 * TB + r) * kPsLd + c] + Ps[(1 * TB + r) * kPsLd + c];
        const double b = Ps[(2 * TB + r) * kPsLd + c] + Ps[(3 * TB + r) * kPsLd + c];
        y[c] = row[k0 + c] - (a + b);
    }
#pragma unroll
    for (int cp = 0; cp + 1 < MB; cp++) {  // column-oriented substitution inside the strip: independent FMAs per step
        const double2* l2 = reinterpret_cast<const double2*>(Lt + (k0 + cp) * kLtLd + k0);
        double l[MB];
#pragma unroll
        for (int c = (cp + 1) & ~1; c < MB; c += 2) {
            const double2 t = l2[c >> 1];
            l[c] = t.x;
            l[c + 1] = t.y;
        }
#pragma unroll
        for (int c = cp + 1; c < MB; c++) y[c] = fma(-y[cp], l[c], y[c]);
    }
#pragma unroll
    for (int c = 0; c < MB; c++) row[k0 + c] = y[c];
}

__device__ __forceinline__ void col_tile_solve(double* Cs, const double* Lt, double* Ps, uint32_t nc) {
    for (uint32_t k0 = 0; k0 < nc; k0 += MB) {
        col_strip_partial(Cs, Lt, Ps, k0);
        __syncthreads();
        if (threadIdx.x < TB) {
            double y[MB];
            col_strip_finish(Cs, Lt, Ps, k0, threadIdx.x, y);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kColThreads)
mf_col_kernel(MfDev D, const uint4* __restrict__ tasks) {
    extern __shared__ __align__(16) double smd_col[];
    double* Cs = smd_col;            // [TB][kYsLd] rows of the tile (Y, unscaled)
    double* Ls = Cs + TB * kYsLd;    // [TB][kLtLd] L_kk (unit lower) transposed
    double* Ps = Ls + TB * kLtLd;    // [4][TB][kPsLd] partial sums of a strip
    __shared__ double invd[TB];
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, row0 = t.y, col0 = t.z, nr = t.w & 0xFFFFu, nc = t.w >> 16;
    const uint32_t f = __ldg(D.f + s);
    double* P = D.pan + __ldg(D.pan_off + s);
    const double* Lkk = P + (size_t)col0 * f + col0;
    double* C = P + (size_t)col0 * f + row0;
    const uint32_t tid = threadIdx.x;
    const uint32_t ncp = (nc + MB - 1) & ~(uint32_t)(MB - 1);
    for (uint32_t e = tid; e < ncp * TB; e += kColThreads) {
        const uint32_t ii = e & 63, j = e >> 6;
        Cs[ii * kYsLd + j] = (ii < nr && j < nc) ? C[(size_t)j * f + ii] : 0.0;
        Ls[j * kLtLd + ii] = (ii < nc && j < ii) ? Lkk[(size_t)j * f + ii] : 0.0;  // strictly lower part, transposed, zero padded
    }
    if (tid < nc) invd[tid] = fast_rcp(Lkk[(size_t)tid * f + tid]);
    __syncthreads();
    col_tile_solve(Cs, Ls, Ps, nc);
    __syncthreads();
    for (uint32_t e = tid; e < nc * TB; e += kColThreads) {
        const uint32_t ii = e & 63, j = e >> 6;
        if (ii < nr) C[(size_t)j * f + ii] = Cs[ii * kYsLd + j] * invd[j];
    }
}

// Right-looking update with pivot block kb (columns pc0..pc0+K of the panel, K <= 64) of one
// 64x64 tile to its lower right: target rows row0.., target columns tcol0.. (front coordinates);
// the target lives in the panel when tcol0 < ns and in U_s otherwise.  t.w packs nrA | nrB << 8 |
// K << 16 | (pc0 / 64) << 24.
__global__ void __launch_bounds__(kTileThreads)
mf_rupd_kernel(MfDev D, const uint4* __restrict__ tasks) {
    __shared__ TileBuf buf;
    FK_STAMP(40);
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, row0 = t.y, tcol0 = t.z;
    const uint32_t nrA = (t.w & 0xFFu) + 1, nrB = ((t.w >> 8) & 0xFFu) + 1, K = ((t.w >> 16) & 0xFFu) + 1, pc0 = (t.w >> 24) * TB;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), r = f - ns;
    double* P = D.pan + __ldg(D.pan_off + s);
    double* T;      // target tile base: element (i, j) at T[j * ld + i]
    uint32_t ld;
    if (tcol0 < ns) {
        T = P + (size_t)tcol0 * f + row0;
        ld = f;
    } else {
        T = D.upd + __ldg(D.upd_off + s) + (size_t)(tcol0 - ns) * r + (row0 - ns);
        ld = r;
    }
    const bool diag_tile = row0 == tcol0;
    const uint32_t tid = threadIdx.x, tx = tid & 15, c4 = (tid >> 4) * 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t ri = tile_row(tx, i);
            const bool ok = ri < nrA && c4 + j < nrB && (!diag_tile || ri >= c4 + j);
            acc[i][j] = ok ? T[(size_t)(c4 + j) * ld + ri] : 0.0;
        }
    FK_STAMP(41);
    tile_kloop(acc, P + (size_t)pc0 * f, f, row0, nrA, tcol0, nrB, K, pc0, buf);
    FK_STAMP(42);
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t ri = tile_row(tx, i);
            const bool ok = ri < nrA && c4 + j < nrB && (!diag_tile || ri >= c4 + j);
            if (ok) T[(size_t)(c4 + j) * ld + ri] = acc[i][j];
        }
    FK_STAMP(43);
}

// ---- tile dataflow factorisation of the levels near the root ---------------------------------------------
// On the levels that hold a few wide supernodes the launch-per-step schedule (diag, panel, update per 64
// pivot columns: 16 launches for a 260-column supernode) is a chain of short kernels separated by launch
// gaps.  Here every 64x64 tile of a front is owned by ONE CTA for its whole life (left-looking per tile):
// it loads the assembled tile into registers, subtracts the products of the finished pivot blocks to its
// left as their panel tiles are published (polled in place, operands staged through shared memory by the
// same register-blocked micro-kernel as mf_rupd_kernel), then factorises it (diagonal tile), solves it
// against the published diagonal tile (panel tile) or stores it (update-matrix tile).  Tiles are numbered
// column by column, so a CTA only ever waits for CTAs with a smaller block index: the schedule cannot
// deadlock however many CTAs are resident.  One launch per level instead of 1 + 3 per pivot block.
#ifdef FK_CHAIN_PROFILE
#define FK_FSTAMP(k) do { if (threadIdx.x == 0) ((long long*)D.ubuf)[bid * 4 + (k)] = global_ns(); } while (0)
// harness only: per-strip stamps of the panel tile with ticket 1 (slots behind 1024 tiles)
#define FK_PSTAMP(strip, k) do { if (threadIdx.x == 0 && bid == 1) ((long long*)D.ubuf)[4096 + (strip) * 4 + (k)] = global_ns(); } while (0)
#else
#define FK_FSTAMP(k) do { } while (0)
#define FK_PSTAMP(strip, k) do { } while (0)
#endif
constexpr size_t kFlowSmem = (TB * kYsLd + TB * kLtLd + TB + kPsDoubles) * sizeof(double);
static_assert(kFlowSmem >= sizeof(TileBuf) && kFlowSmem >= kDiagSmem, "the staging buffers and the diagonal-tile arrays alias the solve tiles");

__global__ void __launch_bounds__(kTileThreads)
mf_flow_kernel(MfDev D, const uint4* __restrict__ tasks, double* __restrict__ pub, const uint32_t* __restrict__ asm_ptr,
               const uint4* __restrict__ asm_ent, uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(16) double smf[];
    TileBuf& buf = *reinterpret_cast<TileBuf*>(smf);
    const uint32_t bid = take_ticket(ticket);
    const uint4 t = __ldg(tasks + bid);
    const uint32_t s = t.x, row0 = t.y, tcol0 = t.z;
    const uint32_t nrA = (t.w & 0xFFu) + 1, nrB = ((t.w >> 8) & 0xFFu) + 1;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), r = f - ns;
    const uint64_t po = __ldg(D.pan_off + s);
    double* P = D.pan + po;
    double* Pp = pub + po;
    const bool in_panel = tcol0 < ns;
    const uint32_t nwait = in_panel ? tcol0 / TB : (ns + TB - 1) / TB;
    double* T;      // target tile base: element (i, j) at T[j * ld + i]
    uint32_t ld;
    if (in_panel) {
        T = P + (size_t)tcol0 * f + row0;
        ld = f;
    } else {
        T = D.upd + __ldg(D.upd_off + s) + (size_t)(tcol0 - ns) * r + (row0 - ns);
        ld = r;
    }
    const bool diag_tile = row0 == tcol0;
    const uint32_t tid = threadIdx.x, tx = tid & 15, c4 = (tid >> 4) * 4;
    double acc[4][4];
    FK_FSTAMP(0);
    {
        // Extend-add of the children (replaces mf_asm_kernel on these levels): the tile's assembled values go to
        // shared memory, every child adds the part of its update matrix that lands in this tile (index ranges
        // precomputed on the host; children in ascending order, distinct targets within a child), and the sum is
        // the tile's starting value.  Update-matrix tiles start from zero, so U_s needs no separate clearing.
        double* S = smf;  // [TB][kTsLd]
        const uint32_t e0 = __ldg(asm_ptr + bid), e1 = __ldg(asm_ptr + bid + 1);
        for (uint32_t e = tid; e < TB * TB; e += kTileThreads) {
            const uint32_t ri = e & 63, cj = e >> 6;
            const bool ok = in_panel && ri < nrA && cj < nrB && (!diag_tile || ri >= cj);
            S[ri * kTsLd + cj] = ok ? T[(size_t)cj * ld + ri] : 0.0;
        }
        __syncthreads();
        for (uint32_t q = e0; q < e1; q++) {
            const uint4 A = __ldg(asm_ent + q);
            const uint32_t c = A.x, a0 = A.y & 0xFFFFu, na = A.y >> 16, b0 = A.z & 0xFFFFu, nb = A.z >> 16;
            const uint32_t rc = __ldg(D.f + c) - __ldg(D.ns + c);
            const double* Uc = D.upd + __ldg(D.upd_off + c);
            const uint32_t* relc = D.rel + __ldg(D.rel_off + c);
            for (uint32_t x = tid; x < na * nb; x += kTileThreads) {
                const uint32_t a = a0 + x % na, b = b0 + x / na;
                if (a >= b) S[(__ldg(relc + a) - row0) * kTsLd + (__ldg(relc + b) - tcol0)] += Uc[(size_t)b * rc + a];
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = S[tile_row(tx, i) * kTsLd + c4 + j];
        __syncthreads();  // S aliases the staging buffers of the update loop
    }
    for (uint32_t b = 0; b + 1 < nwait; b++)
        tile_kloop<true>(acc, Pp + (size_t)(TB * b) * f, f, row0, nrA, tcol0, nrB, min((uint32_t)TB, ns - TB * b), TB * b, buf, D.status);
    FK_FSTAMP(1);
    for (uint32_t b = nwait ? nwait - 1 : 0; b < nwait; b++)
        tile_kloop<true>(acc, Pp + (size_t)(TB * b) * f, f, row0, nrA, tcol0, nrB, min((uint32_t)TB, ns - TB * b), TB * b, buf, D.status);
    if (!in_panel) {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t ri = tile_row(tx, i);
                if (ri < nrA && c4 + j < nrB && (!diag_tile || ri >= c4 + j)) T[(size_t)(c4 + j) * ld + ri] = acc[i][j];
            }
        return;
    }
    FK_FSTAMP(2);
    const uint32_t nc = nrB;  // pivot columns of this block
    if (diag_tile) {
        double* Cs = smf;
        double* Ys = smf + TB * kTsLd;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t ri = tile_row(tx, i);
                if (ri < nc && c4 + j <= ri) Cs[ri * kTsLd + c4 + j] = acc[i][j];
            }
        diag_tile_factor<true>(Cs, Ys, nc, T, Pp + (size_t)tcol0 * f + row0, f, D.status);
        FK_FSTAMP(3);
        return;
    }
    double* Cs = smf;                 // [TB][kYsLd]
    double* Ls = Cs + TB * kYsLd;     // [TB][kLtLd] the diagonal tile's unit-lower factor, transposed
    double* invd = Ls + TB * kLtLd;   // [TB]
    double* Ps = invd + TB;           // [4][TB][kPsLd] partial sums of a strip
    for (uint32_t e = tid; e < TB * kYsLd; e += kTileThreads) Cs[e] = 0.0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t ri = tile_row(tx, i);
            if (ri < nrA && c4 + j < nc) Cs[ri * kYsLd + c4 + j] = acc[i][j];
        }
    for (uint32_t e = tid; e < TB * kLtLd; e += kTileThreads) Ls[e] = 0.0;
    {   // follow the diagonal tile micro-panel by micro-panel: its 8-column strip (rows >= k0: strictly lower part + D) is
        // polled as the owner of the diagonal tile publishes it (the loads of strip k0 + 8 are issued before strip k0 is
        // processed and verified afterwards, so the L2 round trip overlaps the substitution), this tile's rows take the
        // step, and the strip's columns of THIS tile -- final from here on -- are stored and published at once: the tiles
        // to the right consume them chunk by chunk while the rest of the substitution is still running
        const double* Lkk = Pp + (size_t)tcol0 * f + tcol0;
        const uint32_t ii = tid & 63, cA = tid >> 6;  // entries (ii, k0 + cA) and (ii, k0 + cA + 4)
        double* Tp = Pp + (size_t)tcol0 * f + row0;
        double v[2];
        auto issue = [&](uint32_t k0) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t j = k0 + cA + 4u * u;
                v[u] = (k0 < nc && ii < nc && j < nc && j <= ii) ? ld_relaxed(Lkk + (size_t)j * f + ii) : 0.0;
            }
        };
        // A strip's columns of THIS tile are final once its y is in Cs: stored and published by all threads (two values each,
        // lanes = consecutive rows), so the tiles to the right consume them chunk by chunk while the substitution goes on.
        auto publish = [&](uint32_t k0) {
            const uint32_t r = tid & 63u;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t c = k0 + 2u * (tid >> 6) + (uint32_t)u;
                if (r < nrA && c < nc) {
                    const double out = Cs[r * kYsLd + c] * invd[c];
                    T[(size_t)c * f + r] = out;
                    st_relaxed(Tp + (size_t)c * f + r, out);
                }
            }
        };
        issue(0);
        for (uint32_t k0 = 0, st = 0; k0 < nc; k0 += MB, st++) {
            (void)st;
            SpinGuard guard(D.status);
            FK_PSTAMP(st, 0);
            while ((is_unpublished(v[0]) || is_unpublished(v[1])) && !guard.expired()) issue(k0);
            FK_PSTAMP(st, 1);
            {   // every strip has its own columns of Ls and entries of invd: nothing a reader of an earlier strip still needs
                const uint32_t j0 = k0 + cA, j1 = k0 + cA + 4;
                if (j0 < ii) Ls[j0 * kLtLd + ii] = v[0];
                else if (j0 == ii && ii < nc) invd[ii] = fast_rcp(v[0]);
                if (j1 < ii) Ls[j1 * kLtLd + ii] = v[1];
                else if (j1 == ii && ii < nc) invd[ii] = fast_rcp(v[1]);
            }
            __syncthreads();  // the strip of Ls (and, the first time, the zero fill; later, the previous strip's y) is in place
            issue(k0 + MB);   // (after the barrier: a barrier waits for the loads issued before it)
            if (k0) publish(k0 - MB);  // the previous strip's columns of this tile, by all 256 threads (its y is behind the barrier)
            FK_PSTAMP(st, 2);
            col_strip_partial(Cs, Ls, Ps, k0);
            __syncthreads();
            if (tid < TB) {
                double y[MB];
                col_strip_finish(Cs, Ls, Ps, k0, tid, y);
                FK_PSTAMP(st, 3);
            }
        }
        __syncthreads();
        publish(((nc - 1) / MB) * MB);
    }
    FK_FSTAMP(3);
}

// ---- triangular solves ------------------------------------------------------------------------------------
// Forward: y = L^-1 b with per-supernode update vectors (multifrontal style: a parent gathers its
// children's vectors through the relative indices, in child order, so no atomics).  Backward:
// z = L^-T D^-1 y, gathering the ancestors' solution through the front's row list.
// G = threads cooperating on one supernode (32: a warp, 256: a CTA).
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp();
    else __syncthreads();
}

// Pivot-block triangular solves of a wide supernode (ns >= 64) by one CTA of 256 threads, in
// micro-panels of 8 columns: every thread solves the 8x8 block redundantly in registers, then
// updates the rows it owns.  The panel entries of the NEXT micro-panel are fetched into registers
// while the current one is processed (the loads do not depend on the solution), so the chain per
// micro-panel is one barrier plus ~30 dependent FMAs instead of a DRAM round trip.
constexpr int kPR = 4;  // prefetched rows per thread (covers 1024 rows below a micro-panel)

// L y = t (unit lower, ns x ns, column-major with leading dimension f).  t is consumed; y receives
// the solution.  l88[2][64] is scratch.
__device__ __forceinline__ void pivot_forward_256(const double* __restrict__ P, uint32_t f, uint32_t ns, double* t, double* y,
                                                  double* l88) {
    const uint32_t tid = threadIdx.x;
    const uint32_t pc = tid & 7, pcp = (tid >> 3) & 7;  // threads < 64 stage L88[pc][pcp]
    double nb88 = 0.0, nl[kPR][MB];
    auto prefetch = [&](uint32_t k0) {
        nb88 = (tid < 64 && pcp < pc && k0 + pc < ns) ? P[(size_t)(k0 + pcp) * f + k0 + pc] : 0.0;
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            const uint32_t i = k0 + MB + tid + 256 * r;
#pragma unroll
            for (int c = 0; c < MB; c++) nl[r][c] = (i < ns && k0 + c < ns) ? P[(size_t)(k0 + c) * f + i] : 0.0;
        }
    };
    prefetch(0);
    for (uint32_t k0 = 0; k0 < ns; k0 += MB) {
        double* lb = l88 + ((k0 >> 3) & 1) * 64;
        double cl[kPR][MB];
#pragma unroll
        for (int r = 0; r < kPR; r++)
#pragma unroll
            for (int c = 0; c < MB; c++) cl[r][c] = nl[r][c];
        if (tid < 64) lb[pc * 8 + pcp] = nb88;
        if (k0 + MB < ns) prefetch(k0 + MB);
        __syncthreads();
        double yy[MB];
#pragma unroll
        for (int c = 0; c < MB; c++) yy[c] = k0 + c < ns ? t[k0 + c] : 0.0;
#pragma unroll
        for (int cp = 0; cp + 1 < MB; cp++)
#pragma unroll
            for (int c = cp + 1; c < MB; c++) yy[c] = fma(-lb[c * 8 + cp], yy[cp], yy[c]);
        if (tid < MB && k0 + tid < ns) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < MB; c++) v = tid == (uint32_t)c ? yy[c] : v;
            y[k0 + tid] = v;
        }
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            const uint32_t i = k0 + MB + tid + 256 * r;
            if (i < ns) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < MB; c++) acc = fma(cl[r][c], yy[c], acc);
                t[i] -= acc;
            }
        }
        for (uint32_t i = k0 + MB + tid + 256 * kPR; i < ns; i += 256) {  // very wide supernodes: direct loads
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < MB; c++)
                if (k0 + c < ns) acc = fma(P[(size_t)(k0 + c) * f + i], yy[c], acc);
            t[i] -= acc;
        }
    }
    __syncthreads();
}

// L^T z = t (t already holds D^-1 y minus the contribution of the rows below the pivot block).
// red[8][8] and l88[2][64] are scratch; z receives the solution.
__device__ __forceinline__ void pivot_backward_256(const double* __restrict__ P, uint32_t f, uint32_t ns, const double* t, double* z,
                                                   double* l88, double* red) {
    const uint32_t tid = threadIdx.x, wp = tid >> 5, ln = tid & 31;
    const uint32_t pc = tid & 7, pcp = (tid >> 3) & 7;
    double nb88 = 0.0, nl[kPR][MB];
    auto prefetch = [&](uint32_t k0) {
        nb88 = (tid < 64 && pcp < pc && k0 + pc < ns) ? P[(size_t)(k0 + pcp) * f + k0 + pc] : 0.0;
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            const uint32_t i = k0 + MB + tid + 256 * r;
#pragma unroll
            for (int c = 0; c < MB; c++) nl[r][c] = (i < ns && k0 + c < ns) ? P[(size_t)(k0 + c) * f + i] : 0.0;
        }
    };
    const uint32_t last = ((ns - 1) / MB) * MB;
    prefetch(last);
    for (uint32_t kk = 0; kk <= last; kk += MB) {
        const uint32_t k0 = last - kk;
        double* lb = l88 + ((k0 >> 3) & 1) * 64;
        double part[MB];
#pragma unroll
        for (int c = 0; c < MB; c++) part[c] = 0.0;
        // the z of the rows below this micro-panel are final (written before the previous barrier)
#pragma unroll
        for (int r = 0; r < kPR; r++) {
            const uint32_t i = k0 + MB + tid + 256 * r;
            if (i < ns) {
                const double zi = z[i];
#pragma unroll
                for (int c = 0; c < MB; c++) part[c] = fma(nl[r][c], zi, part[c]);
            }
        }
        for (uint32_t i = k0 + MB + tid + 256 * kPR; i < ns; i += 256) {
            const double zi = z[i];
#pragma unroll
            for (int c = 0; c < MB; c++)
                if (k0 + c < ns) part[c] = fma(P[(size_t)(k0 + c) * f + i], zi, part[c]);
        }
        if (tid < 64) lb[pc * 8 + pcp] = nb88;
        if (k0 >= MB) prefetch(k0 - MB);
#pragma unroll
        for (int c = 0; c < MB; c++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part[c] += __shfl_xor_sync(0xFFFFFFFFu, part[c], o);
        }
        if (ln == 0) {
#pragma unroll
            for (int c = 0; c < MB; c++) red[wp * 8 + c] = part[c];
        }
        __syncthreads();
        double zz[MB];
#pragma unroll
        for (int c = MB - 1; c >= 0; c--) {
            double v = 0.0;
            if (k0 + c < ns) {
                double sum = 0.0;
#pragma unroll
                for (int w8 = 0; w8 < 8; w8++) sum += red[w8 * 8 + c];
                v = t[k0 + c] - sum;
            }
#pragma unroll
            for (int cq = c + 1; cq < MB; cq++) v = fma(-lb[cq * 8 + c], zz[cq], v);
            zz[c] = v;
        }
        if (tid < MB && k0 + tid < ns) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < MB; c++) v = tid == (uint32_t)c ? zz[c] : v;
            z[k0 + tid] = v;
        }
        __syncthreads();
    }
}

template <int G, bool WIDE = false>
__device__ __forceinline__ void forward_supernode(const MfDev& D, uint32_t s, double* __restrict__ w, double* t, double* tri,
                                                  uint32_t gt, bool split = false, double* wide = nullptr) {
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s);
    const double* P = D.pan + __ldg(D.pan_off + s);
    for (uint32_t i = gt; i < f; i += G) t[i] = i < ns ? w[c0 + i] : 0.0;
    group_sync<G>();
    for (uint32_t cq = __ldg(D.child_ptr + s); cq < __ldg(D.child_ptr + s + 1); cq++) {
        const uint32_t c = __ldg(D.child + cq);
        const uint32_t rc = __ldg(D.f + c) - __ldg(D.ns + c), ro = __ldg(D.rel_off + c);
        for (uint32_t a = gt; a < rc; a += G) t[__ldg(D.rel + ro + a)] += D.ubuf[ro + a];
        group_sync<G>();
    }
    const uint32_t rlim = split ? ns : f;  // split: the rows below the pivot block are updated by mf_fwd_upd_kernel
    if (WIDE && G == 256 && ns >= 64) {  // wide pivot block: micro-panel solve with register prefetch
        double* y2 = wide;
        pivot_forward_256(P, f, ns, t, y2, tri);
        for (uint32_t i = gt; i < ns; i += G) t[i] = y2[i];
        group_sync<G>();
        for (uint32_t i = ns + gt; i < rlim; i += G) {
            double acc = 0.0;
#pragma unroll 8
            for (uint32_t k = 0; k < ns; k++) acc = fma(P[(size_t)k * f + i], t[k], acc);
            t[i] -= acc;
        }
        group_sync<G>();
    } else
    for (uint32_t k0 = 0; k0 < ns; k0 += 32) {
        const uint32_t nb = min(32u, ns - k0);
        // stage the nb x nb unit-lower triangle: tri[i][k]
        for (uint32_t e = gt; e < nb * 32; e += G) {
            const uint32_t i = e & 31, k = e >> 5;
            if (i < nb && i > k) tri[i * 33 + k] = P[(size_t)(k0 + k) * f + k0 + i];
        }
        group_sync<G>();
        if (gt < 32) {
            double ti = gt < nb ? t[k0 + gt] : 0.0;
            for (uint32_t k = 0; k + 1 < nb; k++) {
                const double yk = __shfl_sync(0xFFFFFFFFu, ti, k);
                if (gt > k && gt < nb) ti = fma(-tri[gt * 33 + k], yk, ti);
            }
            if (gt < nb) t[k0 + gt] = ti;
        }
        group_sync<G>();
        // rows below the block: t[i] -= L[i, block] . y[block]
        for (uint32_t i = k0 + nb + gt; i < rlim; i += G) {
            double acc = 0.0;
#pragma unroll 8
            for (uint32_t k = 0; k < nb; k++) acc = fma(P[(size_t)(k0 + k) * f + i], t[k0 + k], acc);
            t[i] -= acc;
        }
        group_sync<G>();
    }
    const uint32_t ro = __ldg(D.rel_off + s);
    for (uint32_t i = gt; i < f; i += G) {
        if (i < ns) w[c0 + i] = t[i];
        else D.ubuf[ro + i - ns] = t[i];
    }
    group_sync<G>();
}

template <int G, bool WIDE = false>
__device__ __forceinline__ void backward_supernode(const MfDev& D, uint32_t s, double* __restrict__ w, double* __restrict__ delta,
                                                   const int32_t* __restrict__ perm, double* t, double* tri, uint32_t gt,
                                                   bool split = false, const double* __restrict__ tmp = nullptr, double* wide = nullptr) {
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s);
    const double* P = D.pan + __ldg(D.pan_off + s);
    const uint32_t* rows = D.rows + __ldg(D.rows_off + s);
    // t[i >= ns] = z of the ancestors; t[i < ns] = y_i / d_i  (split: the ancestors' part was
    // already reduced into tmp by mf_bwd_dot_kernel)
    const uint32_t rlim = split ? ns : f;
    for (uint32_t i = gt; i < rlim; i += G) {
        if (i < ns) t[i] = w[c0 + i] / P[(size_t)i * f + i] - (split ? tmp[c0 + i] : 0.0);
        else t[i] = w[__ldg(rows + i)];
    }
    group_sync<G>();
    const uint32_t nblk = (ns + 31) / 32;
    if (WIDE && G == 256 && ns >= 64) {  // wide pivot block
        const uint32_t wp = gt >> 5, ln = gt & 31;
        for (uint32_t k = wp; k < ns && rlim > ns; k += G / 32) {  // rows below the pivot block (not split)
            const double* col = P + (size_t)k * f;
            double acc = 0.0;
            for (uint32_t i = ns + ln; i < rlim; i += 32) acc = fma(col[i], t[i], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
            if (ln == 0) t[k] -= acc;
        }
        group_sync<G>();
        double* z2 = wide;
        pivot_backward_256(P, f, ns, t, z2, tri, tri + 128);
        for (uint32_t i = gt; i < ns; i += G) t[i] = z2[i];
        group_sync<G>();
    } else
    for (uint32_t bi = nblk; bi-- > 0;) {
        const uint32_t k0 = bi * 32, nb = min(32u, ns - k0);
        // column dots with everything below the block: one warp per column
        {
            const uint32_t wp = gt >> 5, ln = gt & 31;
            for (uint32_t k = wp; k < nb; k += G / 32) {
                const double* col = P + (size_t)(k0 + k) * f;
                double acc = 0.0;
                for (uint32_t i = k0 + nb + ln; i < rlim; i += 32) acc = fma(col[i], t[i], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
                if (ln == 0) t[k0 + k] -= acc;
            }
        }
        for (uint32_t e = gt; e < nb * 32; e += G) {
            const uint32_t i = e & 31, k = e >> 5;
            if (i < nb && i > k) tri[i * 33 + k] = P[(size_t)(k0 + k) * f + k0 + i];
        }
        group_sync<G>();
        if (gt < 32) {
            double zi = gt < nb ? t[k0 + gt] : 0.0;
            for (uint32_t k = nb; k-- > 1;) {
                const double zk = __shfl_sync(0xFFFFFFFFu, zi, k);
                if (gt < k) zi = fma(-tri[k * 33 + gt], zk, zi);
            }
            if (gt < nb) t[k0 + gt] = zi;
        }
        group_sync<G>();
    }
    for (uint32_t i = gt; i < ns; i += G) {
        const double z = t[i];
        w[c0 + i] = z;
        delta[__ldg(perm + c0 + i)] = z;
    }
    group_sync<G>();
}

// ---- chained solves of the wide levels near the root -------------------------------------------------
// Near the root a level holds a handful of supernodes with hundreds of pivot columns each; one CTA per
// supernode (mf_big_solve_kernel) then walks a long dependent chain alone.  Here a front is cut into
// 64-row chunks, one CTA each, all resident at once (the host only takes this path when the level
// needs fewer CTAs than the device has SMs).  Forward: the CTA of pivot chunk c applies the updates of
// the pivot blocks b < c as their solutions are published (polled in place, see below; the panel tile
// of the next block is prefetched into registers while waiting), multiplies with the
// precomputed inverse of its own 64x64 unit triangle (mf_chain_inv_kernel, once per factorisation)
// and publishes y_c; the CTAs of the rows below the pivot block consume all blocks and leave the
// update vector.  Backward mirrors it from the last pivot block up.  A thread owns a quarter of one
// row (lanes 4r..4r+3), so every reduction is two shuffles in a fixed order and the loop over the
// awaited blocks has no CTA barrier; one link of the chain costs about a microsecond.
#ifdef FK_CHAIN_PROFILE
#define FK_CSTAMP(k) do { if (threadIdx.x == 0) ((long long*)D.upd)[blockIdx.x * 8 + (k)] = global_ns(); } while (0)
#else
#define FK_CSTAMP(k) do { } while (0)
#endif
// the 16 values q, q+4, ... of a published 64-vector (entries >= count read as zero)
__device__ __forceinline__ void poll16(const double* base, uint32_t q, uint32_t count, double (&v)[16], int* status) {
    bool missing;
    SpinGuard guard(status);
    do {
        missing = false;
#pragma unroll
        for (int u = 0; u < 16; u++) v[u] = q + 4u * u < count ? ld_relaxed(base + q + 4 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < 16; u++) missing = missing || is_unpublished(v[u]);
    } while (missing && !guard.expired());
}
__device__ __forceinline__ double quad_sum(const double (&a4)[4]) {
    double v = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    return v;
}

// X = L_kk^-1 of one 64x64 unit-lower pivot tile (row-major, zero / identity padded), one thread per column.
constexpr size_t kChainInvSmem = 2 * (size_t)TB * kTsLd * sizeof(double);
__global__ void __launch_bounds__(TB)
mf_chain_inv_kernel(MfDev D, const uint4* __restrict__ tasks) {
    extern __shared__ double smi[];
    double* Ls = smi;
    double* Xs = smi + TB * kTsLd;
    const uint4 tk = __ldg(tasks + blockIdx.x);
    const uint32_t s = tk.x, col0 = tk.y, nc = tk.z;
    const uint32_t f = __ldg(D.f + s);
    const double* Lkk = D.pan + __ldg(D.pan_off + s) + (size_t)col0 * f + col0;
    const uint32_t j = threadIdx.x;
    for (uint32_t e = j; e < TB * TB; e += TB) {
        const uint32_t ii = e & 63, k = e >> 6;
        Ls[ii * kTsLd + k] = (ii < nc && k < ii) ? Lkk[(size_t)k * f + ii] : 0.0;
    }
    for (uint32_t i = 0; i < TB; i++) Xs[i * kTsLd + j] = i == j ? 1.0 : 0.0;
    __syncthreads();
    for (uint32_t i = j + 1; i < TB; i++) {
        double a0 = 0.0, a1 = 0.0;
        uint32_t k = j;
        for (; k + 1 < i; k += 2) {
            a0 = fma(Ls[i * kTsLd + k], Xs[k * kTsLd + j], a0);
            a1 = fma(Ls[i * kTsLd + k + 1], Xs[(k + 1) * kTsLd + j], a1);
        }
        if (k < i) a0 = fma(Ls[i * kTsLd + k], Xs[k * kTsLd + j], a0);
        Xs[i * kTsLd + j] = -(a0 + a1);
    }
    __syncthreads();
    double* X = D.winv + (size_t)(__ldg(D.winv_blk + s) + col0 / TB) * (TB * TB);
    for (uint32_t e = j; e < TB * TB; e += TB) X[e] = Xs[(e >> 6) * kTsLd + (e & 63)];
}

__global__ void __launch_bounds__(256)
mf_chain_fwd_kernel(MfDev D, const uint4* __restrict__ tasks, double* __restrict__ w, double* __restrict__ pub,
                    uint32_t* __restrict__ ticket) {
    __shared__ double t[TB];
    const uint4 tk = __ldg(tasks + take_ticket(ticket));
    const uint32_t s = tk.x, row0 = tk.y, nr = tk.z;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s), blk0 = __ldg(D.winv_blk + s);
    const double* P = D.pan + __ldg(D.pan_off + s);
    const uint32_t B = (ns + TB - 1) / TB;
    const bool pivot = row0 < ns;
    const uint32_t nwait = pivot ? row0 / TB : B;
    const uint32_t tid = threadIdx.x, row = tid >> 2, q = tid & 3;
    double nl[16], li[16];
    auto prefetch = [&](uint32_t b) {
        const uint32_t nc = min((uint32_t)TB, ns - TB * b);
        const double* base = P + (size_t)(TB * b) * f + row0 + row;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint32_t k = q + 4u * u;
            nl[u] = (row < nr && k < nc) ? __ldg(base + (size_t)k * f) : 0.0;
        }
    };
    FK_CSTAMP(0);
    if (nwait) prefetch(0);
    if (pivot) {
        const double* X = D.winv + (size_t)(blk0 + row0 / TB) * (TB * TB) + row * TB + q;
#pragma unroll
        for (int u = 0; u < 16; u++) li[u] = __ldg(X + 4 * u);
    }
    // the chunk's part of the front vector: right-hand side (pivot rows) + the children's update vectors
    if (tid < TB) t[tid] = (pivot && tid < nr) ? w[c0 + row0 + tid] : 0.0;
    __syncthreads();
    for (uint32_t cq = __ldg(D.child_ptr + s); cq < __ldg(D.child_ptr + s + 1); cq++) {
        const uint32_t c = __ldg(D.child + cq);
        const uint32_t rc = __ldg(D.f + c) - __ldg(D.ns + c), ro = __ldg(D.rel_off + c);
        // the child's rows inside this chunk (distinct targets; no dependent searches: every thread
        // scans a strided part of the child's list)
        for (uint32_t a = tid; a < rc; a += 256) {
            const uint32_t r = __ldg(D.rel + ro + a) - row0;
            if (r < nr) t[r] += D.ubuf[ro + a];
        }
        __syncthreads();
    }
    double tr = t[row];
    for (uint32_t b = 0; b < nwait; b++) {
        double cl[16];
#pragma unroll
        for (int u = 0; u < 16; u++) cl[u] = nl[u];
        if (b + 1 < nwait) prefetch(b + 1);
        if (b + 1 == nwait) FK_CSTAMP(1);
        double yv[16];
        poll16(pub + c0 + TB * b, q, min((uint32_t)TB, ns - TB * b), yv, D.status);
        if (b + 1 == nwait) FK_CSTAMP(2);
        double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 16; u++) a4[u & 3] = fma(cl[u], yv[u], a4[u & 3]);
        tr -= quad_sum(a4);
    }
    FK_CSTAMP(3);
    if (pivot) {
        __syncthreads();  // (the gather phase's reads of t are long done; this orders the writes below)
        if (q == 0) t[row] = tr;
        __syncthreads();
        double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 16; u++) a4[u & 3] = fma(li[u], t[q + 4 * u], a4[u & 3]);
        const double yr = quad_sum(a4);
        FK_CSTAMP(4);
        if (q == 0 && row < nr) {
            st_relaxed(pub + c0 + row0 + row, yr);
            w[c0 + row0 + row] = yr;
        }
        FK_CSTAMP(5);
    } else if (q == 0 && row < nr) {
        D.ubuf[__ldg(D.rel_off + s) + row0 - ns + row] = tr;
    }
}

// tasks: pivot chunks only, the last chunk of a supernode first.  tmp[c0 + j] holds the dot products of
// column j with the ancestors' solution (mf_bwd_dot_kernel) when the front has rows below the pivot block.
__global__ void __launch_bounds__(256)
mf_chain_bwd_kernel(MfDev D, const uint4* __restrict__ tasks, double* __restrict__ w, double* __restrict__ delta,
                    const int32_t* __restrict__ perm, const double* __restrict__ tmp, double* __restrict__ pub,
                    uint32_t* __restrict__ ticket) {
    __shared__ double t[TB];
    const uint4 tk = __ldg(tasks + take_ticket(ticket));
    const uint32_t s = tk.x, col0 = tk.y, nc = tk.z;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s), blk0 = __ldg(D.winv_blk + s);
    const double* P = D.pan + __ldg(D.pan_off + s);
    const uint32_t B = (ns + TB - 1) / TB, c = col0 / TB;
    const uint32_t tid = threadIdx.x, col = tid >> 2, q = tid & 3;
    double nl[16], li[16];
    auto prefetch = [&](uint32_t b) {  // L(rows of block b, my column), rows q, q+4, ...
        const uint32_t nrb = min((uint32_t)TB, ns - TB * b);
        const double* base = P + (size_t)(col0 + col) * f + TB * b + q;
#pragma unroll
        for (int u = 0; u < 16; u++) nl[u] = (col < nc && q + 4u * u < nrb) ? __ldg(base + 4 * u) : 0.0;
    };
    if (c + 1 < B) prefetch(B - 1);
    {   // column `col` of L_cc^-1 (= row of its transpose), rows q, q+4, ...
        const double* X = D.winv + (size_t)(blk0 + c) * (TB * TB) + (size_t)q * TB + col;
#pragma unroll
        for (int u = 0; u < 16; u++) li[u] = __ldg(X + 4 * u * TB);
    }
    double tr = 0.0;
    if (col < nc) {
        tr = w[c0 + col0 + col] / __ldg(P + (size_t)(col0 + col) * f + col0 + col);
        if (f > ns) tr -= tmp[c0 + col0 + col];
    }
    for (uint32_t b = B - 1; b > c; b--) {
        double cl[16];
#pragma unroll
        for (int u = 0; u < 16; u++) cl[u] = nl[u];
        if (b - 1 > c) prefetch(b - 1);
        double zv[16];
        poll16(pub + c0 + TB * b, q, min((uint32_t)TB, ns - TB * b), zv, D.status);
        double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 16; u++) a4[u & 3] = fma(cl[u], zv[u], a4[u & 3]);
        tr -= quad_sum(a4);
    }
    if (q == 0) t[col] = tr;
    __syncthreads();
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < 16; u++) a4[u & 3] = fma(li[u], t[q + 4 * u], a4[u & 3]);
    const double zz = quad_sum(a4);
    if (q == 0 && col < nc) {
        st_relaxed(pub + c0 + col0 + col, zz);
        w[c0 + col0 + col] = zz;
        delta[__ldg(perm + c0 + col0 + col)] = zz;
    }
}

template <bool FORWARD>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
mf_small_solve_kernel(MfDev D, const uint32_t* __restrict__ sub_ptr, const uint32_t* __restrict__ sub_list, uint32_t nsub,
                      double* __restrict__ w, double* __restrict__ delta, const int32_t* __restrict__ perm) {
    extern __shared__ double sms[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* tri = sms + warp * (32 * 33 + kSmallFront);
    double* t = tri + 32 * 33;
    const uint32_t sub = blockIdx.x * kWarpsPerCta + warp;
    if (sub >= nsub) return;
    const uint32_t q0 = __ldg(sub_ptr + sub), q1 = __ldg(sub_ptr + sub + 1);
    if (FORWARD) {
        for (uint32_t q = q0; q < q1; q++) forward_supernode<32>(D, __ldg(sub_list + q), w, t, tri, lane);
    } else {
        for (uint32_t q = q1; q-- > q0;) backward_supernode<32>(D, __ldg(sub_list + q), w, delta, perm, t, tri, lane);
    }
}

// Supernodes whose panel below the pivot block is large are "split": the pivot-block part runs in
// mf_big_solve_kernel (one CTA), the (f - ns) x ns rectangle in the two kernels below on many CTAs.
constexpr uint32_t kSplitEntries = 16384;
__device__ __forceinline__ bool is_split(uint32_t f, uint32_t ns) { return (f - ns) * ns > kSplitEntries; }

// G threads per supernode: 256 in general, 64 on levels whose fronts all have order <= 64 (the many
// small fronts near the leaves: four times as many supernodes per SM).
template <bool FORWARD, bool WIDE, int G = 256>
__global__ void __launch_bounds__(G)
mf_big_solve_kernel(MfDev D, const uint32_t* __restrict__ list, double* __restrict__ w, double* __restrict__ delta,
                    const int32_t* __restrict__ perm, const double* __restrict__ tmp, uint32_t max_front) {
    extern __shared__ double smd[];
    double* tri = smd;            // [32*33] (wide pivot blocks: L88 staging + reduction scratch)
    double* t = smd + 32 * 33;    // [max front]
    double* wide = t + max_front; // [max front] second vector of the wide pivot-block solves
    const uint32_t s = __ldg(list + blockIdx.x);
    const bool split = is_split(__ldg(D.f + s), __ldg(D.ns + s));
    if (FORWARD) forward_supernode<G, WIDE>(D, s, w, t, tri, threadIdx.x, split, wide);
    else backward_supernode<G, WIDE>(D, s, w, delta, perm, t, tri, threadIdx.x, split, tmp, wide);
}

// Forward, split supernodes: u_s[r0 .. r0+64) -= L21[rows, :] y_s.  256 threads = 64 rows x 4
// interleaved column quarters; the quarters are summed in fixed order.
__global__ void __launch_bounds__(256)
mf_fwd_upd_kernel(MfDev D, const uint4* __restrict__ tasks, const double* __restrict__ w) {
    extern __shared__ double smd[];
    __shared__ double part[4][64];
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, r0 = t.y, nrows = t.z;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s);
    const double* P = D.pan + __ldg(D.pan_off + s);
    double* ys = smd;
    for (uint32_t k = threadIdx.x; k < ns; k += 256) ys[k] = w[c0 + k];
    __syncthreads();
    const uint32_t row = threadIdx.x & 63, q = threadIdx.x >> 6;
    double acc = 0.0;
    if (row < nrows) {
        const double* base = P + ns + r0 + row;
#pragma unroll 8
        for (uint32_t k = q; k < ns; k += 4) acc = fma(base[(size_t)k * f], ys[k], acc);
    }
    part[q][row] = acc;
    __syncthreads();
    if (q == 0 && row < nrows) {
        double* u = D.ubuf + __ldg(D.rel_off + s) + r0 + row;
        *u -= ((part[0][row] + part[1][row]) + part[2][row]) + part[3][row];
    }
}

// Backward, split supernodes: tmp[c0 + j] = sum_{i >= ns} L[i][j] z[rows[i]] for 32 pivot columns per
// CTA, one warp per column at a time.
__global__ void __launch_bounds__(256)
mf_bwd_dot_kernel(MfDev D, const uint4* __restrict__ tasks, const double* __restrict__ w, double* __restrict__ tmp) {
    extern __shared__ double smd[];
    const uint4 t = __ldg(tasks + blockIdx.x);
    const uint32_t s = t.x, j0 = t.y, ncols = t.z;
    const uint32_t f = __ldg(D.f + s), ns = __ldg(D.ns + s), c0 = __ldg(D.c0 + s), r = f - ns;
    const double* P = D.pan + __ldg(D.pan_off + s);
    const uint32_t* rows = D.rows + __ldg(D.rows_off + s) + ns;
    double* zt = smd;
    for (uint32_t a = threadIdx.x; a < r; a += 256) zt[a] = w[__ldg(rows + a)];
    __syncthreads();
    const uint32_t wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    for (uint32_t jj = wp; jj < ncols; jj += 8) {
        const double* col = P + (size_t)(j0 + jj) * f + ns;
        double acc = 0.0;
#pragma unroll 4
        for (uint32_t a = ln; a < r; a += 32) acc = fma(col[a], zt[a], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (ln == 0) tmp[c0 + j0 + jj] = acc;
    }
}

template <class T>
cudaError_t upload_vec(const std::vector<T>& v, const T** out, std::vector<void*>& owned) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    owned.push_back(d);
    if (!v.empty()) e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T*)d;
    return e;
}

template <class T>
cudaError_t alloc_vec(T** out, size_t count, std::vector<void*>& owned) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, count) * sizeof(T));
    if (e != cudaSuccess) return e;
    owned.push_back(d);
    *out = (T*)d;
    return e;
}

constexpr size_t kColSmem = (TB * kYsLd + TB * kLtLd + kPsDoubles) * sizeof(double);
constexpr size_t kSmallFactorSmem = (size_t)kWarpsPerCta * kSmallFront * kSmallLd * sizeof(double);
constexpr size_t kSmallSolveSmem = (size_t)kWarpsPerCta * (32 * 33 + kSmallFront) * sizeof(double);

}  // namespace

Multifrontal::~Multifrontal() {
    if (factor_graph_) cudaGraphExecDestroy(factor_graph_);
    if (solve_graph_) cudaGraphExecDestroy(solve_graph_);
    for (void* p : owned_) cudaFree(p);
}

#define MF_CU(call)                      \
    do {                                 \
        cudaError_t e_ = (call);         \
        if (e_ != cudaSuccess) {         \
            if (err) *err = #call;       \
            return e_;                   \
        }                                \
    } while (0)

// Host-only part of init(): supernodes, tree, relative indices, task lists (no CUDA calls; probed by
// fk_topology_supernodal for the CPU tests).
cudaError_t Multifrontal::build_symbolic(const Topology& t, std::string* err) {
    const uint32_t n = t.n_free;
    n_ = n;
    sym = MfSymbolic();
    static const bool sym_timing = std::getenv("FK_SYM_TIMING") != nullptr;
    auto sym_t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!sym_timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mf symbolic] %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - sym_t0).count());
        sym_t0 = now;
    };
    std::vector<uint32_t>&c0 = sym.c0, &ns = sym.ns, &f = sym.f, &rows_off = sym.rows_off, &rows = sym.rows, &rel_off = sym.rel_off,
                         &rel = sym.rel, &child_ptr = sym.child_ptr, &child = sym.child, &winv_blk = sym.winv_blk, &sub_ptr = sym.sub_ptr,
                         &sub_list = sym.sub_list, &level_list = sym.level_list, &level = sym.level, &tasks = sym.tasks;
    std::vector<uint64_t>&pan_off = sym.pan_off, &upd_off = sym.upd_off;
    std::vector<int32_t>& sparent = sym.sparent;
    std::vector<uint8_t>& big = sym.big;
    const std::vector<uint32_t>& lc = t.l_colptr;
    const std::vector<uint32_t>& lr = t.l_rowidx;
    // ---- supernodes: column j joins its predecessor's supernode when parent[j-1] == j and
    // struct(j) == struct(j-1) \ {j-1}
    std::vector<uint32_t> col2sn(n);
    for (uint32_t j = 0; j < n; j++) {
        const bool joins = j > 0 && t.parent[j - 1] == (int32_t)j && (lc[j + 1] - lc[j]) + 1 == (lc[j] - lc[j - 1]);
        if (!joins) c0.push_back(j);
        col2sn[j] = (uint32_t)c0.size() - 1;
    }
    const uint32_t S = (uint32_t)c0.size();
    ns.assign(S, 0); f.assign(S, 0); rows_off.assign(S + 1, 0); rel_off.assign(S + 1, 0);
    pan_off.assign(S, 0); upd_off.assign(S, 0);
    sparent.assign(S, -1);
    uint64_t pan_total = 0, upd_total = 0;
    flops_ = 0;
    for (uint32_t s = 0; s < S; s++) {
        const uint32_t last = (s + 1 < S ? c0[s + 1] : n) - 1;
        ns[s] = last - c0[s] + 1;
        f[s] = lc[c0[s] + 1] - lc[c0[s]];
        rows_off[s + 1] = rows_off[s] + f[s];
        rel_off[s + 1] = rel_off[s] + (f[s] - ns[s]);
        pan_off[s] = pan_total;
        pan_total += (uint64_t)f[s] * ns[s];
        upd_off[s] = upd_total;
        upd_total += (uint64_t)(f[s] - ns[s]) * (f[s] - ns[s]);
        sparent[s] = t.parent[last] >= 0 ? (int32_t)col2sn[t.parent[last]] : -1;
        if (ns[s] > 255u * TB) {
            if (err) *err = "multifrontal: supernode wider than 16,320 columns";
            return cudaErrorInvalidValue;
        }
    }
    for (uint32_t j = 0; j < n; j++) flops_ += (uint64_t)(lc[j + 1] - lc[j]) * (lc[j + 1] - lc[j]);
    pan_total_ = pan_total;
    rows.assign(rows_off[S], 0);
    for (uint32_t s = 0; s < S; s++) std::copy(lr.begin() + lc[c0[s]], lr.begin() + lc[c0[s]] + f[s], rows.begin() + rows_off[s]);
    lap("supernodes");
    // L position -> panel offset; diagonal positions
    diag_map_.assign(n, 0);
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t s = col2sn[j], k = j - c0[s];
        const uint64_t base = pan_off[s] + (uint64_t)k * f[s] + k;
        diag_map_[j] = base;
    }
    lap("position maps");
    // children (ascending) and relative indices
    child_ptr.assign(S + 1, 0);
    for (uint32_t s = 0; s < S; s++)
        if (sparent[s] >= 0) child_ptr[sparent[s] + 1]++;
    for (uint32_t s = 0; s < S; s++) child_ptr[s + 1] += child_ptr[s];
    child.resize(child_ptr[S]);
    {
        std::vector<uint32_t> fill(child_ptr.begin(), child_ptr.end() - 1);
        for (uint32_t s = 0; s < S; s++)
            if (sparent[s] >= 0) child[fill[sparent[s]]++] = s;
    }
    rel.assign(rel_off[S], 0);
    for (uint32_t s = 0; s < S; s++) {
        if (sparent[s] < 0) {
            if (f[s] != ns[s]) {
                if (err) *err = "multifrontal: root supernode with an update block";
                return cudaErrorInvalidValue;
            }
            continue;
        }
        const uint32_t p = (uint32_t)sparent[s];
        const uint32_t* rp = rows.data() + rows_off[p];
        const uint32_t* rs = rows.data() + rows_off[s] + ns[s];
        uint32_t pos = 0;
        for (uint32_t a = 0; a < f[s] - ns[s]; a++) {
            while (pos < f[p] && rp[pos] < rs[a]) pos++;
            if (pos >= f[p] || rp[pos] != rs[a]) {
                if (err) *err = "multifrontal: update row missing from the parent front";
                return cudaErrorInvalidValue;
            }
            rel[rel_off[s] + a] = pos;
        }
    }
    lap("children, relative indices");
    // ---- small subtrees / big levels
    big.assign(S, 0);
    for (uint32_t s = 0; s < S; s++) {
        if (f[s] > (uint32_t)kSmallFront) big[s] = 1;
        if (big[s] && sparent[s] >= 0) big[sparent[s]] = 1;
    }
    std::vector<int32_t> sub_of(S, -1);
    uint32_t nsub = 0;
    for (uint32_t s = S; s-- > 0;) {
        if (big[s]) continue;
        if (sparent[s] < 0 || big[sparent[s]]) sub_of[s] = (int32_t)nsub++;
        else sub_of[s] = sub_of[sparent[s]];
    }
    std::vector<uint64_t> sub_work(nsub, 0);
    std::vector<uint32_t> sub_cnt(nsub, 0);
    for (uint32_t s = 0; s < S; s++)
        if (!big[s]) {
            sub_work[sub_of[s]] += (uint64_t)f[s] * f[s] * ns[s] + 64;
            sub_cnt[sub_of[s]]++;
        }
    std::vector<uint32_t> sub_order(nsub);
    std::iota(sub_order.begin(), sub_order.end(), 0u);
    std::stable_sort(sub_order.begin(), sub_order.end(), [&](uint32_t a, uint32_t b) { return sub_work[a] > sub_work[b]; });
    std::vector<uint32_t> sub_rank(nsub);
    sub_ptr.assign(nsub + 1, 0);
    for (uint32_t k = 0; k < nsub; k++) sub_rank[sub_order[k]] = k;
    for (uint32_t k = 0; k < nsub; k++) sub_ptr[k + 1] = sub_ptr[k] + sub_cnt[sub_order[k]];
    sub_list.assign(sub_ptr[nsub], 0);
    {
        std::vector<uint32_t> fill(sub_ptr.begin(), sub_ptr.end() - 1);
        for (uint32_t s = 0; s < S; s++)
            if (!big[s]) sub_list[fill[sub_rank[sub_of[s]]]++] = s;
    }
    nsub_ = nsub;
    level.assign(S, 0); winv_blk.assign(S, 0);
    uint32_t nlevels = 0, nbig = 0, max_front = 1, winv_blocks = 0;
    for (uint32_t s = 0; s < S; s++) {
        max_front = std::max(max_front, f[s]);
        if (!big[s]) continue;
        nbig++;
        winv_blk[s] = winv_blocks;
        winv_blocks += (ns[s] + TB - 1) / TB;
        nlevels = std::max(nlevels, level[s] + 1);
        if (sparent[s] >= 0) level[sparent[s]] = std::max(level[sparent[s]], level[s] + 1);
    }
    std::vector<std::vector<uint32_t>> by_level(nlevels);
    for (uint32_t s = 0; s < S; s++)
        if (big[s]) by_level[level[s]].push_back(s);
    lap("subtrees, levels");
    // ---- task lists of the factorisation
    tasks.clear();  // uint4 each
    auto push_task = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
        tasks.push_back(a); tasks.push_back(b); tasks.push_back(c); tasks.push_back(d);
    };
    factor_seq_.clear();
    fwd_tasks_.clear();
    level_wide_.clear();
    level_narrow_.clear();
    bwd_tasks_.clear();
    level_ptr_.assign(1, 0);
    level_list.clear();
    level_seq_ptr_.clear();
    for (uint32_t l = 0; l < nlevels; l++) {
        const std::vector<uint32_t>& L = by_level[l];
        level_seq_ptr_.push_back((uint32_t)factor_seq_.size());
        for (uint32_t s : L) level_list.push_back(s);
        level_ptr_.push_back((uint32_t)level_list.size());
        {
            bool wide = false;
            for (uint32_t s : L) wide = wide || ns[s] >= 64;
            level_wide_.push_back(wide);
            bool narrow = true;
            for (uint32_t s : L) narrow = narrow && f[s] <= 64;
            level_narrow_.push_back(narrow);
        }
        {   // solve tasks of the split supernodes of this level
            uint32_t first = (uint32_t)(tasks.size() / 4);
            for (uint32_t s : L)
                if ((uint64_t)(f[s] - ns[s]) * ns[s] > kSplitEntries)
                    for (uint32_t r0 = 0; r0 < f[s] - ns[s]; r0 += 64) push_task(s, r0, std::min<uint32_t>(64, f[s] - ns[s] - r0), 0);
            fwd_tasks_.push_back({first, (uint32_t)(tasks.size() / 4) - first});
            first = (uint32_t)(tasks.size() / 4);
            for (uint32_t s : L)
                if ((uint64_t)(f[s] - ns[s]) * ns[s] > kSplitEntries)
                    for (uint32_t j0 = 0; j0 < ns[s]; j0 += 32) push_task(s, j0, std::min<uint32_t>(32, ns[s] - j0), 0);
            bwd_tasks_.push_back({first, (uint32_t)(tasks.size() / 4) - first});
        }
        uint32_t first = (uint32_t)(tasks.size() / 4), maxnsb = 0;
        for (uint32_t s : L) {
            maxnsb = std::max(maxnsb, (ns[s] + TB - 1) / TB);
            const bool has_children = child_ptr[s + 1] > child_ptr[s];
            const uint32_t cbw = f[s] > 256 ? 8 : TB;  // big fronts: one warp per target column, many CTAs
            for (uint32_t lo = 0; lo < f[s]; lo += cbw) {
                const uint32_t hi = std::min(f[s], lo + cbw);
                if (!has_children && hi <= ns[s]) continue;  // nothing to zero, nothing to add
                push_task(s, lo, hi, 0);
            }
        }
        uint32_t count = (uint32_t)(tasks.size() / 4) - first;
        if (count) factor_seq_.push_back({0, first, count});
        // blocks of a front after pivot block kb: later pivot blocks, then 64-row blocks of the update part
        auto blocks_after = [&](uint32_t s, uint32_t kb, std::vector<std::pair<uint32_t, uint32_t>>& out) {
            out.clear();
            for (uint32_t r0 = (kb + 1) * TB; r0 < ns[s]; r0 += TB) out.push_back({r0, std::min<uint32_t>(TB, ns[s] - r0)});
            for (uint32_t r0 = ns[s]; r0 < f[s]; r0 += TB) out.push_back({r0, std::min<uint32_t>(TB, f[s] - r0)});
        };
        std::vector<std::pair<uint32_t, uint32_t>> blk;
        for (uint32_t kb = 0; kb < maxnsb; kb++) {
            first = (uint32_t)(tasks.size() / 4);
            for (uint32_t s : L)
                if (kb * TB < ns[s]) push_task(s, kb * TB, kb * TB, std::min<uint32_t>(TB, ns[s] - kb * TB) << 16);
            count = (uint32_t)(tasks.size() / 4) - first;
            if (count) factor_seq_.push_back({1, first, count});
            first = (uint32_t)(tasks.size() / 4);
            for (uint32_t s : L) {
                if (kb * TB >= ns[s]) continue;
                const uint32_t nc = std::min<uint32_t>(TB, ns[s] - kb * TB);
                blocks_after(s, kb, blk);
                for (auto& b : blk) push_task(s, b.first, kb * TB, b.second | (nc << 16));
            }
            count = (uint32_t)(tasks.size() / 4) - first;
            if (count) factor_seq_.push_back({2, first, count});
            first = (uint32_t)(tasks.size() / 4);
            for (uint32_t s : L) {
                if (kb * TB >= ns[s]) continue;
                const uint32_t nc = std::min<uint32_t>(TB, ns[s] - kb * TB);
                blocks_after(s, kb, blk);
                for (size_t bj = 0; bj < blk.size(); bj++)
                    for (size_t bi = bj; bi < blk.size(); bi++)
                        push_task(s, blk[bi].first, blk[bj].first,
                                  (blk[bi].second - 1) | ((blk[bj].second - 1) << 8) | ((nc - 1) << 16) | (kb << 24));
            }
            count = (uint32_t)(tasks.size() / 4) - first;
            if (count) factor_seq_.push_back({3, first, count});
        }
    }
    lap("factor task lists");
    level_seq_ptr_.push_back((uint32_t)factor_seq_.size());
    // ---- chained solves: wide levels whose 64-row chunks all fit on the device at once
    chain_.assign(nlevels, ChainLevel());
    chain_tasks_.clear();
    std::fill(winv_blk.begin(), winv_blk.end(), 0u);
    winv_blocks = 0;
    inv_first_ = inv_count_ = 0;
    flow_asm_ptr_.clear();
    flow_asm_.clear();
    flow_pub_lo_ = ~0ull;
    flow_pub_hi_ = 0;
    {
        static const bool off = std::getenv("FK_NO_CHAIN") != nullptr;  // debug / A-B knob
        auto push_chain = [&](uint32_t a, uint32_t b, uint32_t c) {
            chain_tasks_.push_back(a); chain_tasks_.push_back(b); chain_tasks_.push_back(c); chain_tasks_.push_back(0);
        };
        std::vector<uint32_t> chained;
        for (uint32_t l = 0; l < nlevels && !off; l++) {
            if (!level_wide_[l]) continue;
            const std::vector<uint32_t>& L = by_level[l];
            uint64_t ctas = 0;
            for (uint32_t s : L) ctas += (ns[s] + TB - 1) / TB + (f[s] - ns[s] + TB - 1) / TB;
            static const uint32_t chain_max = [] {  // A/B knob; the default was a residency bound before the tickets
                const char* e = std::getenv("FK_CHAIN_MAX_CTAS");
                return (uint32_t)(e ? std::max(1, std::atoi(e)) : (int)kChainMaxCtas);
            }();
            if (ctas > chain_max) continue;
            ChainLevel& c = chain_[l];
            c.on = true;
            for (uint32_t s : L) {
                chained.push_back(s);
                winv_blk[s] = winv_blocks;  // first 64x64 inverse block / first flag of the supernode
                winv_blocks += (ns[s] + TB - 1) / TB;
            }
            c.fwd_first = (uint32_t)(chain_tasks_.size() / 4);
            for (uint32_t s : L) {
                for (uint32_t r0 = 0; r0 < ns[s]; r0 += TB) push_chain(s, r0, std::min<uint32_t>(TB, ns[s] - r0));
                for (uint32_t r0 = ns[s]; r0 < f[s]; r0 += TB) push_chain(s, r0, std::min<uint32_t>(TB, f[s] - r0));
            }
            c.fwd_count = (uint32_t)(chain_tasks_.size() / 4) - c.fwd_first;
            c.dot_first = (uint32_t)(chain_tasks_.size() / 4);
            for (uint32_t s : L)
                if (f[s] > ns[s])
                    for (uint32_t j0 = 0; j0 < ns[s]; j0 += 8) push_chain(s, j0, std::min<uint32_t>(8, ns[s] - j0));
            c.dot_count = (uint32_t)(chain_tasks_.size() / 4) - c.dot_first;
            c.bwd_first = (uint32_t)(chain_tasks_.size() / 4);
            for (uint32_t s : L)
                for (uint32_t b = (ns[s] + TB - 1) / TB; b-- > 0;) push_chain(s, b * TB, std::min<uint32_t>(TB, ns[s] - b * TB));
            c.bwd_count = (uint32_t)(chain_tasks_.size() / 4) - c.bwd_first;
        }
        // tile dataflow factorisation: every level whose fronts are larger than one tile (no residency condition:
        // tiles are numbered column by column, dependencies point to smaller block indices)
        static const bool no_flow = std::getenv("FK_NO_FLOW") != nullptr;  // debug / A-B knob
        // Measured on the 400x250 lattice (5 factorisations): dataflow on the levels with <= ~1,000 tiles 26.1 ms, on
        // every level with a front above 256 / 128 / 64 rows 26.7 / 27.0 / 27.5 ms, on none 30.3 ms: levels with many
        // supernodes are throughput bound and better served by the wide launch-per-step kernels.
        static const uint32_t flow_max_tiles = [] {
            const char* e = std::getenv("FK_FLOW_MAX_TILES");
            return (uint32_t)(e ? std::max(0, std::atoi(e)) : 1024);
        }();
        for (uint32_t l = 0; l < nlevels && !off && !no_flow; l++) {
            const std::vector<uint32_t>& L = by_level[l];
            uint64_t tiles = 0;
            uint32_t fmax = 0;
            for (uint32_t s : L) {
                const uint64_t nb = (ns[s] + TB - 1) / TB + (f[s] - ns[s] + TB - 1) / TB;
                tiles += nb * (nb + 1) / 2;
                fmax = std::max(fmax, f[s]);
            }
            if (fmax <= (uint32_t)TB || tiles > flow_max_tiles) continue;
            ChainLevel& c = chain_[l];
            c.flow_first = (uint32_t)(chain_tasks_.size() / 4);
            std::vector<std::pair<uint32_t, uint32_t>> blk;
            for (uint32_t s : L) {
                blk.clear();
                for (uint32_t r0 = 0; r0 < ns[s]; r0 += TB) blk.push_back({r0, std::min<uint32_t>(TB, ns[s] - r0)});
                for (uint32_t r0 = ns[s]; r0 < f[s]; r0 += TB) blk.push_back({r0, std::min<uint32_t>(TB, f[s] - r0)});
                for (size_t bj = 0; bj < blk.size(); bj++)
                    for (size_t bi = bj; bi < blk.size(); bi++) {
                        chain_tasks_.push_back(s); chain_tasks_.push_back(blk[bi].first); chain_tasks_.push_back(blk[bj].first);
                        chain_tasks_.push_back((blk[bi].second - 1) | ((blk[bj].second - 1) << 8));
                    }
                flow_pub_lo_ = std::min<uint64_t>(flow_pub_lo_, pan_off[s]);
                flow_pub_hi_ = std::max<uint64_t>(flow_pub_hi_, pan_off[s] + (uint64_t)f[s] * ns[s]);
            }
            c.flow_count = (uint32_t)(chain_tasks_.size() / 4) - c.flow_first;
            // extend-add ranges of every (tile, child): rows / columns of the child's update matrix inside the tile
            c.asm_first = (uint32_t)flow_asm_ptr_.size();
            for (uint32_t q = c.flow_first; q < c.flow_first + c.flow_count; q++) {
                const uint32_t s = chain_tasks_[4 * (size_t)q], row0 = chain_tasks_[4 * (size_t)q + 1], tcol0 = chain_tasks_[4 * (size_t)q + 2];
                const uint32_t w = chain_tasks_[4 * (size_t)q + 3], nrA = (w & 0xFFu) + 1, nrB = ((w >> 8) & 0xFFu) + 1;
                flow_asm_ptr_.push_back((uint32_t)(flow_asm_.size() / 4));
                for (uint32_t cq = child_ptr[s]; cq < child_ptr[s + 1]; cq++) {
                    const uint32_t ch = child[cq];
                    const uint32_t* rb = rel.data() + rel_off[ch];
                    const uint32_t* re = rel.data() + rel_off[ch + 1];
                    const uint32_t a0 = (uint32_t)(std::lower_bound(rb, re, row0) - rb), a1 = (uint32_t)(std::lower_bound(rb, re, row0 + nrA) - rb);
                    const uint32_t b0 = (uint32_t)(std::lower_bound(rb, re, tcol0) - rb), b1 = (uint32_t)(std::lower_bound(rb, re, tcol0 + nrB) - rb);
                    if (a1 <= a0 || b1 <= b0 || a1 - 1 < b0) continue;  // nothing of this child in the tile's lower part
                    if (a0 >= (1u << 16) || b0 >= (1u << 16)) {
                        if (err) *err = "multifrontal: update matrix too large for the packed extend-add ranges";
                        return cudaErrorInvalidValue;
                    }
                    flow_asm_.push_back(ch); flow_asm_.push_back(a0 | ((a1 - a0) << 16)); flow_asm_.push_back(b0 | ((b1 - b0) << 16)); flow_asm_.push_back(0);
                }
            }
            flow_asm_ptr_.push_back((uint32_t)(flow_asm_.size() / 4));
        }
        inv_first_ = (uint32_t)(chain_tasks_.size() / 4);
        for (uint32_t s : chained)
            for (uint32_t r0 = 0; r0 < ns[s]; r0 += TB) push_chain(s, r0, std::min<uint32_t>(TB, ns[s] - r0));
        inv_count_ = (uint32_t)(chain_tasks_.size() / 4) - inv_first_;
    }
    {   // mid-size levels (see mf_mid_factor_kernel)
        static const bool no_mid = std::getenv("FK_NO_MID") != nullptr;  // debug / A-B knob
        mid_.assign(nlevels, MidLevel());
        mid_list_.clear();
        for (uint32_t l = 0; l < nlevels && !no_mid; l++) {
            const std::vector<uint32_t>& L = by_level[l];
            uint32_t fmax = 0;
            for (uint32_t s : L) fmax = std::max(fmax, f[s]);
            if (L.empty() || fmax > (uint32_t)kMidFront) continue;
            chain_[l].flow_count = 0;  // (a mid-size level is never a dataflow level: the same kernel with FK_NO_FLOW and without)
            mid_[l].first = (uint32_t)mid_list_.size();
            mid_[l].count = (uint32_t)L.size();
            mid_[l].ld = fmax | 1u;
            // heavier fronts first: the last wave of the launch is the light one
            std::vector<uint32_t> order(L);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (uint64_t)f[a] * f[a] * ns[a] > (uint64_t)f[b] * f[b] * ns[b]; });
            mid_list_.insert(mid_list_.end(), order.begin(), order.end());
        }
    }
    n_chain_flags_ = winv_blocks;
    lap("chain task lists");
    factor_launches_ = factor_seq_.size() + (nsub ? 1 : 0);
    stats.supernodes = S; stats.small_subtrees = nsub; stats.big = nbig; stats.levels = nlevels;
    stats.max_front = max_front; stats.upd_doubles = upd_total;
    sym.n = n; sym.S = S; sym.nsub = nsub; sym.nlevels = nlevels; sym.nbig = nbig; sym.max_front = max_front;
    sym.winv_blocks = winv_blocks; sym.pan_total = pan_total; sym.upd_total = upd_total;
    return cudaSuccess;
}

cudaError_t Multifrontal::init(const Topology& t, cudaStream_t stream, std::string* err) {
    (void)stream;
    MF_CU(build_symbolic(t, err));
    const uint32_t n = sym.n, S = sym.S, max_front = sym.max_front;
    const uint64_t pan_total = sym.pan_total, upd_total = sym.upd_total;
    const std::vector<uint32_t>&c0 = sym.c0, &ns = sym.ns, &f = sym.f, &rows_off = sym.rows_off, &rows = sym.rows, &rel_off = sym.rel_off,
                               &rel = sym.rel, &child_ptr = sym.child_ptr, &child = sym.child, &winv_blk = sym.winv_blk, &sub_ptr = sym.sub_ptr,
                               &sub_list = sym.sub_list, &level_list = sym.level_list, &tasks = sym.tasks;
    const std::vector<uint64_t>&pan_off = sym.pan_off, &upd_off = sym.upd_off;
    // ---- upload
    dev_.S = S;
    MF_CU(upload_vec(c0, &dev_.c0, owned_));
    MF_CU(upload_vec(ns, &dev_.ns, owned_));
    MF_CU(upload_vec(f, &dev_.f, owned_));
    MF_CU(upload_vec(rows_off, &dev_.rows_off, owned_));
    MF_CU(upload_vec(rows, &dev_.rows, owned_));
    MF_CU(upload_vec(pan_off, &dev_.pan_off, owned_));
    MF_CU(upload_vec(upd_off, &dev_.upd_off, owned_));
    MF_CU(upload_vec(rel_off, &dev_.rel_off, owned_));
    MF_CU(upload_vec(rel, &dev_.rel, owned_));
    MF_CU(upload_vec(child_ptr, &dev_.child_ptr, owned_));
    MF_CU(upload_vec(child, &dev_.child, owned_));
    MF_CU(upload_vec(winv_blk, &dev_.winv_blk, owned_));
    MF_CU(upload_vec(sub_ptr, &d_sub_ptr_, owned_));
    MF_CU(upload_vec(sub_list, &d_sub_list_, owned_));
    MF_CU(upload_vec(level_list, &d_level_list_, owned_));
    {
        const uint32_t* p = nullptr;
        MF_CU(upload_vec(tasks, &p, owned_));
        d_tasks_ = (const uint4*)p;
    }
    MF_CU(alloc_vec(&dev_.pan, pan_total, owned_));
    MF_CU(alloc_vec(&dev_.upd, upd_total, owned_));
    MF_CU(alloc_vec(&dev_.winv, std::max<size_t>(1, (size_t)sym.winv_blocks * TB * TB), owned_));  // chained levels only
    MF_CU(alloc_vec(&dev_.ubuf, rel_off[S], owned_));
    MF_CU(alloc_vec(&dev_.status, 1, owned_));
    MF_CU(alloc_vec(&d_tmp_, n, owned_));
    {
        const uint32_t* p = nullptr;
        MF_CU(upload_vec(chain_tasks_, &p, owned_));
        d_chain_tasks_ = (const uint4*)p;
        MF_CU(alloc_vec(&d_chain_pub_, 2 * (size_t)std::max<uint32_t>(1, n), owned_));
        int dev = 0, sms = 0;
        MF_CU(cudaGetDevice(&dev));
        MF_CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        (void)sms;  // (no residency condition any more: the polling kernels take their tasks by ticket)
        // one ticket counter per polling launch: [level] dataflow factorisation, [L + level] chained forward,
        // [2L + level] chained backward; cleared by a memset node at the head of the factor / solve graph
        n_tickets_ = 3 * (uint32_t)chain_.size() + 1;
        MF_CU(alloc_vec(&d_tickets_, n_tickets_, owned_));
        MF_CU(cudaMemset(d_tickets_, 0, sizeof(uint32_t) * n_tickets_));
        MF_CU(cudaFuncSetAttribute(mf_chain_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainInvSmem));
        MF_CU(cudaFuncSetAttribute(mf_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFlowSmem));
        bool any_flow = false;
        for (const ChainLevel& c : chain_) any_flow = any_flow || c.flow_count;
        if (any_flow) {  // published panel tiles of the dataflow levels (addressed like the panel storage)
            MF_CU(alloc_vec(&d_pan_pub_, (size_t)(flow_pub_hi_ - flow_pub_lo_), owned_));
            MF_CU(upload_vec(flow_asm_ptr_, &d_flow_asm_ptr_, owned_));
            const uint32_t* pe = nullptr;
            MF_CU(upload_vec(flow_asm_, &pe, owned_));
            d_flow_asm_ = (const uint4*)pe;
        }
    }
    MF_CU(cudaFuncSetAttribute(mf_small_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallFactorSmem));
    MF_CU(upload_vec(mid_list_, &d_mid_list_, owned_));
    MF_CU(cudaFuncSetAttribute(mf_mid_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(kMidFront | 1) * (kMidFront | 1) * sizeof(double))));
    MF_CU(cudaFuncSetAttribute(mf_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColSmem));
    MF_CU(cudaFuncSetAttribute(mf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem));
    MF_CU(cudaFuncSetAttribute(mf_small_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSolveSmem));
    MF_CU(cudaFuncSetAttribute(mf_small_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSolveSmem));
    const size_t big_solve_smem = (32 * 33 + 2 * (size_t)max_front) * sizeof(double);
    if (big_solve_smem > 200 * 1024) {
        if (err) *err = "multifrontal: a front is too large for the shared-memory solve vector";
        return cudaErrorInvalidValue;
    }
    MF_CU(cudaFuncSetAttribute(mf_big_solve_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    MF_CU(cudaFuncSetAttribute(mf_big_solve_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    MF_CU(cudaFuncSetAttribute(mf_big_solve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    MF_CU(cudaFuncSetAttribute(mf_big_solve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    MF_CU(cudaFuncSetAttribute(mf_fwd_upd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    MF_CU(cudaFuncSetAttribute(mf_bwd_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_solve_smem));
    return cudaSuccess;
}

cudaError_t Multifrontal::enqueue_factor(cudaStream_t st) {
    // FK_MF_TIMING=1 (debug, implies no graph): CUDA events around every launch, per-kind sums to stderr
    static const bool timing = std::getenv("FK_MF_TIMING") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float sums[5] = {0, 0, 0, 0, 0}, maxs[5] = {0, 0, 0, 0, 0};
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    static const bool detail = timing && std::atoi(std::getenv("FK_MF_TIMING")) >= 3;
    int seq_no = -1;
    uint32_t cur_count = 0;
    auto timed = [&](int kind, auto&& launch) {
        if (timing) cudaEventRecord(e0, st);
        launch();
        if (timing) {
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            sums[kind] += ms;
            maxs[kind] = std::max(maxs[kind], ms);
            if (detail) fprintf(stderr, "[mf factor launch] #%d kind %d ctas %u %.1f us\n", seq_no, kind, cur_count, ms * 1e3f);
        }
    };
    if (nsub_) {
        const uint32_t grid = (nsub_ + kWarpsPerCta - 1) / kWarpsPerCta;
        timed(4, [&] { mf_small_factor_kernel<<<grid, kWarpsPerCta * 32, kSmallFactorSmem, st>>>(dev_, d_sub_ptr_, d_sub_list_, nsub_); });
    }
    if (d_pan_pub_) {
        cudaError_t e = cudaMemsetAsync(d_pan_pub_, 0xFF, (size_t)(flow_pub_hi_ - flow_pub_lo_) * sizeof(double), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(d_tickets_, 0, sizeof(uint32_t) * chain_.size(), st);
        if (e != cudaSuccess) return e;
    }
    const uint32_t nlv = (uint32_t)level_seq_ptr_.size() - 1;
    for (uint32_t lv = 0; lv < nlv; lv++) {
        const bool flow = d_pan_pub_ && chain_[lv].flow_count;
        const bool mid = !flow && lv < mid_.size() && mid_[lv].count;
        if (mid) {  // the level's assembly / diag / panel / update launches are replaced by one launch: a CTA per front
            seq_no++;
            cur_count = mid_[lv].count;
            const uint32_t ld = mid_[lv].ld;
            timed(1, [&] {
                mf_mid_factor_kernel<<<mid_[lv].count, kMidThreads, (size_t)ld * ld * sizeof(double), st>>>(dev_, d_mid_list_ + mid_[lv].first, ld);
            });
            continue;
        }
        for (uint32_t qi = level_seq_ptr_[lv]; qi < level_seq_ptr_[lv + 1]; qi++) {
            const Launch& l = factor_seq_[qi];
            if (flow) continue;  // the level's assembly / diag / panel / update launches are replaced below
            const uint4* tk = d_tasks_ + l.first;
            seq_no++;
            cur_count = l.count;
            timed(l.kind, [&] {
                switch (l.kind) {
                    case 0: mf_asm_kernel<<<l.count, 256, 0, st>>>(dev_, tk); break;
                    case 1: mf_diag_kernel<<<l.count, kDiagThreads, kDiagSmem, st>>>(dev_, tk); break;
                    case 2: mf_col_kernel<<<l.count, kColThreads, kColSmem, st>>>(dev_, tk); break;
                    default: mf_rupd_kernel<<<l.count, kTileThreads, 0, st>>>(dev_, tk); break;
                }
            });
        }
        if (flow) {
            seq_no++;
            cur_count = chain_[lv].flow_count;
            // the published buffer is addressed with the panel offsets: bias the base so that pub + pan_off lands in it
            timed(3, [&] {
                mf_flow_kernel<<<chain_[lv].flow_count, kTileThreads, kFlowSmem, st>>>(dev_, d_chain_tasks_ + chain_[lv].flow_first,
                                                                                    d_pan_pub_ - flow_pub_lo_, d_flow_asm_ptr_ + chain_[lv].asm_first,
                                                                                    d_flow_asm_, d_tickets_ + lv);
            });
        }
    }
    if (inv_count_) timed(4, [&] { mf_chain_inv_kernel<<<inv_count_, TB, kChainInvSmem, st>>>(dev_, d_chain_tasks_ + inv_first_); });
    if (timing) {
        fprintf(stderr, "[mf timing] asm %.3f (max %.3f)  diag %.3f (max %.3f)  col %.3f (max %.3f)  rupd %.3f (max %.3f)  small %.3f ms\n",
                sums[0], maxs[0], sums[1], maxs[1], sums[2], maxs[2], sums[3], maxs[3], sums[4]);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return cudaGetLastError();
}

cudaError_t Multifrontal::enqueue_solve(double* w, double* delta, const int32_t* d_perm, cudaStream_t st) {
    static const bool timing = std::getenv("FK_MF_TIMING") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float sums[6] = {0, 0, 0, 0, 0, 0}, maxs[6] = {0, 0, 0, 0, 0, 0};
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    static const bool detail = timing && std::atoi(std::getenv("FK_MF_TIMING")) >= 2;
    int cur_level = -1;
    auto timed = [&](int kind, auto&& launch) {
        if (timing) cudaEventRecord(e0, st);
        launch();
        if (timing) {
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            sums[kind] += ms;
            maxs[kind] = std::max(maxs[kind], ms);
            if (detail) fprintf(stderr, "[mf solve launch] level %d kind %d %.1f us\n", cur_level, kind, ms * 1e3f);
        }
    };
    const uint32_t sgrid = (nsub_ + kWarpsPerCta - 1) / kWarpsPerCta;
    const size_t smem = (32 * 33 + 2 * (size_t)stats.max_front) * sizeof(double);
    const size_t vsmem = (size_t)stats.max_front * sizeof(double);
    const size_t nsmem = (32 * 33 + 2 * 64) * sizeof(double);  // narrow levels: fronts of order <= 64
    const uint32_t nlevels = (uint32_t)level_ptr_.size() - 1;
    bool any_chain = false;
    for (const ChainLevel& c : chain_) any_chain = any_chain || c.on;
    if (any_chain) {
        cudaError_t e = cudaMemsetAsync(d_chain_pub_, 0xFF, 2 * (size_t)n_ * sizeof(double), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(d_tickets_ + chain_.size(), 0, sizeof(uint32_t) * 2 * chain_.size(), st);
        if (e != cudaSuccess) return e;
    }
    if (nsub_) timed(0, [&] { mf_small_solve_kernel<true><<<sgrid, kWarpsPerCta * 32, kSmallSolveSmem, st>>>(dev_, d_sub_ptr_, d_sub_list_, nsub_, w, delta, d_perm); });
    for (uint32_t l = 0; l < nlevels; l++) {
        const uint32_t cnt = level_ptr_[l + 1] - level_ptr_[l];
        const uint32_t* list = d_level_list_ + level_ptr_[l];
        cur_level = (int)l;
        if (chain_[l].on) {
            timed(1, [&] { mf_chain_fwd_kernel<<<chain_[l].fwd_count, 256, 0, st>>>(dev_, d_chain_tasks_ + chain_[l].fwd_first, w, d_chain_pub_, d_tickets_ + chain_.size() + l); });
            continue;
        }
        timed(1, [&] {
            if (level_wide_[l]) mf_big_solve_kernel<true, true><<<cnt, 256, smem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, stats.max_front);
            else if (level_narrow_[l]) mf_big_solve_kernel<true, false, 64><<<cnt, 64, nsmem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, 64);
            else mf_big_solve_kernel<true, false><<<cnt, 256, smem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, stats.max_front);
        });
        if (fwd_tasks_[l].second) timed(2, [&] { mf_fwd_upd_kernel<<<fwd_tasks_[l].second, 256, vsmem, st>>>(dev_, d_tasks_ + fwd_tasks_[l].first, w); });
    }
    for (uint32_t l = nlevels; l-- > 0;) {
        const uint32_t cnt = level_ptr_[l + 1] - level_ptr_[l];
        const uint32_t* list = d_level_list_ + level_ptr_[l];
        cur_level = (int)l;
        if (chain_[l].on) {
            if (chain_[l].dot_count)
                timed(3, [&] { mf_bwd_dot_kernel<<<chain_[l].dot_count, 256, vsmem, st>>>(dev_, d_chain_tasks_ + chain_[l].dot_first, w, d_tmp_); });
            timed(4, [&] {
                mf_chain_bwd_kernel<<<chain_[l].bwd_count, 256, 0, st>>>(dev_, d_chain_tasks_ + chain_[l].bwd_first, w, delta, d_perm, d_tmp_,
                                                                      d_chain_pub_ + n_, d_tickets_ + 2 * chain_.size() + l);
            });
            continue;
        }
        if (bwd_tasks_[l].second) timed(3, [&] { mf_bwd_dot_kernel<<<bwd_tasks_[l].second, 256, vsmem, st>>>(dev_, d_tasks_ + bwd_tasks_[l].first, w, d_tmp_); });
        timed(4, [&] {
            if (level_wide_[l]) mf_big_solve_kernel<false, true><<<cnt, 256, smem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, stats.max_front);
            else if (level_narrow_[l]) mf_big_solve_kernel<false, false, 64><<<cnt, 64, nsmem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, 64);
            else mf_big_solve_kernel<false, false><<<cnt, 256, smem, st>>>(dev_, list, w, delta, d_perm, d_tmp_, stats.max_front);
        });
    }
    if (nsub_) timed(5, [&] { mf_small_solve_kernel<false><<<sgrid, kWarpsPerCta * 32, kSmallSolveSmem, st>>>(dev_, d_sub_ptr_, d_sub_list_, nsub_, w, delta, d_perm); });
    if (timing) {
        fprintf(stderr, "[mf solve timing] small fwd %.3f  big fwd %.3f (max %.3f)  fwd upd %.3f (max %.3f)  bwd dot %.3f (max %.3f)  big bwd %.3f (max %.3f)  small bwd %.3f ms\n",
                sums[0], sums[1], maxs[1], sums[2], maxs[2], sums[3], maxs[3], sums[4], maxs[4], sums[5]);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return cudaGetLastError();
}

// The launch sequences are static: captured once into CUDA graphs and replayed.
cudaError_t Multifrontal::factor(cudaStream_t st) {
    static const bool no_graph = std::getenv("FK_NO_GRAPH") != nullptr || std::getenv("FK_MF_TIMING") != nullptr;
    if (no_graph || factor_launches_ < 8) return enqueue_factor(st);
    if (!factor_graph_) {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) return e;
        cudaError_t e1 = enqueue_factor(st);
        e = cudaStreamEndCapture(st, &g);
        if (e1 != cudaSuccess) return e1;
        if (e != cudaSuccess) return e;
        e = cudaGraphInstantiate(&factor_graph_, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return e;
    }
    return cudaGraphLaunch(factor_graph_, st);
}

cudaError_t Multifrontal::solve(double* w, double* delta, const int32_t* d_perm, cudaStream_t st) {
    static const bool no_graph = std::getenv("FK_NO_GRAPH") != nullptr || std::getenv("FK_MF_TIMING") != nullptr;
    if (no_graph || level_ptr_.size() < 5) return enqueue_solve(w, delta, d_perm, st);
    if (solve_graph_ && (w != solve_w_ || delta != solve_delta_ || d_perm != solve_perm_)) {
        cudaGraphExecDestroy(solve_graph_);
        solve_graph_ = nullptr;
    }
    if (!solve_graph_) {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) return e;
        cudaError_t e1 = enqueue_solve(w, delta, d_perm, st);
        e = cudaStreamEndCapture(st, &g);
        if (e1 != cudaSuccess) return e1;
        if (e != cudaSuccess) return e;
        e = cudaGraphInstantiate(&solve_graph_, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return e;
        solve_w_ = w; solve_delta_ = delta; solve_perm_ = d_perm;
    }
    return cudaGraphLaunch(solve_graph_, st);
}

}  // namespace fk
