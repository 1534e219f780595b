// Microbenchmark of the 8x8 register LDLt (micro_ldl) and the own-row substitution of the diagonal-tile kernel:
// cycles per call with 1 or 8 warps of a CTA active.
#include "../fiksi_b200/csrc/multifrontal.cu"
#include <cstdio>
using namespace fk;
template <int MODE>
__global__ void bench(double* out, long long* cyc, int warps, int reps) {
    __shared__ double Cs[TB * kTsLd];
    for (uint32_t e = threadIdx.x; e < TB * kTsLd; e += blockDim.x) {
        const uint32_t i = e / kTsLd, j = e % kTsLd;
        Cs[e] = i == j ? 50.0 + i : 1.0 / (1 + (i > j ? i - j : j - i));
    }
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= warps) return;
    double acc = 0.0;
    const uint32_t i = threadIdx.x & 63;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        const uint32_t k0 = (r & 7) * 8;
        double ll[MB][MB], inv[MB], piv[MB], y[MB];
        if (MODE == 0 || MODE == 2) micro_ldl(Cs, kTsLd, k0, 8, ll, inv, piv);
        else {
#pragma unroll
            for (int a = 0; a < MB; a++) {
                inv[a] = 1.0 + a;
#pragma unroll
                for (int b = 0; b < MB; b++) ll[a][b] = Cs[(k0 + a) * kTsLd + k0 + b];
            }
        }
        if (MODE >= 1) {
#pragma unroll
            for (int c = 0; c < MB; c++) y[c] = Cs[i * kTsLd + k0 + c] + acc;
#pragma unroll
            for (int cp = 0; cp + 1 < MB; cp++)
#pragma unroll
                for (int c = cp + 1; c < MB; c++) y[c] = fma(-y[cp], ll[c][cp], y[c]);
            acc += y[7] * inv[7];
        } else {
            acc += inv[7] + ll[7][3] + acc * 1e-30;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
}
// 8 column stores of one row value per call (stride f doubles between columns): plain, st.relaxed.gpu, both
template <int MODE>
__global__ void store_bench(double* T, double* Tp, long long* cyc, uint32_t f, int warps, int reps) {
    if ((int)(threadIdx.x >> 5) >= warps) return;
    const uint32_t i = threadIdx.x & 63, q = threadIdx.x >> 6;
    double v = 1.0 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        const uint32_t k0 = (r & 7) * 8;
#pragma unroll
        for (int c = 0; c < MB; c++) {
            if (MODE == 3 || (uint32_t)(c & 3) == q) {
                const double out = v * (1.0 + c);
                if (MODE != 1) T[(size_t)(k0 + c) * f + i] = out;
                if (MODE != 0) st_relaxed(Tp + (size_t)(k0 + c) * f + i, out);
            }
        }
        v += 1e-9;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
}
template <bool PUB>
__global__ void __launch_bounds__(256) tile_bench(double* T, double* Tp, int* status, long long* cyc, uint32_t f) {
    extern __shared__ __align__(16) double smt[];
    double* Cs = smt;
    double* Ys = smt + TB * kTsLd;
    for (uint32_t e = threadIdx.x; e < TB * kTsLd; e += blockDim.x) {
        const uint32_t i = e / kTsLd, j = e % kTsLd;
        Cs[e] = i == j ? 500.0 + i : 1.0 / (1 + (i > j ? i - j : j - i));
    }
    __syncthreads();
    long long g0, g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    diag_tile_factor<PUB>(Cs, Ys, 64, T, Tp, f, status);
    __syncthreads();
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = g1 - g0; }
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 16);
    {
        const uint32_t f = 832;
        double *T, *Tp; int* st; cudaMalloc(&T, (size_t)f * 64 * 8); cudaMalloc(&Tp, (size_t)f * 64 * 8); cudaMalloc(&st, 4096); cudaMemset(st, 0, 4096);
        cudaFuncSetAttribute(tile_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem);
        cudaFuncSetAttribute(tile_bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem);
        for (int pub = 0; pub < 2; pub++) {
            for (int rep = 0; rep < 3; rep++) {
                if (pub) tile_bench<true><<<1, 256, kDiagSmem>>>(T, Tp, st, c, f);
                else tile_bench<false><<<1, 256, kDiagSmem>>>(T, Tp, st, c, f);
            }
            cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
            printf("diag_tile_factor<%s> 64x64: %lld cycles, %lld ns by %%globaltimer -> SM clock %.0f MHz (%s)\n", pub ? "publishing" : "panel only", h[0], h[1], 1e3 * h[0] / h[1], cudaGetErrorString(cudaGetLastError()));
        }
#ifdef FK_CHAIN_PROFILE
        { long long hs[64]; cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost);
          for (int p = 0; p < 3; p++) { printf("micro-panel %d:", p); for (int k = 1; k <= 4; k++) printf("  stamp %d +%lld", k, hs[8 + p * 8 + k] - hs[8 + p * 8 + k - 1]); if (p) printf("  (since previous panel's stamp 0: %lld)", hs[8 + p * 8] - hs[8 + (p - 1) * 8]); printf("\n"); } }
#endif
        for (int rep = 0; rep < 400; rep++) tile_bench<true><<<1, 256, kDiagSmem>>>(T, Tp, st, c, f);  // ~60 ms of back-to-back launches
        cudaDeviceSynchronize();
        { long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
          printf("after 400 back-to-back launches: %lld cycles, %lld ns -> SM clock %.0f MHz\n", h[0], h[1], 1e3 * h[0] / h[1]); }
    }
    {
        const uint32_t f = 832;
        double *T, *Tp; cudaMalloc(&T, (size_t)f * 64 * 8); cudaMalloc(&Tp, (size_t)f * 64 * 8);
        const char* sn[] = {"plain stores (2 of 8 columns per thread)", "st.relaxed.gpu (2 of 8)", "plain + st.relaxed.gpu (2 of 8)", "plain + st.relaxed.gpu (all 8 per thread)"};
        for (int mode = 0; mode < 4; mode++)
            for (int warps : {2, 8}) {
                for (int rep = 0; rep < 2; rep++) {
                    if (mode == 0) store_bench<0><<<1, 256>>>(T, Tp, c, f, warps, 400);
                    if (mode == 1) store_bench<1><<<1, 256>>>(T, Tp, c, f, warps, 400);
                    if (mode == 2) store_bench<2><<<1, 256>>>(T, Tp, c, f, warps, 400);
                    if (mode == 3) store_bench<3><<<1, 256>>>(T, Tp, c, f, warps, 400);
                }
                cudaDeviceSynchronize();
                long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                printf("%-44s %d warps: %lld cycles per call (%s)\n", sn[mode], warps, h, cudaGetErrorString(cudaGetLastError()));
            }
    }
    const char* names[] = {"micro_ldl only", "substitution only", "micro_ldl + substitution"};
    for (int mode = 0; mode < 3; mode++)
        for (int warps : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) bench<0><<<1, 256>>>(d, c, warps, 400);
                if (mode == 1) bench<1><<<1, 256>>>(d, c, warps, 400);
                if (mode == 2) bench<2><<<1, 256>>>(d, c, warps, 400);
            }
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-26s %d warps: %lld cycles per call (%s)\n", names[mode], warps, h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
