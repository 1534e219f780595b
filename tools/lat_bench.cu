// Latency microbenchmark (one warp): dependent DFMA / DMUL / DADD chains, shuffles, reciprocal, division, sqrt.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double d) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0); r = fma(r, e, r); e = fma(-d, r, 1.0); r = fma(r, e, r); return r;
}
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b, int n) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x * b;
            if (OP == 2) x = x + b;
            if (OP == 3) x = __shfl_sync(0xFFFFFFFFu, x, (threadIdx.x + 1) & 31);
            if (OP == 4) x = fast_rcp(x) + a;
            if (OP == 5) x = a / x + a;
            if (OP == 6) x = sqrt(x) + a;
            if (OP == 7) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }
            if (OP == 8) { float f = __double2float_rn(x); asm volatile("" : "+f"(f)); x = f; x = fma(x, b, a); }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 8);
    const char* names[] = {"dfma", "dmul", "dadd", "shfl64", "fast_rcp+dadd", "div+dadd", "sqrt+dadd", "f64->f32 ffma ->f64", "cvt+dfma"};
    const int n = 1000;
#define RUN(OP) { k<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); k<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-22s %.1f cycles per op\n", names[OP], (double)h / (8.0 * n)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
