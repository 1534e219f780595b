// How fast can ONE warp issue independent DFMAs, and how many warps per SM does the FP64 pipe need?
// One CTA of W warps on one SM; every thread runs CH independent FMA chains for N steps; cycles by clock64.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, int n) {
    double a[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = 1.0 + threadIdx.x * 1e-3 + c;
    const double m = 1.0000001, b = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) a[c] = fma(a[c], m, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += a[c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 8);
    const int n = 4096;
    for (int ch : {1, 2, 4, 8, 16})
        for (int w : {1, 2, 4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; rep++) {
                if (ch == 1) k<1><<<1, 32 * w>>>(o, c, n);
                if (ch == 2) k<2><<<1, 32 * w>>>(o, c, n);
                if (ch == 4) k<4><<<1, 32 * w>>>(o, c, n);
                if (ch == 8) k<8><<<1, 32 * w>>>(o, c, n);
                if (ch == 16) k<16><<<1, 32 * w>>>(o, c, n);
            }
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            const double per_warp_instr = (double)h / ((double)n * ch);
            printf("%2d chains/thread, %2d warps: %.2f cycles per warp-DFMA (one warp's view), %.1f FMA lanes/clk on the SM\n", ch, w, per_warp_instr,
                   32.0 * w * n * ch / (double)h);
        }
    return 0;
}
