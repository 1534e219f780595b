// FP64 throughput of the 4x4 register-blocked outer product (3 register operands per DFMA) on ONE SM,
// for 1..8 warps per SM sub-partition: cycles per warp-level DFMA per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, const double* in, long long* cyc, int iters) {
    double a[4], b[4], acc[4][4];
    for (int u = 0; u < 4; u++) { a[u] = in[threadIdx.x + 32 * u]; b[u] = in[threadIdx.x + 32 * (4 + u)]; }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(-a[i], b[j], acc[i][j]);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *d, *in; long long* c; cudaMalloc(&d, 1024 * 8); cudaMalloc(&in, 1024 * 8 * 8); cudaMalloc(&c, 8); cudaMemset(in, 0, 1024 * 64);
    const int iters = 2000;
    for (int threads : {32, 128, 256, 512, 1024}) {
        k<<<1, threads>>>(d, in, c, iters); k<<<1, threads>>>(d, in, c, iters); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        double warps_per_smsp = threads / 32.0 / 4.0; if (warps_per_smsp < 1) warps_per_smsp = 1;
        double dfma_per_smsp = 64.0 * iters * warps_per_smsp;
        printf("threads %4d: %.2f cycles per warp-DFMA per SMSP (%.1f DFMA lanes/clk/SM)\n", threads, h / dfma_per_smsp, 32.0 * 4 / (h / dfma_per_smsp));
    }
    return 0;
}
