// Debug harness: times the phases of mf_diag_kernel on one 64x64 SPD tile (build: see tools/run_diag_harness.sh)
#include "../fiksi_b200/csrc/multifrontal.cu"
#include <cstdio>
using namespace fk;
int main() {
    const uint32_t f = 64, ns = 64;
    std::vector<double> A(f * ns, 0.0);
    for (uint32_t j = 0; j < ns; j++) for (uint32_t i = j; i < f; i++) A[j * f + i] = (i == j) ? 70.0 + i : 1.0 / (1 + i - j);
    MfDev D{};
    uint32_t h_c0 = 0, h_ns = ns, h_f = f, h_w = 0; uint64_t h_off = 0;
    uint32_t *d32; cudaMalloc(&d32, 64); uint64_t* d64; cudaMalloc(&d64, 64);
    cudaMemcpy(d32, &h_c0, 4, cudaMemcpyHostToDevice); cudaMemcpy(d32 + 1, &h_ns, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d32 + 2, &h_f, 4, cudaMemcpyHostToDevice); cudaMemcpy(d32 + 3, &h_w, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d64, &h_off, 8, cudaMemcpyHostToDevice);
    D.S = 1; D.c0 = d32; D.ns = d32 + 1; D.f = d32 + 2; D.winv_blk = d32 + 3; D.pan_off = d64;
    cudaMalloc(&D.pan, A.size() * 8); cudaMalloc(&D.winv, 4096 * 8); cudaMalloc(&D.ubuf, 64 * 8); cudaMalloc(&D.status, 4);
    cudaMemset(D.status, 0, 4);
    uint4 task{0, 0, 0, 64u << 16}; uint4* dt; cudaMalloc(&dt, 16); cudaMemcpy(dt, &task, 16, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(D.pan, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        mf_diag_kernel<<<1, kDiagThreads>>>(D, dt);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("rep %d: %.1f us (%s)\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    // check: L D L^T == A
    std::vector<double> L(A.size()); cudaMemcpy(L.data(), D.pan, A.size() * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (uint32_t i = 0; i < f; i++) for (uint32_t j = 0; j <= i; j++) {
        double sacc = 0; for (uint32_t k = 0; k <= j; k++) { double lik = (i == k) ? 1.0 : L[k * f + i], ljk = (j == k) ? 1.0 : L[k * f + j]; sacc += lik * L[k * f + k] * ljk; }
        maxerr = fmax(maxerr, fabs(sacc - A[j * f + i]));
    }
    printf("max |LDL^T - A| = %.3e\n", maxerr);
    return 0;
}
