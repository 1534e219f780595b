// Debug harness: mf_chain_fwd_kernel on one root-like supernode (f = ns = 832, 13 pivot chunks), with
// %globaltimer stamps per CTA: where does the time of one link of the chain go?
#define FK_CHAIN_PROFILE 1
#include "../fiksi_b200/csrc/multifrontal.cu"
#include <cstdio>
using namespace fk;
int main() {
    const uint32_t ns = 832, f = 832, B = ns / 64;
    std::vector<double> A((size_t)f * ns, 0.0);
    for (uint32_t j = 0; j < ns; j++) for (uint32_t i = j; i < f; i++) A[(size_t)j * f + i] = (i == j) ? 2.0 : 0.5 / (1 + i - j);
    MfDev D{};
    uint32_t h32[8] = {0, ns, f, 0, 0, 0, 0, 0};  // c0, ns, f, winv_blk, child_ptr[0..1], rel_off[0..1]
    uint64_t h_off = 0;
    uint32_t* d32; cudaMalloc(&d32, 64); uint64_t* d64; cudaMalloc(&d64, 64);
    cudaMemcpy(d32, h32, 32, cudaMemcpyHostToDevice); cudaMemcpy(d64, &h_off, 8, cudaMemcpyHostToDevice);
    D.S = 1; D.c0 = d32; D.ns = d32 + 1; D.f = d32 + 2; D.winv_blk = d32 + 3; D.child_ptr = d32 + 4; D.child = d32 + 6; D.rel_off = d32 + 6; D.rel = d32 + 6;
    D.pan_off = d64;
    cudaMalloc(&D.pan, A.size() * 8); cudaMemcpy(D.pan, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&D.ubuf, 64 * 8); cudaMalloc(&D.upd, 4096 * 8); cudaMalloc(&D.winv, (size_t)B * 4096 * 8);
    std::vector<uint4> tasks;
    for (uint32_t c = 0; c < B; c++) tasks.push_back({0, c * 64, 64, 0});
    uint4* dt; cudaMalloc(&dt, tasks.size() * 16); cudaMemcpy(dt, tasks.data(), tasks.size() * 16, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(mf_chain_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainInvSmem);
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(a);
            mf_chain_inv_kernel<<<B, TB, kChainInvSmem>>>(D, dt);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("inverse blocks: %.1f us (%s)\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        }
    }
    double* w; cudaMalloc(&w, ns * 8);
    double* pub; cudaMalloc(&pub, ns * 8);
    std::vector<double> rhs(ns, 1.0);
    uint32_t* ticket; cudaMalloc(&ticket, 4);  // the kernels take their tasks by ticket
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 4; rep++) {
        cudaMemcpy(w, rhs.data(), ns * 8, cudaMemcpyHostToDevice);
        cudaMemset(pub, 0xFF, ns * 8);
        cudaMemset(ticket, 0, 4);
        cudaEventRecord(a);
        mf_chain_fwd_kernel<<<B, 256>>>(D, dt, w, pub, ticket);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("rep %d: chain fwd %.1f us (%s)\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<long long> st(B * 8);
    cudaMemcpy(st.data(), D.upd, st.size() * 8, cudaMemcpyDeviceToHost);
    const long long t0 = st[0];
    for (uint32_t c = 0; c < B; c++) {
        printf("chunk %2u: start %6lld  wait-begin %6lld  flag %6lld  t-final %6lld  solved %6lld  released %6lld (ns)\n", c, st[c * 8] - t0,
               c ? st[c * 8 + 1] - t0 : 0, c ? st[c * 8 + 2] - t0 : 0, st[c * 8 + 3] - t0, st[c * 8 + 4] - t0, st[c * 8 + 5] - t0);
    }
    // check against a host forward substitution
    std::vector<double> y(ns), got(ns);
    for (uint32_t i = 0; i < ns; i++) { double v = 1.0; for (uint32_t k = 0; k < i; k++) v -= A[(size_t)k * f + i] * y[k]; y[i] = v; }
    cudaMemcpy(got.data(), w, ns * 8, cudaMemcpyDeviceToHost);
    double err = 0; for (uint32_t i = 0; i < ns; i++) err = fmax(err, fabs(got[i] - y[i]) / (1 + fabs(y[i])));
    printf("max rel err vs host = %.3e\n", err);
    return 0;
}
