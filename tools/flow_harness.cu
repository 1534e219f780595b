// Debug harness: mf_flow_kernel on one root-like supernode (f = ns = 832: 13 pivot blocks, 91 tiles) with
// %globaltimer stamps per tile CTA: where does one link of the factorisation chain (diagonal tile -> panel
// tile -> next diagonal tile) spend its time?  Checks L D L^T = A.
#define FK_CHAIN_PROFILE 1
#include "../fiksi_b200/csrc/multifrontal.cu"
#include <cstdio>
using namespace fk;
int main() {
    const uint32_t ns = 832, f = 832, B = ns / 64;
    std::vector<double> A((size_t)f * ns, 0.0);
    for (uint32_t j = 0; j < ns; j++) for (uint32_t i = j; i < f; i++) A[(size_t)j * f + i] = (i == j) ? 900.0 + i : 1.0 / (1 + i - j);
    MfDev D{};
    uint32_t h32[4] = {0, ns, f, 0};
    uint64_t h_off = 0;
    uint32_t* d32; cudaMalloc(&d32, 64); uint64_t* d64; cudaMalloc(&d64, 64);
    cudaMemcpy(d32, h32, 16, cudaMemcpyHostToDevice); cudaMemcpy(d64, &h_off, 8, cudaMemcpyHostToDevice); cudaMemcpy(d64 + 1, &h_off, 8, cudaMemcpyHostToDevice);
    D.S = 1; D.c0 = d32; D.ns = d32 + 1; D.f = d32 + 2; D.winv_blk = d32 + 3; D.pan_off = d64; D.upd_off = d64 + 1;
    cudaMalloc(&D.pan, A.size() * 8); cudaMalloc(&D.ubuf, 8192 * 8); cudaMalloc(&D.upd, 64); cudaMalloc(&D.status, 4096); cudaMemset(D.status, 0, 4096);
    std::vector<uint4> tasks;
    for (uint32_t bj = 0; bj < B; bj++) for (uint32_t bi = bj; bi < B; bi++) tasks.push_back({0, bi * 64, bj * 64, 63u | (63u << 8)});
    uint4* dt; cudaMalloc(&dt, tasks.size() * 16); cudaMemcpy(dt, tasks.data(), tasks.size() * 16, cudaMemcpyHostToDevice);
    double* pub; cudaMalloc(&pub, A.size() * 8);
    uint32_t* aptr; cudaMalloc(&aptr, (tasks.size() + 1) * 4); cudaMemset(aptr, 0, (tasks.size() + 1) * 4);  // no children
    cudaFuncSetAttribute(mf_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFlowSmem);
    long long* dst = (long long*)D.status + 8;
    uint32_t* ticket; cudaMalloc(&ticket, 4);  // the kernel takes its tiles by ticket
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 4; rep++) {
        cudaMemcpy(D.pan, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMemset(pub, 0xFF, A.size() * 8);
        cudaEventRecord(a);
        cudaMemset(ticket, 0, 4);
        mf_flow_kernel<<<(unsigned)tasks.size(), kTileThreads, kFlowSmem>>>(D, dt, pub, aptr, nullptr, ticket);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("rep %d: flow factor %.1f us, %zu tiles (%s)\n", rep, ms * 1e3, tasks.size(), cudaGetErrorString(cudaGetLastError()));
    }
    {   // the first diagonal tile alone: one CTA, nothing else on the device
        cudaMemcpy(D.pan, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMemset(pub, 0xFF, A.size() * 8);
        cudaMemset(ticket, 0, 4);
        mf_flow_kernel<<<1, kTileThreads, kFlowSmem>>>(D, dt, pub, aptr, nullptr, ticket);
        cudaDeviceSynchronize();
        long long one[4]; cudaMemcpy(one, D.ubuf, sizeof(one), cudaMemcpyDeviceToHost);
        printf("tile (0,0) alone: %lld ns from the end of the update loop to the end (%s)\n", one[3] - one[2], cudaGetErrorString(cudaGetLastError()));
        cudaMemcpy(D.pan, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMemset(pub, 0xFF, A.size() * 8);
        cudaMemset(ticket, 0, 4);
        mf_flow_kernel<<<(unsigned)tasks.size(), kTileThreads, kFlowSmem>>>(D, dt, pub, aptr, nullptr, ticket);
        cudaDeviceSynchronize();
    }
    std::vector<long long> st(tasks.size() * 4);
    cudaMemcpy(st.data(), D.ubuf, st.size() * 8, cudaMemcpyDeviceToHost);
    const long long t0 = st[0];
    size_t at = 0;
    for (uint32_t bj = 0; bj < B; bj++)
        for (uint32_t bi = bj; bi < B; bi++, at++)
            if (bi <= bj + 1)
                printf("tile (%2u,%2u) %s: start %7lld  earlier-blocks-done %7lld  last-block-done %7lld  finished %7lld (ns)\n", bi, bj, bi == bj ? "diag " : "panel",
                       st[at * 4] - t0, st[at * 4 + 1] - t0, st[at * 4 + 2] - t0, st[at * 4 + 3] - t0);
    {   // per-strip stamps of panel tile (1,0): strip entered, strip's values present, substitution starts, substitution done
        long long ps[32]; cudaMemcpy(ps, (long long*)D.ubuf + 4096, sizeof(ps), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 8; m++) printf("panel (1,0) strip %d: enter %7lld  values present %7lld  substitution %7lld .. %7lld (ns)\n", m, ps[m*4] - t0, ps[m*4+1] - t0, ps[m*4+2] - t0, ps[m*4+3] - t0);
    }
    {
        long long ds[24]; cudaMemcpy(ds, dst, sizeof(ds), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 3; m++) printf("diag tile (0,0) micro-panel %d: 8x8 LDL %lld, own row %lld + stores %lld, barrier %lld, rank-8 update %lld cycles; next starts +%lld\n", m, ds[m*8+1]-ds[m*8], ds[m*8+5]-ds[m*8+1], ds[m*8+2]-ds[m*8+5], ds[m*8+3]-ds[m*8+2], ds[m*8+4]-ds[m*8+3], m < 2 ? ds[(m+1)*8]-ds[m*8+4] : 0);
    }
    std::vector<double> L(A.size());
    cudaMemcpy(L.data(), D.pan, A.size() * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (uint32_t i = 0; i < f; i += 7) for (uint32_t j = 0; j <= i; j += 5) {
        double sacc = 0;
        for (uint32_t k = 0; k <= j; k++) { double lik = (i == k) ? 1.0 : L[(size_t)k * f + i], ljk = (j == k) ? 1.0 : L[(size_t)k * f + j]; sacc += lik * L[(size_t)k * f + k] * ljk; }
        maxerr = fmax(maxerr, fabs(sacc - A[(size_t)j * f + i]));
    }
    printf("max |L D L^T - A| (sampled) = %.3e\n", maxerr);
    return 0;
}
