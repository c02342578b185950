// Phase timing of the one-launch sampler (globaltimer stamps of CTA 0 / thread 0).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPFC_SAMPLE_STAMPS -I include \
//        -I face-recognition-pytorch_b200/csrc tools/probe/sample_phases.cu -o /tmp/sample_phases && /tmp/sample_phases
#include "../../face-recognition-pytorch_b200/csrc/pfc_sample.cu"
#include <stdio.h>
#include <vector>

int main() {
    const int cases[][3] = {{45029, 4502, 1024}, {257489, 51497, 4096}, {2000000, 400000, 4096}};
    const char* names[] = {"start", "stage+bitmap", "hist0", "sync", "merge0", "hist1+sync", "merge1", "hist2+sync", "merge2",
                           "count+sync", "compact", "sync", "remap"};
    for (int mode : {8, 16})
    for (auto& c : cases) {
        pfc_sample_debug_cluster(mode);
        const int nl = c[0], ns = c[1], B = c[2];
        std::vector<float> perm(nl);
        std::vector<int32_t> lab(B);
        uint32_t x = 12345;
        for (auto& v : perm) { x = x * 1664525u + 1013904223u; v = (x >> 8) * (1.0f / 16777216.0f); }
        for (auto& v : lab) { x = x * 1664525u + 1013904223u; v = (x >> 4) % (8u * nl) < (uint32_t)nl ? (x >> 4) % nl : -1; }
        float* dperm; int32_t *dlab, *dn, *drem; int64_t* didx; void* ws;
        cudaMalloc(&dperm, nl * 4); cudaMalloc(&dlab, B * 4); cudaMalloc(&dn, 4); cudaMalloc(&drem, B * 4);
        cudaMalloc(&didx, sizeof(int64_t) * (ns > B ? ns : B)); cudaMalloc(&ws, pfc_sample_workspace_bytes(nl));
        cudaMemcpy(dperm, perm.data(), nl * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dlab, lab.data(), B * 4, cudaMemcpyHostToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float ms = 0;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            int rc = pfc_sample(dperm, dlab, B, nl, ns, didx, dn, drem, ws, pfc_sample_workspace_bytes(nl), 0);
            cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            if (rc) printf("rc=%d\n", rc);
        }
        unsigned long long st[16];
        cudaMemcpyFromSymbol(st, pfc::g_stamps, sizeof(st));
        int n; cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
        printf("cluster %d launches %d nl=%d k=%d B=%d: n_out=%d, event time %.1f us; phases (us):", mode, pfc_sample_launches(nl), nl, ns, B, n, ms * 1e3);
        for (int i = 1; i <= 12; ++i) printf(" %s %.1f", names[i], (st[i] - st[i - 1]) * 1e-3);
        printf(" | total %.1f (%s)\n", (st[12] - st[0]) * 1e-3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
