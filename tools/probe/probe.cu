// EXPERIMENT helper (not part of the product library): do CTAs of a small register-only kernel become resident on an SM
// that already runs one persistent GEMM CTA?  stamp_kernel records %globaltimer; probe_kernel records, per CTA, the SM it
// ran on and when it started / finished (it spins `spin_ns` so that placement, not duration, is what is measured).
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}

__global__ void stamp_kernel(unsigned long long* out) { *out = gtime(); }

template <int REGS_PAD>
__global__ void __launch_bounds__(128) probe_kernel(unsigned long long* rec, int spin_ns, float* sink) {
    // REGS_PAD live floats keep the register allocation of the kernel at a chosen size
    float v[REGS_PAD];
#pragma unroll
    for (int i = 0; i < REGS_PAD; ++i) v[i] = threadIdx.x * 0.5f + i;
    const unsigned long long t0 = gtime();
    unsigned long long t1 = t0;
    while (t1 - t0 < (unsigned long long)spin_ns) {
#pragma unroll
        for (int i = 0; i < REGS_PAD; ++i) v[i] = v[i] * 1.0001f + 0.5f;
        t1 = gtime();
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < REGS_PAD; ++i) s += v[i];
    if (s == 12345.678f) *sink = s;
    if (threadIdx.x == 0) {
        rec[3 * blockIdx.x + 0] = smid();
        rec[3 * blockIdx.x + 1] = t0;
        rec[3 * blockIdx.x + 2] = t1;
    }
}

extern "C" {
int probe_stamp(unsigned long long* out, void* stream) {
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out);
    return (int)cudaGetLastError();
}
// regs: 0 -> 18 registers, 1 -> 47, 3 -> ~63, 2 -> 79
int probe_launch(unsigned long long* rec, int ctas, int spin_ns, int regs, float* sink, void* stream) {
    if (regs == 0) probe_kernel<8><<<ctas, 128, 0, (cudaStream_t)stream>>>(rec, spin_ns, sink);
    else if (regs == 1) probe_kernel<40><<<ctas, 128, 0, (cudaStream_t)stream>>>(rec, spin_ns, sink);
    else if (regs == 3) probe_kernel<56><<<ctas, 128, 0, (cudaStream_t)stream>>>(rec, spin_ns, sink);
    else probe_kernel<72><<<ctas, 128, 0, (cudaStream_t)stream>>>(rec, spin_ns, sink);
    return (int)cudaGetLastError();
}
}
