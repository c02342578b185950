// Device pieces shared by pfc_rows.cu and pfc_peer.cu: the backward coefficients of one row (pfc_backward_prepare) and the
// loss reduction in loss_kernel's order, so that the kernels that fuse "statistics -> loss -> coefficients" into one launch
// produce the same bits as the separate launches.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfc {

struct PrepArgs {
    const float* grad_loss;     // device scalar d loss (null = 1)
    float s;
    int B, d;
    const int32_t* labels;      // shard-local target, -1 = another rank's
    const float* tgt_raw;
    int margin_kind;
    float cos_m, sin_m, theta;
    const __nv_bfloat16* xn;    // bf16, or fp16 bits when xn_f16 (the reference's AMP operands)
    int xn_f16;
    __nv_bfloat16* xs;          // always bf16: c_i can be far below the fp16 range
    float* coef;
    __nv_bfloat16* E;
    int n_pad;
};

// Backward coefficients of one row (nets/PartialFC.py:464-484 and the autograd of nets/ArcFace.py:80-91, :204):
//   c_i = g * s / (B * L_i);  Xs_i = c_i * Xn_i (bf16);  E'[i, y_i] = -dm_i * mask_i * Lothers_i
//   dm_i = d(margin)/dt = cos m + sin m * t / sqrt(1 - t^2)  if t > cos(pi - m) else 1   (CosFace: 1)
//   mask_i = 1 if -1 <= raw <= 1 (clamp backward) else 0
// Called by `vlanes` consecutive threads (vlane = 0 .. vlanes-1) that share the row.
__device__ __forceinline__ void prepare_row(const PrepArgs& a, int row, float L, float others, int vlane, int vlanes) {
    const float g = a.grad_loss ? a.grad_loss[0] : 1.f;
    const float c = g * a.s / (static_cast<float>(a.B) * L);
    const int nv = a.d >> 2;
    for (int k = vlane; k < nv; k += vlanes) {
        const uint2 raw = *reinterpret_cast<const uint2*>(a.xn + static_cast<size_t>(row) * a.d + 4 * k);
        float2 f0, f1;
        if (a.xn_f16) {
            f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        } else {
            f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        }
        __nv_bfloat162 q0 = __floats2bfloat162_rn(f0.x * c, f0.y * c), q1 = __floats2bfloat162_rn(f1.x * c, f1.y * c);
        uint2 out;
        out.x = *reinterpret_cast<uint32_t*>(&q0);
        out.y = *reinterpret_cast<uint32_t*>(&q1);
        *reinterpret_cast<uint2*>(a.xs + static_cast<size_t>(row) * a.d + 4 * k) = out;
    }
    if (vlane == 0) {
        a.coef[row] = c;
        const int lbl = a.labels[row];
        if (lbl >= 0) {
            const float raw = a.tgt_raw[row];
            const float mask = (fabsf(raw) <= 1.f) ? 1.f : 0.f;
            const float t = fminf(fmaxf(raw, -1.f), 1.f);
            float dm = 1.f;
            if (a.margin_kind == 0 && t > a.theta) dm = a.cos_m + a.sin_m * t / sqrtf(fmaxf(1.f - t * t, 1e-12f));
            // class-blocked spill: E'[class / 64][row][class % 64]
            a.E[(static_cast<size_t>(lbl >> 6) * a.B + row) * 64 + (lbl & 63)] = __float2bfloat16_rn(-dm * mask * others);
        }
    }
}

__device__ __forceinline__ float warp_sum_l(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// loss = -mean_i log(max(p_i, 1e-30)), p_i = target e / row sum (nets/PartialFC.py:454-461), from the finished [B][2]
// statistics, run by ONE CTA of kThreads threads (a divisor of 1024) standing in for loss_kernel's 1024: thread t covers the
// rows of loss_kernel's threads t, t + kThreads, ... in the same order, and the partial sums are combined in loss_kernel's
// order (warp tree, then 32 warp sums) -> the same bits.  Also writes row_L.  All threads of the CTA must call it.
template <int kThreads>
__device__ __forceinline__ void loss_from_stats(const float* stats, int B, float* __restrict__ row_L,
                                                float* __restrict__ loss) {
    __shared__ float wsum[32];
    for (int v = 0; v < 1024 / kThreads; ++v) {
        const int vt = threadIdx.x + kThreads * v;       // virtual thread id of loss_kernel
        float acc = 0.f;
        for (int i = vt; i < B; i += 1024) {
            const float others = __ldcg(stats + 2 * i), te = __ldcg(stats + 2 * i + 1);
            const float L = others + te;
            row_L[i] = L;
            acc -= logf(fmaxf(te / L, 1e-30f));
        }
        acc = warp_sum_l(acc);
        if ((threadIdx.x & 31) == 0) wsum[vt >> 5] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const float v = warp_sum_l(wsum[threadIdx.x]);
        if (threadIdx.x == 0) loss[0] = v / static_cast<float>(B);
    }
}

}  // namespace pfc
