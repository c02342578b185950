// HBM-bound row kernels of the head: fused L2-normalise (fp32 -> bf16 + 1/norm), label localisation,
// row-statistics / loss reduction, backward coefficient + target patch, normalise-backward for dX and dW,
// fused SGD / AdamW update that also emits the next step's normalised bf16 shard, and row gather / scatter
// for the sampled (PartialFC r<1) path.
//
// Layout: every matrix is row-major with d (embedding size) contiguous; one warp owns one row and moves it
// with 16-byte loads (lane l handles float4 #l, #l+32, ...), reductions are warp shuffles.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include "pfc_internal.h"
#include "pfc_launch.cuh"
#include "pfc_prepare.cuh"

namespace pfc {

constexpr int ROW_WARPS = 8;           // rows per CTA
constexpr int MAXV = 8;                // float4 per lane -> d <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4_stream(const float* p) {   // read-once data: do not keep in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 pack4_bf16(float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    return r;
}
// the GEMM operands are bf16 by default and fp16 in the reference's AMP mode (conf.mixed_precision, nets/PartialFC.py:198)
__device__ __forceinline__ uint2 pack4_op(float4 v, int f16) {
    if (!f16) return pack4_bf16(v);
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    return r;
}
// bf16 twin of an fp16 operand row: rounds through fp16 first, i.e. bit for bit what cast_f16_bf16_kernel makes of the
// fp16 row that pack4_op(v, 1) writes
__device__ __forceinline__ uint2 pack4_bf16_of_f16(float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return pack4_bf16(make_float4(fa.x, fa.y, fb.x, fb.y));
}
__device__ __forceinline__ float4 unpack4_bf16(uint2 r) {
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// ---------------------------------------------------------------------------------------------------------
// xn = x / max(||x||, 1e-12) as bf16, inv = 1 / max(||x||, 1e-12)          (nets/PartialFC.py:199-200)
// optional gather: row r reads x[index[r]]                                   (nets/PartialFC.py:120)
// optional label localisation of the same rows (single-GPU step: one launch fewer), nets/PartialFC.py:188-193
__global__ void __launch_bounds__(ROW_WARPS * 32)
l2norm_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ index, int rows, int d,
                   __nv_bfloat16* __restrict__ xn, float* __restrict__ inv_norm,
                   const int64_t* __restrict__ labels, int64_t class_start, int num_local,
                   int32_t* __restrict__ labels_local, int f16) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (labels != nullptr && lane == 0) {
        const int64_t l = labels[row] - class_start;
        labels_local[row] = (l >= 0 && l < num_local) ? static_cast<int32_t>(l) : -1;
    }
    const int64_t src = index ? index[row] : row;
    const float* xr = x + src * d;
    const int nv = d >> 2;
    float4 v[MAXV];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int c = lane + 32 * j;
        if (c < nv) {
            v[j] = ld4_stream(xr + 4 * c);
            ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
        }
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    __nv_bfloat16* o = xn + static_cast<size_t>(row) * d;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int c = lane + 32 * j;
        if (c < nv) {
            const float4 q = make_float4(v[j].x / denom, v[j].y / denom, v[j].z / denom, v[j].w / denom);
            *reinterpret_cast<uint2*>(o + 4 * c) = pack4_op(q, f16);
        }
    }
    if (lane == 0) inv_norm[row] = 1.f / denom;
}

// fp16 -> bf16, 8 elements per thread: the dX contraction needs the shard in the spill's format (AMP mode only)
__global__ void __launch_bounds__(256)
cast_f16_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n8) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 v = src[i];
    uint4 o;
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&in[k]));
        __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
        out[k] = *reinterpret_cast<uint32_t*>(&b);
    }
    o.x = out[0]; o.y = out[1]; o.z = out[2]; o.w = out[3];
    dst[i] = o;
}

// labels -> shard-local ids, -1 for classes owned by another rank           (nets/PartialFC.py:188-193)
__global__ void localize_labels_kernel(const int64_t* __restrict__ labels, int B, int64_t class_start,
                                       int num_local, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int64_t l = labels[i] - class_start;
    out[i] = (l >= 0 && l < num_local) ? static_cast<int32_t>(l) : -1;
}

// stats[i] = { sum over the part_sum slabs of part_sum[t][i],  target e (0 when the target lives on another rank) }
// Fixed summation order -> bit-reproducible.  CTA = 8 rows x 32 slab groups (a first version with 32 rows per CTA had
// only B/32 CTAs in flight and took 14 us for 3 MB).
constexpr int RS_ROWS = 8, RS_GROUPS = 64;
__device__ __forceinline__ float row_stats_sum(const float* __restrict__ part_sum, int n_tiles, int B, int B_pad,
                                               float (*red)[RS_ROWS + 1]) {
    const int r = threadIdx.x & (RS_ROWS - 1), g = threadIdx.x / RS_ROWS;
    const int row = blockIdx.x * RS_ROWS + r;
    float s = 0.f;
    if (row < B) {
        int t = g;
        for (; t + 3 * RS_GROUPS < n_tiles; t += 4 * RS_GROUPS) {        // four independent loads in flight
            const float a = part_sum[static_cast<size_t>(t) * B_pad + row];
            const float b = part_sum[static_cast<size_t>(t + RS_GROUPS) * B_pad + row];
            const float c = part_sum[static_cast<size_t>(t + 2 * RS_GROUPS) * B_pad + row];
            const float d = part_sum[static_cast<size_t>(t + 3 * RS_GROUPS) * B_pad + row];
            s += a; s += b; s += c; s += d;
        }
        for (; t < n_tiles; t += RS_GROUPS) s += part_sum[static_cast<size_t>(t) * B_pad + row];
    }
    red[g][r] = s;
    __syncthreads();
    float tot = 0.f;
    if (g == 0) {
#pragma unroll
        for (int k = 0; k < RS_GROUPS; ++k) tot += red[k][r];
    }
    return tot;      // valid for g == 0
}

__global__ void __launch_bounds__(RS_ROWS * RS_GROUPS)
row_stats_kernel(const float* __restrict__ part_sum, int n_tiles, int B, int B_pad,
                 const int32_t* __restrict__ labels, const float* __restrict__ tgt_e, float* __restrict__ stats) {
    __shared__ float red[RS_GROUPS][RS_ROWS + 1];
    const float tot = row_stats_sum(part_sum, n_tiles, B, B_pad, red);
    const int row = blockIdx.x * RS_ROWS + (threadIdx.x & (RS_ROWS - 1));
    if (threadIdx.x < RS_ROWS && row < B) {
        stats[2 * row] = tot;
        stats[2 * row + 1] = (labels[row] >= 0) ? tgt_e[row] : 0.f;
    }
}

// Single-GPU step: row statistics AND the loss in one launch.  Every CTA writes the stats of its rows, then takes a
// ticket; the last CTA through sees all of them (threadfence + L2 loads) and forms row_L and the loss exactly like
// loss_kernel (same per-thread row assignment and reduction tree -> same bits as the two-kernel path).
// kPrepare: the CTA also forms the backward coefficients of its 8 rows (pfc_backward_prepare: on one GPU the row sum is
// final as soon as the CTA has it) -- two warps per row -- so the no-autograd step has one launch fewer.
template <bool kPrepare>
__global__ void __launch_bounds__(RS_ROWS * RS_GROUPS)
row_stats_loss_kernel(const float* __restrict__ part_sum, int n_tiles, int B, int B_pad,
                      const int32_t* __restrict__ labels, const float* __restrict__ tgt_e, float* stats,
                      float* __restrict__ row_L, float* __restrict__ loss, unsigned int* ticket, PrepArgs pa) {
    __shared__ float red[RS_GROUPS][RS_ROWS + 1];
    __shared__ float row_o[RS_ROWS], row_t[RS_ROWS];
    __shared__ bool last;
    const float tot = row_stats_sum(part_sum, n_tiles, B, B_pad, red);
    const int row = blockIdx.x * RS_ROWS + (threadIdx.x & (RS_ROWS - 1));
    if (threadIdx.x < RS_ROWS && row < B) {
        const float te = (labels[row] >= 0) ? tgt_e[row] : 0.f;
        stats[2 * row] = tot;
        stats[2 * row + 1] = te;
        if (kPrepare) { row_o[threadIdx.x] = tot; row_t[threadIdx.x] = te; }
    }
    if (kPrepare) {
        __syncthreads();
        constexpr int kWarpsPerRow = (RS_ROWS * RS_GROUPS / 32) / RS_ROWS;    // 16 warps, 8 rows: two warps per row
        const int w = threadIdx.x >> 5, r = w / kWarpsPerRow;
        const int prow = blockIdx.x * RS_ROWS + r;
        if (prow < B) {
            const float L = row_o[r] + row_t[r];
            row_L[prow] = L;
            prepare_row(pa, prow, L, row_o[r], (w % kWarpsPerRow) * 32 + (threadIdx.x & 31), kWarpsPerRow * 32);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    loss_from_stats<RS_ROWS * RS_GROUPS>(stats, B, row_L, loss);
    if (threadIdx.x == 0) *ticket = 0;
}

// loss = -mean_i log(max(p_i, 1e-30)),  p_i = target e / row sum             (nets/PartialFC.py:454-461)
// stats is the (all-reduced) [B][2] array; also emits L_i = stats[i][0] + stats[i][1].
__global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ stats, int B, float* __restrict__ row_L, float* __restrict__ loss) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) {
        const float others = stats[2 * i], te = stats[2 * i + 1];
        const float L = others + te;
        row_L[i] = L;
        const float p = fmaxf(te / L, 1e-30f);
        acc -= logf(p);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = red[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = v / static_cast<float>(B);
    }
}

// Backward coefficients of every row: prepare_row (pfc_prepare.cuh), one warp per row.
__global__ void __launch_bounds__(ROW_WARPS * 32)
backward_prepare_kernel(const float* __restrict__ stats, const float* __restrict__ row_L, PrepArgs pa) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= pa.B) return;
    prepare_row(pa, row, row_L[row], stats[2 * row], threadIdx.x & 31, 32);
}

// d = 512 fast path of dx_finalize_kernel: FOUR warps per row (one float4 per lane and slab), two rows per CTA -- the
// one-warp-per-row kernel below has only `rows` warps in flight (7 per SM at B = 1024) and ran at 1.9 TB/s.
__global__ void __launch_bounds__(256)
dx_finalize_d512_kernel(const float* __restrict__ partial, int splits, size_t split_stride, const float* __restrict__ coef,
                        const float* __restrict__ x, const float* __restrict__ inv_norm, float scale, int rows,
                        float* __restrict__ out) {
    __shared__ float part[2][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = warp >> 2, q = warp & 3;
    const int row = blockIdx.x * 2 + r;
    const bool ok = row < rows;
    const size_t off = static_cast<size_t>(ok ? row : 0) * 512 + q * 128 + lane * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), xv = a;
    float dot = 0.f, inv = 0.f;
    if (ok) {
        const float* pp = partial + off;
        for (int z = 0; z < splits; z += 8) {                // eight split slabs in flight, summed in slab order
            float4 p[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                p[u] = (z + u < splits) ? ld4_stream(pp + (z + u) * split_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { a.x += p[u].x; a.y += p[u].y; a.z += p[u].z; a.w += p[u].w; }
        }
        const float c = coef ? coef[row] : 1.f;
        a.x *= c; a.y *= c; a.z *= c; a.w *= c;
        if (x) {
            inv = inv_norm[row];
            xv = ld4(x + off);
            xv.x *= inv; xv.y *= inv; xv.z *= inv; xv.w *= inv;
            dot = xv.x * a.x + xv.y * a.y + xv.z * a.z + xv.w * a.w;
        }
    }
    if (x) {                                                 // uniform: row dot = sum of the four quarter sums
        dot = warp_sum(dot);
        if (lane == 0) part[r][q] = dot;
        __syncthreads();
        dot = (part[r][0] + part[r][1]) + (part[r][2] + part[r][3]);
    }
    if (!ok) return;
    const float m = x ? scale * inv : scale;
    if (x) { a.x -= xv.x * dot; a.y -= xv.y * dot; a.z -= xv.z * dot; a.w -= xv.w * dot; }
    a.x *= m; a.y *= m; a.z *= m; a.w *= m;
    st4(out + off, a);
}

// dX: sum the class-split partials, scale by c_i, and (when x is given) apply the normalise backward
//   dx = scale * (dxn - xn (xn . dxn)) / ||x||,   xn = x * inv_norm          (autograd of F.normalize)
// scale = world_size reproduces AllGatherFunc.backward's "grad_out *= len(grad_list)" (nets/PartialFC.py:521).
__global__ void __launch_bounds__(ROW_WARPS * 32)
dx_finalize_kernel(const float* __restrict__ partial, int splits, size_t split_stride, const float* __restrict__ coef,
                   const float* __restrict__ x, const float* __restrict__ inv_norm, float scale, int rows, int d,
                   float* __restrict__ out) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = d >> 2;
    const float c = coef ? coef[row] : 1.f;
    float4 g[MAXV], xv[MAXV];
    float dot = 0.f;
    const float inv = x ? inv_norm[row] : 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* pp = partial + static_cast<size_t>(row) * d + 4 * k;
            for (int z = 0; z < splits; z += 4) {            // four split slabs in flight, summed in slab order
                float4 p[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    p[u] = (z + u < splits) ? ld4_stream(pp + (z + u) * split_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) { a.x += p[u].x; a.y += p[u].y; a.z += p[u].z; a.w += p[u].w; }
            }
            a.x *= c; a.y *= c; a.z *= c; a.w *= c;
            g[j] = a;
            if (x) {
                float4 q = ld4(x + static_cast<size_t>(row) * d + 4 * k);
                q.x *= inv; q.y *= inv; q.z *= inv; q.w *= inv;
                xv[j] = q;
                dot += q.x * a.x + q.y * a.y + q.z * a.z + q.w * a.w;
            }
        }
    }
    if (x) dot = warp_sum(dot);
    const float m = x ? scale * inv : scale;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            float4 a = g[j];
            if (x) {
                a.x -= xv[j].x * dot; a.y -= xv[j].y * dot; a.z -= xv[j].z * dot; a.w -= xv[j].w * dot;
            }
            a.x *= m; a.y *= m; a.z *= m; a.w *= m;
            st4(out + static_cast<size_t>(row) * d + 4 * k, a);
        }
    }
}

// dW normalise-backward (+ optional fused optimiser step, + next step's bf16 normalised shard).
//   dw  = (dwn - wn (wn . dwn)) / ||w||
//   SGD   (torch.optim.SGD, dampening 0, no nesterov):  g = dw + wd w;  buf = mu buf + g;  w -= lr buf
//   AdamW (torch.optim.AdamW):  w *= 1 - lr wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//                               w -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
enum { OPT_NONE = 0, OPT_SGD = 1, OPT_ADAMW = 2, OPT_ADAM = 3 };
struct OptArgs {
    int kind;
    float lr, momentum, wd;         // SGD
    float beta1, beta2, eps, bc1, bc2_sqrt;   // Adam(W): bc1 = 1-b1^t, bc2_sqrt = sqrt(1-b2^t)
    const float* grad_scale;        // device scalar: the loss scale the gradient carries (divided out first), or null
    const int* step_dev;            // Adam(W): device step counter (the update is step step_dev[0] + 1), or null
    const int64_t* index;           // sampled shards: row r of (dwn, inv_norm_w, wn_next) is row index[r] of (w, state), or null
    int wn_f16;                     // wn_next is written as fp16 instead of bf16
    __nv_bfloat16* wn_copy_b;       // wn_f16 only: bf16 twin of wn_next for the next step's dX contraction, or null
};

__device__ __forceinline__ float opt_inv_grad_scale(const float* grad_scale) {
    return grad_scale ? 1.f / grad_scale[0] : 1.f;
}

__global__ void __launch_bounds__(ROW_WARPS * 32)
dw_finalize_kernel(const float* __restrict__ dwn, float* __restrict__ w, const float* inv_norm_w,
                   int rows, int d, OptArgs opt, float* __restrict__ dw_out, float* __restrict__ st1,
                   float* __restrict__ st2, __nv_bfloat16* __restrict__ wn_next, float* inv_norm_next) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = d >> 2;
    const float inv = inv_norm_w[row];
    const size_t base = static_cast<size_t>(row) * d;
    // in-place update of a sampled shard: the weights / optimizer state live in the full [num_local, d] arrays
    const size_t wbase = opt.index ? static_cast<size_t>(opt.index[row]) * d : base;
    float4 g[MAXV], wv[MAXV];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            g[j] = ld4_stream(dwn + base + 4 * k);
            wv[j] = ld4(w + wbase + 4 * k);
            dot += wv[j].x * g[j].x + wv[j].y * g[j].y + wv[j].z * g[j].z + wv[j].w * g[j].w;
        }
    }
    dot = warp_sum(dot) * inv;                       // wn . dwn
    const float wscale = dot * inv;                  // (wn . dwn) * wn = w * (dot * inv)
    const float gs = inv * opt_inv_grad_scale(opt.grad_scale);
    float bc1 = opt.bc1, bc2_sqrt = opt.bc2_sqrt;
    if (opt.step_dev != nullptr) {                   // CUDA-graph replay: the step count lives on the device
        // in double, like the host computes the same two numbers for an eager step (pfc_dw_adam)
        const double t = static_cast<double>(opt.step_dev[0] + 1);
        bc1 = static_cast<float>(1.0 - pow(static_cast<double>(opt.beta1), t));
        bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(opt.beta2), t)));
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            float4 a;
            a.x = (g[j].x - wv[j].x * wscale) * gs;
            a.y = (g[j].y - wv[j].y * wscale) * gs;
            a.z = (g[j].z - wv[j].z * wscale) * gs;
            a.w = (g[j].w - wv[j].w * wscale) * gs;
            if (opt.kind == OPT_NONE) {
                st4(dw_out + base + 4 * k, a);
                continue;
            }
            float4 wq = wv[j];
            if (opt.kind == OPT_SGD) {
                a.x += opt.wd * wq.x; a.y += opt.wd * wq.y; a.z += opt.wd * wq.z; a.w += opt.wd * wq.w;
                float4 b = a;
                if (opt.momentum != 0.f) {
                    b = ld4(st1 + wbase + 4 * k);
                    b.x = opt.momentum * b.x + a.x; b.y = opt.momentum * b.y + a.y;
                    b.z = opt.momentum * b.z + a.z; b.w = opt.momentum * b.w + a.w;
                    st4(st1 + wbase + 4 * k, b);
                }
                wq.x -= opt.lr * b.x; wq.y -= opt.lr * b.y; wq.z -= opt.lr * b.z; wq.w -= opt.lr * b.w;
            } else {
                if (opt.kind == OPT_ADAMW) {
                    const float dec = 1.f - opt.lr * opt.wd;
                    wq.x *= dec; wq.y *= dec; wq.z *= dec; wq.w *= dec;
                } else {
                    a.x += opt.wd * wq.x; a.y += opt.wd * wq.y; a.z += opt.wd * wq.z; a.w += opt.wd * wq.w;
                }
                float4 m = ld4(st1 + wbase + 4 * k), v = ld4(st2 + wbase + 4 * k);
                const float b1 = opt.beta1, b2 = opt.beta2;
                m.x = b1 * m.x + (1.f - b1) * a.x; m.y = b1 * m.y + (1.f - b1) * a.y;
                m.z = b1 * m.z + (1.f - b1) * a.z; m.w = b1 * m.w + (1.f - b1) * a.w;
                v.x = b2 * v.x + (1.f - b2) * a.x * a.x; v.y = b2 * v.y + (1.f - b2) * a.y * a.y;
                v.z = b2 * v.z + (1.f - b2) * a.z * a.z; v.w = b2 * v.w + (1.f - b2) * a.w * a.w;
                st4(st1 + wbase + 4 * k, m);
                st4(st2 + wbase + 4 * k, v);
                const float step = opt.lr / bc1;
                wq.x -= step * m.x / (sqrtf(v.x) / bc2_sqrt + opt.eps);
                wq.y -= step * m.y / (sqrtf(v.y) / bc2_sqrt + opt.eps);
                wq.z -= step * m.z / (sqrtf(v.z) / bc2_sqrt + opt.eps);
                wq.w -= step * m.w / (sqrtf(v.w) / bc2_sqrt + opt.eps);
            }
            st4(w + wbase + 4 * k, wq);
            wv[j] = wq;
            ss += wq.x * wq.x + wq.y * wq.y + wq.z * wq.z + wq.w * wq.w;
        }
    }
    if (opt.kind == OPT_NONE || wn_next == nullptr) return;
    // next step's normalised operand straight from the registers: saves re-reading the fp32 shard
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            const float4 q = make_float4(wv[j].x / denom, wv[j].y / denom, wv[j].z / denom, wv[j].w / denom);
            *reinterpret_cast<uint2*>(wn_next + base + 4 * k) = pack4_op(q, opt.wn_f16);
            if (opt.wn_copy_b != nullptr)
                *reinterpret_cast<uint2*>(opt.wn_copy_b + base + 4 * k) = pack4_bf16_of_f16(q);
        }
    }
    if (lane == 0) inv_norm_next[row] = 1.f / denom;
}

// Specialised fused SGD row kernel for d = 128 * NV: all three input streams (dwn, w, momentum) are requested
// up front (12 x 16-byte loads per lane in flight at d = 512) so a warp is never waiting on a dependent phase,
// arrays are sized exactly (no predicated dead registers), 4 rows per CTA for occupancy.
// kL2 (PFC_L2_GRAD, bf16 gradient only): the gradient rows are expected in L2 (the dW GEMM stored them evict_last), the
// fp32 state streams through with evict_first so that it does not push them out, and each gradient row is DISCARDED
// from L2 once its warp has consumed it -- its dirty lines never reach HBM (the next step rewrites the buffer).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4_hint(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

template <int NV, bool kGradBf16, bool kL2 = false>
__global__ void __launch_bounds__(128)
dw_sgd_rows_kernel(const void* __restrict__ dwn_, float* __restrict__ w, float* __restrict__ mom,
                   const float* inv_norm_w, int rows, float lr, float momentum, float wd,
                   const float* __restrict__ grad_scale, __nv_bfloat16* __restrict__ wn_next, float* inv_norm_next,
                   const int64_t* __restrict__ index, int wn_f16, __nv_bfloat16* __restrict__ wn_copy_b) {
    const float inv_grad_scale = opt_inv_grad_scale(grad_scale);
    constexpr int d = 128 * NV;
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    // grid-stride over rows: with one row per warp and a full grid this is a single trip; the persistent launch
    // (a few CTAs per SM, see pfc_debug_sgd_persistent) loops so that the kernel can share SMs with a GEMM
    for (int row = blockIdx.x * warps + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps) {
    const size_t base = static_cast<size_t>(row) * d;
    // sampled shard updated in place: w / momentum rows live at index[row] of the full arrays
    const size_t wbase = index ? static_cast<size_t>(index[row]) * d : base;
    float4 g[NV], wv[NV], mv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        if constexpr (kGradBf16) {
            uint2 r;
            const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(dwn_) + base + 4 * (lane + 32 * j);
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(gp));
            g[j] = unpack4_bf16(r);
        } else {
            g[j] = ld4_stream(static_cast<const float*>(dwn_) + base + 4 * (lane + 32 * j));
        }
    }
    uint64_t pol = 0;
    if constexpr (kL2) pol = l2_evict_first_policy();
#pragma unroll
    for (int j = 0; j < NV; ++j)
        wv[j] = kL2 ? ld4_hint(w + wbase + 4 * (lane + 32 * j), pol) : ld4(w + wbase + 4 * (lane + 32 * j));
#pragma unroll
    for (int j = 0; j < NV; ++j)
        mv[j] = kL2 ? ld4_hint(mom + wbase + 4 * (lane + 32 * j), pol) : ld4(mom + wbase + 4 * (lane + 32 * j));
    const float inv = inv_norm_w[row];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
        dot += wv[j].x * g[j].x + wv[j].y * g[j].y + wv[j].z * g[j].z + wv[j].w * g[j].w;
    dot = warp_sum(dot) * inv;
    const float wscale = dot * inv;
    const float gs = inv * inv_grad_scale;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        float4 a, b = mv[j], q = wv[j];
        a.x = (g[j].x - q.x * wscale) * gs + wd * q.x;
        a.y = (g[j].y - q.y * wscale) * gs + wd * q.y;
        a.z = (g[j].z - q.z * wscale) * gs + wd * q.z;
        a.w = (g[j].w - q.w * wscale) * gs + wd * q.w;
        b.x = momentum * b.x + a.x; b.y = momentum * b.y + a.y;
        b.z = momentum * b.z + a.z; b.w = momentum * b.w + a.w;
        q.x -= lr * b.x; q.y -= lr * b.y; q.z -= lr * b.z; q.w -= lr * b.w;
        if constexpr (kL2) {
            st4_hint(mom + wbase + 4 * (lane + 32 * j), b, pol);
            st4_hint(w + wbase + 4 * (lane + 32 * j), q, pol);
        } else {
            st4(mom + wbase + 4 * (lane + 32 * j), b);
            st4(w + wbase + 4 * (lane + 32 * j), q);
        }
        wv[j] = q;
        ss += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    }
    if constexpr (kL2 && kGradBf16) {
        // every lane's gradient loads have returned (their values went into the update above): drop the row's lines
        __syncwarp();
        constexpr int kLines = d * 2 / 128;
        if (lane < kLines)
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(static_cast<const __nv_bfloat16*>(dwn_) + base + lane * 64)
                         : "memory");
    }
    if (wn_next == nullptr) continue;
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float4 q = make_float4(wv[j].x / denom, wv[j].y / denom, wv[j].z / denom, wv[j].w / denom);
        *reinterpret_cast<uint2*>(wn_next + base + 4 * (lane + 32 * j)) = pack4_op(q, wn_f16);
        if (wn_copy_b != nullptr)      // AMP mode: the dX contraction reads the shard as bf16 (see pfc_backward_dx)
            *reinterpret_cast<uint2*>(wn_copy_b + base + 4 * (lane + 32 * j)) = pack4_bf16_of_f16(q);
    }
    if (lane == 0) inv_norm_next[row] = 1.f / denom;
    }
}

static int g_sgd_persistent_warps = 0;   // 0: one row per warp, full grid; > 0: that many warps per SM, grid-stride

template <int NV>
static void launch_dw_sgd_rows(const void* dwn, bool bf16, float* w, float* mom, const float* inv_norm_w, int rows,
                               float lr, float momentum, float wd, const float* igs, __nv_bfloat16* wn_next,
                               float* inv_next, const int64_t* index, int wn_f16, __nv_bfloat16* wn_copy_b,
                               cudaStream_t st) {
    int grid = (rows + 3) / 4, block = 128;
    if (g_sgd_persistent_warps > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        block = 32 * g_sgd_persistent_warps;
        if (block > 128) block = 128;                 // __launch_bounds__(128): several CTAs per SM instead
        const int ctas_per_sm = (32 * g_sgd_persistent_warps + block - 1) / block;
        grid = sms * ctas_per_sm;
        const int need = (rows + block / 32 - 1) / (block / 32);
        if (grid > need) grid = need;
    }
    if (bf16 && pfc_l2_grad_enabled())
        launch_step_kernel(PDL_UPDATE, dw_sgd_rows_kernel<NV, true, true>, grid, block, 0, st, dwn, w, mom, inv_norm_w, rows, lr,
                           momentum, wd, igs, wn_next, inv_next, index, wn_f16, wn_copy_b);
    else if (bf16)
        launch_step_kernel(PDL_UPDATE, dw_sgd_rows_kernel<NV, true>, grid, block, 0, st, dwn, w, mom, inv_norm_w, rows, lr, momentum,
                           wd, igs, wn_next, inv_next, index, wn_f16, wn_copy_b);
    else
        launch_step_kernel(PDL_UPDATE, dw_sgd_rows_kernel<NV, false>, grid, block, 0, st, dwn, w, mom, inv_norm_w, rows, lr, momentum,
                           wd, igs, wn_next, inv_next, index, wn_f16, wn_copy_b);
}

// dst[r] = src[index[r]]  /  dst[index[r]] = src[r]   (nets/PartialFC.py:120-121, :142-143), up to 3 tensors at once
struct RowSet { const float* src[3]; float* dst[3]; int count; };

template <bool kScatter>
__global__ void __launch_bounds__(ROW_WARPS * 32)
move_rows_kernel(RowSet set, const int64_t* __restrict__ index, int rows, int d) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t other = index[row];
    const int nv = d >> 2;
    for (int t = 0; t < set.count; ++t) {
        const float* s = set.src[t] + (kScatter ? static_cast<int64_t>(row) : other) * d;
        float* o = set.dst[t] + (kScatter ? other : static_cast<int64_t>(row)) * d;
        for (int k = lane; k < nv; k += 32) st4(o + 4 * k, ld4_stream(s + 4 * k));
    }
}

// Stand-alone margin module (nets/ArcFace.py:76-91, :100-105, :27-61): out = s * margin(logits), gate = d out / d in.
__global__ void margin_apply_kernel(const float* __restrict__ in, const int64_t* __restrict__ labels, int B, int n,
                                    int margin_kind, float s, float cos_m, float sin_m, float theta, float sinmm,
                                    float m3, float thr, float* __restrict__ out, float* __restrict__ gate) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<size_t>(B) * n) return;
    const int i = static_cast<int>(idx / n), c = static_cast<int>(idx - static_cast<size_t>(i) * n);
    const float t = in[idx];
    float v = t, g = 1.f;
    if (labels[i] == c) {
        if (margin_kind == 0) {
            if (t > theta) {
                const float st = sqrtf(1.f - t * t);
                v = t * cos_m - st * sin_m;
                g = cos_m + sin_m * t / st;
            } else {
                v = t - sinmm;
            }
        } else {
            v = t - m3;
        }
    } else if (thr > 0.f && t > thr) {
        v = 0.f;
        g = 0.f;
    }
    out[idx] = v * s;
    if (gate) gate[idx] = g * s;
}

static inline int row_grid(int rows) { return (rows + ROW_WARPS - 1) / ROW_WARPS; }
static inline int check_launch() { return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH; }
static inline bool bad_d(int d) { return d <= 0 || (d & 7) || d > 128 * MAXV; }

}  // namespace pfc

using namespace pfc;

extern "C" {

// not part of the public header: > 0 runs the fused SGD row kernel as a persistent grid with that many warps per SM
// (used when it is overlapped with a GEMM on another stream), 0 restores the full grid
void pfc_debug_sgd_persistent(int warps_per_sm) { g_sgd_persistent_warps = warps_per_sm; }

int pfc_l2norm_rows(const float* x, const int64_t* index, int rows, int d, void* xn, float* inv_norm, int fp16_operands,
                    void* stream) {
    if (rows < 0 || bad_d(d)) return PFC_ERR_SHAPE;
    if (rows == 0) return PFC_OK;
    launch_step_kernel(PDL_NORMALISE, l2norm_rows_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        x, index, rows, d, reinterpret_cast<__nv_bfloat16*>(xn), inv_norm, nullptr, 0, 0, nullptr, fp16_operands);
    return check_launch();
}

int pfc_l2norm_rows_localize(const float* x, int rows, int d, void* xn, float* inv_norm, const int64_t* labels,
                             int64_t class_start, int num_local, int32_t* labels_local, int fp16_operands,
                             void* stream) {
    if (rows <= 0 || bad_d(d) || !labels || !labels_local) return PFC_ERR_SHAPE;
    launch_step_kernel(PDL_NORMALISE, l2norm_rows_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        x, nullptr, rows, d, reinterpret_cast<__nv_bfloat16*>(xn), inv_norm, labels, class_start, num_local,
        labels_local, fp16_operands);
    return check_launch();
}

int pfc_cast_f16_to_bf16(const void* src_f16, void* dst_bf16, size_t elems, void* stream) {
    if (elems == 0) return PFC_OK;
    if (elems % 8 || (reinterpret_cast<uintptr_t>(src_f16) & 15) || (reinterpret_cast<uintptr_t>(dst_bf16) & 15))
        return PFC_ERR_ALIGNMENT;
    const size_t n8 = elems / 8;
    cast_f16_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(src_f16), static_cast<uint4*>(dst_bf16), n8);
    return check_launch();
}

int pfc_localize_labels(const int64_t* labels, int B, int64_t class_start, int num_local, int32_t* labels_local,
                        void* stream) {
    if (B <= 0) return PFC_ERR_SHAPE;
    launch_step_kernel(PDL_LABELS, localize_labels_kernel, (B + 255) / 256, 256, 0, (cudaStream_t)stream,
        labels, B, class_start, num_local, labels_local);
    return check_launch();
}

int pfc_row_stats(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                  float* stats, void* stream) {
    if (B <= 0 || n_tiles <= 0) return PFC_ERR_SHAPE;
    const int B_pad = (B + 127) / 128 * 128;
    launch_step_kernel(PDL_STATS, row_stats_kernel, (B + RS_ROWS - 1) / RS_ROWS, RS_ROWS * RS_GROUPS, 0, (cudaStream_t)stream,
        part_sum, n_tiles, B, B_pad, labels_local, tgt_e, stats);
    return check_launch();
}

static PrepArgs prep_args(const float* grad_loss, float s, int B, int d, const int32_t* labels_local, const float* tgt_raw,
                          int margin_kind, float m2, const void* xn, void* xs, float* coef, void* E, int n_pad,
                          int fp16_operands) {
    const double pi = 3.14159265358979323846;
    PrepArgs a;
    a.grad_loss = grad_loss; a.s = s; a.B = B; a.d = d; a.labels = labels_local; a.tgt_raw = tgt_raw;
    a.margin_kind = margin_kind;
    a.cos_m = (float)cos((double)m2); a.sin_m = (float)sin((double)m2); a.theta = (float)cos(pi - (double)m2);
    a.xn = reinterpret_cast<const __nv_bfloat16*>(xn); a.xs = reinterpret_cast<__nv_bfloat16*>(xs);
    a.coef = coef; a.E = reinterpret_cast<__nv_bfloat16*>(E); a.n_pad = n_pad;
    a.xn_f16 = fp16_operands;
    return a;
}

int pfc_row_stats_loss(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                       float* stats, float* row_L, float* loss, unsigned int* ticket, void* stream) {
    if (B <= 0 || n_tiles <= 0 || !ticket) return PFC_ERR_SHAPE;
    const int B_pad = (B + 127) / 128 * 128;
    launch_step_kernel(PDL_STATS, row_stats_loss_kernel<false>, (B + RS_ROWS - 1) / RS_ROWS, RS_ROWS * RS_GROUPS, 0,
        (cudaStream_t)stream, part_sum, n_tiles, B, B_pad, labels_local, tgt_e, stats, row_L, loss, ticket, PrepArgs{});
    return check_launch();
}

int pfc_row_stats_loss_prepare(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                               float* stats, float* row_L, float* loss, unsigned int* ticket, const float* grad_loss,
                               float s, int d, const float* tgt_raw, int margin_kind, float m2, const void* xn, void* xs,
                               float* coef, void* E, int n_pad, int fp16_operands, void* stream) {
    if (B <= 0 || n_tiles <= 0 || !ticket || bad_d(d)) return PFC_ERR_SHAPE;
    const int B_pad = (B + 127) / 128 * 128;
    launch_step_kernel(PDL_STATS, row_stats_loss_kernel<true>, (B + RS_ROWS - 1) / RS_ROWS, RS_ROWS * RS_GROUPS, 0,
        (cudaStream_t)stream, part_sum, n_tiles, B, B_pad, labels_local, tgt_e, stats, row_L, loss, ticket,
        prep_args(grad_loss, s, B, d, labels_local, tgt_raw, margin_kind, m2, xn, xs, coef, E, n_pad, fp16_operands));
    return check_launch();
}

int pfc_loss(const float* stats, int B, float* row_L, float* loss, void* stream) {
    if (B <= 0) return PFC_ERR_SHAPE;
    launch_step_kernel(PDL_STATS, loss_kernel, 1, 1024, 0, (cudaStream_t)stream,
        stats, B, row_L, loss);
    return check_launch();
}

int pfc_backward_prepare(const float* stats, const float* row_L, const float* grad_loss, float s, int B, int d,
                         const int32_t* labels_local, const float* tgt_raw, int margin_kind, float m2,
                         const void* xn, void* xs, float* coef, void* E, int n_pad, int fp16_operands, void* stream) {
    if (B <= 0 || bad_d(d)) return PFC_ERR_SHAPE;
    launch_step_kernel(PDL_PREPARE, backward_prepare_kernel, row_grid(B), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        stats, row_L, prep_args(grad_loss, s, B, d, labels_local, tgt_raw, margin_kind, m2, xn, xs, coef, E, n_pad,
                                fp16_operands));
    return check_launch();
}

int pfc_dx_finalize(const float* partial, int splits, const float* coef, const float* x, const float* inv_norm,
                    float scale, int rows, int rows_total, int d, float* out, void* stream) {
    if (rows <= 0 || bad_d(d) || splits <= 0 || rows_total < rows) return PFC_ERR_SHAPE;
    if (d == 512) {
        launch_step_kernel(PDL_DX_FINAL, dx_finalize_d512_kernel, (rows + 1) / 2, 256, 0, (cudaStream_t)stream,
        partial, splits, static_cast<size_t>(rows_total) * d, coef, x, inv_norm, scale, rows, out);
        return check_launch();
    }
    launch_step_kernel(PDL_DX_FINAL, dx_finalize_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        partial, splits, static_cast<size_t>(rows_total) * d, coef, x, inv_norm, scale, rows, d, out);
    return check_launch();
}

int pfc_dw_finalize(const float* dwn, const float* w, const float* inv_norm_w, int rows, int d, float inv_grad_scale,
                    float* dw, void* stream) {
    if (rows <= 0 || bad_d(d)) return PFC_ERR_SHAPE;
    if (inv_grad_scale != 1.f) return PFC_ERR_SHAPE;   // the un-fused gradient keeps the loss scale (GradScaler.unscale_)
    OptArgs o = {};
    o.kind = OPT_NONE;
    launch_step_kernel(PDL_UPDATE, dw_finalize_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        dwn, const_cast<float*>(w), inv_norm_w, rows, d, o, dw, nullptr, nullptr, nullptr, nullptr);
    return check_launch();
}

int pfc_dw_sgd(const void* dwn, int dwn_bf16, float* w, float* mom, const float* inv_norm_w, int rows, int d, float lr,
               float momentum, float weight_decay, const float* grad_scale, void* wn_next, float* inv_norm_next,
               const int64_t* index, int fp16_operands, void* wn_next_copy_bf16, void* stream) {
    if (rows <= 0 || bad_d(d)) return PFC_ERR_SHAPE;
    if (wn_next_copy_bf16 && (!fp16_operands || !wn_next)) return PFC_ERR_SHAPE;   // the twin of an fp16 shard only
    __nv_bfloat16* wcb = reinterpret_cast<__nv_bfloat16*>(wn_next_copy_bf16);
    if (d % 128 == 0 && mom != nullptr) {
        cudaStream_t st = (cudaStream_t)stream;
        __nv_bfloat16* wnn = reinterpret_cast<__nv_bfloat16*>(wn_next);
        const bool bf = dwn_bf16 != 0;
#define PFC_SGD_CASE(NV) \
    launch_dw_sgd_rows<NV>(dwn, bf, w, mom, inv_norm_w, rows, lr, momentum, weight_decay, grad_scale, wnn, \
                           inv_norm_next, index, fp16_operands, wcb, st)
        switch (d / 128) {
            case 1: PFC_SGD_CASE(1); break;
            case 2: PFC_SGD_CASE(2); break;
            case 3: PFC_SGD_CASE(3); break;
            case 4: PFC_SGD_CASE(4); break;
            case 5: PFC_SGD_CASE(5); break;
            case 6: PFC_SGD_CASE(6); break;
            case 7: PFC_SGD_CASE(7); break;
            default: PFC_SGD_CASE(8); break;
        }
#undef PFC_SGD_CASE
        return check_launch();
    }
    if (dwn_bf16) return PFC_ERR_SHAPE;   // the generic row kernel reads an fp32 gradient
    OptArgs o = {};
    o.kind = OPT_SGD;
    o.lr = lr; o.momentum = momentum; o.wd = weight_decay; o.grad_scale = grad_scale;
    o.index = index;
    o.wn_f16 = fp16_operands;
    o.wn_copy_b = wcb;
    launch_step_kernel(PDL_UPDATE, dw_finalize_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        static_cast<const float*>(dwn), w, inv_norm_w, rows, d, o, nullptr, mom, nullptr,
        reinterpret_cast<__nv_bfloat16*>(wn_next), inv_norm_next);
    return check_launch();
}

int pfc_dw_adam(const float* dwn, float* w, float* exp_avg, float* exp_avg_sq, const float* inv_norm_w, int rows,
                int d, float lr, float beta1, float beta2, float eps, float weight_decay, int step, int decoupled,
                const float* grad_scale, void* wn_next, float* inv_norm_next, const int* step_dev,
                const int64_t* index, int fp16_operands, void* wn_next_copy_bf16, void* stream) {
    if (rows <= 0 || bad_d(d) || (step <= 0 && !step_dev)) return PFC_ERR_SHAPE;
    if (wn_next_copy_bf16 && (!fp16_operands || !wn_next)) return PFC_ERR_SHAPE;
    if (step <= 0) step = 1;
    OptArgs o = {};
    o.kind = decoupled ? OPT_ADAMW : OPT_ADAM;
    o.lr = lr; o.wd = weight_decay; o.beta1 = beta1; o.beta2 = beta2; o.eps = eps;
    o.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    o.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    o.grad_scale = grad_scale;
    o.step_dev = step_dev;
    o.index = index;
    o.wn_f16 = fp16_operands;
    o.wn_copy_b = reinterpret_cast<__nv_bfloat16*>(wn_next_copy_bf16);
    launch_step_kernel(PDL_UPDATE, dw_finalize_kernel, row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream,
        dwn, w, inv_norm_w, rows, d, o, nullptr, exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(wn_next),
        inv_norm_next);
    return check_launch();
}

int pfc_gather_rows(const float* const* src, float* const* dst, int count, const int64_t* index, int rows, int d,
                    void* stream) {
    if (rows < 0 || count < 1 || count > 3 || d <= 0 || (d & 3)) return PFC_ERR_SHAPE;
    if (rows == 0) return PFC_OK;
    RowSet s = {};
    s.count = count;
    for (int i = 0; i < count; ++i) { s.src[i] = src[i]; s.dst[i] = dst[i]; }
    move_rows_kernel<false><<<row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(s, index, rows, d);
    return check_launch();
}

int pfc_scatter_rows(const float* const* src, float* const* dst, int count, const int64_t* index, int rows, int d,
                     void* stream) {
    if (rows < 0 || count < 1 || count > 3 || d <= 0 || (d & 3)) return PFC_ERR_SHAPE;
    if (rows == 0) return PFC_OK;
    RowSet s = {};
    s.count = count;
    for (int i = 0; i < count; ++i) { s.src[i] = src[i]; s.dst[i] = dst[i]; }
    move_rows_kernel<true><<<row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(s, index, rows, d);
    return check_launch();
}

int pfc_margin_apply(const float* logits, const int64_t* labels, int B, int n, int margin_kind, float s, float m2,
                     float m3, float interclass_filtering_threshold, float* out, float* gate, void* stream) {
    if (B <= 0 || n <= 0) return PFC_ERR_SHAPE;
    const double pi = 3.14159265358979323846;
    const size_t total = static_cast<size_t>(B) * n;
    margin_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        logits, labels, B, n, margin_kind, s, (float)cos((double)m2), (float)sin((double)m2),
        (float)cos(pi - (double)m2), (float)(sin(pi - (double)m2) * (double)m2), m3, interclass_filtering_threshold,
        out, gate);
    return check_launch();
}

}  // extern "C"
