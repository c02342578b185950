// Peer-memory (NVLink / NVSwitch) versions of the head's three exchanges, fused into the kernels that produce the
// data: every rank stores its contribution straight into every peer's symmetric buffer, a flag barrier publishes
// the stores, and the consumer is a plain local kernel.  They replace the NCCL all-gather / all-reduce /
// reduce-scatter of the step (reference: nets/PartialFC.py:182-186, :448/:453/:459, :505-522), each of which costs
// ~30 us inside a CUDA graph at 8 ranks for a payload of a few KB to 2 MB.
//
//   pfc_peer_l2norm_gather : xn = normalise(x) as bf16 -> rows [rank*b, rank*b+b) of EVERY rank's xn_all (+ labels)
//   pfc_peer_row_stats     : this rank's [B,2] softmax statistics -> slot `rank` of EVERY rank's slots[W][B][2]
//   pfc_peer_loss          : sum of the W slots in rank order (bit-identical on all ranks) -> stats, row_L, loss
//   pfc_peer_dx_scatter    : c_i * sum_z partial[z][i,:] for the rows owned by rank r -> slot `rank` of rank r's
//                            dx_slots[W][b][d]   (the consumer is pfc_dx_finalize with splits = W)
//   pfc_peer_barrier       : all ranks' previous stores are visible everywhere after it (one tiny kernel per rank)
// The consumers of the exchanged data take the barrier at their OWN start instead of behind a separate launch
// (pfc_peer_localize_labels, pfc_peer_loss, pfc_peer_dx_finalize): CTA 0 signals the peers, every CTA polls this
// rank's flags, and the last CTA through advances the epoch -- three launches fewer per step.
//
// Peer buffers come from torch's symmetric-memory allocator (host code passes the mapped device pointers); the
// barrier uses one uint32 flag per (receiver, sender) pair and a per-rank epoch counter kept in device memory, so the
// same launch sequence can be replayed from a CUDA graph.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "pfc_internal.h"
#include "pfc_launch.cuh"
#include "pfc_prepare.cuh"

namespace pfc {

constexpr int MAX_PEERS = 16;
struct PeerPtrs {
    void* p[MAX_PEERS];
};

// How long a rank waits for its peers at a flag barrier before it gives up (printf + trap, which surfaces as a CUDA
// error on this rank instead of a silent hang).  NCCL's own watchdog default is 10 minutes; a barrier here is reached by
// every rank once per step, so the only legitimate long waits are host-side stalls of a peer (data-loader hiccup,
// rank-0 checkpoint / validation, a debugger).  Default 10 minutes of SM clock at 2 GHz; pfc_peer_set_timeout_ms()
// changes it (0 = wait forever).  After the first ~100 us a waiting thread backs off with nanosleep.
__device__ long long g_peer_timeout_cycles = 1200000000000LL;

__device__ __forceinline__ void peer_wait_flag(const uint32_t* mine, uint32_t ep, int rank, int from) {
    uint32_t v;
    const long long t0 = clock64();
    const long long limit = *reinterpret_cast<volatile long long*>(&g_peer_timeout_cycles);
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (static_cast<int32_t>(v - ep) >= 0) break;
        if (++spins > 256) {
            __nanosleep(spins > 65536 ? 2000 : 100);
            if (limit > 0 && clock64() - t0 > limit) {
                printf("pfc: peer barrier timed out (rank %d waiting for %d, epoch %u, saw %u)\n", rank, from, ep, v);
                __trap();
            }
        }
    }
}

__device__ __forceinline__ float warp_sum_p(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One CTA, W threads.  flags.p[q] -> rank q's flag array uint32[W]; counter -> this rank's epoch (device memory).
__global__ void peer_barrier_kernel(PeerPtrs flags, uint32_t* counter, int rank, int W) {
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) {
        ep_s = *counter + 1;
        *counter = ep_s;
    }
    __syncthreads();
    const uint32_t ep = ep_s;
    __threadfence_system();
    if (threadIdx.x < W) {
        uint32_t* remote = static_cast<uint32_t*>(flags.p[threadIdx.x]) + rank;    // my slot in peer's array
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(ep) : "memory");
        peer_wait_flag(static_cast<const uint32_t*>(flags.p[rank]) + threadIdx.x, ep, rank, threadIdx.x);
    }
    __threadfence_system();
}

// Entry barrier of a consumer kernel (every thread of every CTA calls it first).  The producing kernel of THIS rank
// has completed (stream order), so its stores only need the system-scope fence before the flags go out.
// state[0] = epoch counter, state[1] = ticket of the CTAs that have passed (both zeroed once by the host).
__device__ __forceinline__ void peer_entry_barrier(const PeerPtrs& flags, uint32_t* state, int rank, int W) {
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) ep_s = *reinterpret_cast<volatile uint32_t*>(state) + 1;
    __syncthreads();
    const uint32_t ep = ep_s;
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        uint32_t* remote = static_cast<uint32_t*>(flags.p[threadIdx.x]) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(ep) : "memory");
    }
    if (threadIdx.x < W) {
        peer_wait_flag(static_cast<const uint32_t*>(flags.p[rank]) + threadIdx.x, ep, rank, threadIdx.x);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // every CTA has read the epoch before it takes a ticket, so the last one may advance it
        if (atomicAdd(state + 1, 1u) == gridDim.x - 1) {
            state[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(state) = ep;
        }
    }
}

// barrier + labels -> shard-local ids (-1 for classes of another rank), nets/PartialFC.py:188-193
__global__ void __launch_bounds__(256)
peer_localize_labels_kernel(PeerPtrs flags, uint32_t* state, int rank, int W, const int64_t* labels, int B,
                            int64_t class_start, int num_local, int32_t* __restrict__ out) {
    peer_entry_barrier(flags, state, rank, W);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int64_t l = reinterpret_cast<const volatile int64_t*>(labels)[i] - class_start;
    out[i] = (l >= 0 && l < num_local) ? static_cast<int32_t>(l) : -1;
}

// warp per row; lane l handles float4 #l, #l+32, ... (d <= 1024)
__global__ void __launch_bounds__(256)
peer_l2norm_gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int b, int d, int rank, int W,
                          PeerPtrs xn_all, PeerPtrs labels_all, float* __restrict__ inv_norm, int f16) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= b) return;
    const float* xr = x + static_cast<size_t>(row) * d;
    const int nv = d >> 2;
    float4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (c < nv) {
            v[j] = *reinterpret_cast<const float4*>(xr + 4 * c);
            ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
        }
    }
    ss = warp_sum_p(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    const size_t grow = static_cast<size_t>(rank) * b + row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (c < nv) {
            uint2 pk;
            if (f16) {       // the reference's AMP operands (nets/PartialFC.py:198)
                __half2 a = __floats2half2_rn(v[j].x / denom, v[j].y / denom);
                __half2 bb = __floats2half2_rn(v[j].z / denom, v[j].w / denom);
                pk.x = *reinterpret_cast<uint32_t*>(&a);
                pk.y = *reinterpret_cast<uint32_t*>(&bb);
            } else {
                __nv_bfloat162 a = __floats2bfloat162_rn(v[j].x / denom, v[j].y / denom);
                __nv_bfloat162 bb = __floats2bfloat162_rn(v[j].z / denom, v[j].w / denom);
                pk.x = *reinterpret_cast<uint32_t*>(&a);
                pk.y = *reinterpret_cast<uint32_t*>(&bb);
            }
            for (int q = 0; q < W; ++q)
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(xn_all.p[q]) + grow * d + 4 * c) = pk;
        }
    }
    if (lane == 0) {
        inv_norm[row] = 1.f / denom;
        const int64_t l = labels[row];
        for (int q = 0; q < W; ++q) static_cast<int64_t*>(labels_all.p[q])[grow] = l;
    }
}

// same reduction as row_stats_kernel (pfc_rows.cu: 8 rows x 32 slab groups per CTA, fixed summation order), result
// stored into slot `rank` of every peer
constexpr int PRS_ROWS = 8, PRS_GROUPS = 64;
__global__ void __launch_bounds__(PRS_ROWS * PRS_GROUPS)
peer_row_stats_kernel(const float* __restrict__ part_sum, int n_tiles, int B, int B_pad,
                      const int32_t* __restrict__ labels, const float* __restrict__ tgt_e, int rank, int W,
                      PeerPtrs slots) {
    __shared__ float red[PRS_GROUPS][PRS_ROWS + 1];
    const int r = threadIdx.x & (PRS_ROWS - 1), g = threadIdx.x / PRS_ROWS;
    const int row = blockIdx.x * PRS_ROWS + r;
    float s = 0.f;
    if (row < B) {
        int t = g;
        for (; t + 3 * PRS_GROUPS < n_tiles; t += 4 * PRS_GROUPS) {
            const float a = part_sum[static_cast<size_t>(t) * B_pad + row];
            const float b = part_sum[static_cast<size_t>(t + PRS_GROUPS) * B_pad + row];
            const float c = part_sum[static_cast<size_t>(t + 2 * PRS_GROUPS) * B_pad + row];
            const float d = part_sum[static_cast<size_t>(t + 3 * PRS_GROUPS) * B_pad + row];
            s += a; s += b; s += c; s += d;
        }
        for (; t < n_tiles; t += PRS_GROUPS) s += part_sum[static_cast<size_t>(t) * B_pad + row];
    }
    red[g][r] = s;
    __syncthreads();
    if (g == 0 && row < B) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < PRS_GROUPS; ++k) tot += red[k][r];
        const float2 v = make_float2(tot, (labels[row] >= 0) ? tgt_e[row] : 0.f);
        for (int q = 0; q < W; ++q)
            reinterpret_cast<float2*>(static_cast<float*>(slots.p[q]) + static_cast<size_t>(rank) * B * 2)[row] = v;
    }
}

// stats[i] = sum_r slots[r][i] (rank order), row_L, loss -- the local half of the exchange + pfc_loss
template <bool kBarrier>
__global__ void __launch_bounds__(1024)
peer_loss_kernel(PeerPtrs flags, uint32_t* state, int rank, const float* slots, int W, int B, float* __restrict__ stats,
                 float* __restrict__ row_L, float* __restrict__ loss) {
    __shared__ float red[32];
    if (kBarrier) peer_entry_barrier(flags, state, rank, W);
    float acc = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) {
        float others = 0.f, te = 0.f;
        for (int r = 0; r < W; ++r) {
            // peers wrote the slots: plain (not read-only-cache) loads
            float2 v;
            asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];"
                         : "=f"(v.x), "=f"(v.y) : "l"(slots + (static_cast<size_t>(r) * B + i) * 2));
            others += v.x;
            te += v.y;
        }
        stats[2 * i] = others;
        stats[2 * i + 1] = te;
        const float L = others + te;
        row_L[i] = L;
        acc -= logf(fmaxf(te / L, 1e-30f));
    }
    acc = warp_sum_p(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = warp_sum_p(red[threadIdx.x]);
        if (threadIdx.x == 0) loss[0] = v / static_cast<float>(B);
    }
}

// The no-autograd step's "barrier -> loss -> coefficients" in one launch: every CTA takes the entry barrier, then one warp
// per row sums the W slots of its row in rank order (identical bits on all ranks), writes stats / row_L and forms the
// backward coefficients of the row (pfc_backward_prepare); the last CTA through (ticket) reduces the loss from the finished
// statistics in loss_kernel's order.  Replaces pfc_peer_loss (one 1024-thread CTA) + pfc_backward_prepare.
constexpr int PLP_WARPS = 8;
__global__ void __launch_bounds__(PLP_WARPS * 32)
peer_loss_prepare_kernel(PeerPtrs flags, uint32_t* state, int rank, const float* slots, int W, int B, float* stats,
                         float* __restrict__ row_L, float* __restrict__ loss, unsigned int* ticket, PrepArgs pa) {
    __shared__ bool last;
    peer_entry_barrier(flags, state, rank, W);
    const int row = blockIdx.x * PLP_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row < B) {
        float2 v = make_float2(0.f, 0.f);
        if (lane < W)      // peers wrote the slots: plain (not read-only-cache) loads, one rank per lane
            asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];"
                         : "=f"(v.x), "=f"(v.y) : "l"(slots + (static_cast<size_t>(lane) * B + row) * 2));
        float others = 0.f, te = 0.f;
        for (int r = 0; r < W; ++r) {                       // rank order, like peer_loss_kernel
            others += __shfl_sync(0xffffffffu, v.x, r);
            te += __shfl_sync(0xffffffffu, v.y, r);
        }
        const float L = others + te;
        if (lane == 0) {
            stats[2 * row] = others;
            stats[2 * row + 1] = te;
            row_L[row] = L;
        }
        prepare_row(pa, row, L, others, lane, 32);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    loss_from_stats<PLP_WARPS * 32>(stats, B, row_L, loss);
    if (threadIdx.x == 0) *ticket = 0;
}

// row i of the global batch belongs to rank i / b; its scaled dXn partial goes to that rank's slot `rank`
__global__ void __launch_bounds__(256)
peer_dx_scatter_kernel(const float* __restrict__ partial, int splits, size_t split_stride,
                       const float* __restrict__ coef, int B, int b, int d, int rank, PeerPtrs dx_slots) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const int dst = row / b, lr = row - dst * b;
    float* out = static_cast<float*>(dx_slots.p[dst]) + (static_cast<size_t>(rank) * b + lr) * d;
    const float c = coef[row];
    for (int k = lane; k < (d >> 2); k += 32) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* pp = partial + static_cast<size_t>(row) * d + 4 * k;
        for (int z = 0; z < splits; z += 4) {                // four split slabs in flight, summed in slab order
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                p[u] = (z + u < splits) ? *reinterpret_cast<const float4*>(pp + (z + u) * split_stride)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) { a.x += p[u].x; a.y += p[u].y; a.z += p[u].z; a.w += p[u].w; }
        }
        a.x *= c; a.y *= c; a.z *= c; a.w *= c;
        *reinterpret_cast<float4*>(out + 4 * k) = a;
    }
}

// barrier + dx = W * normalize_backward(sum over the W slots of this rank's dx_slots [W][b][d])   (:505-522)
__global__ void __launch_bounds__(256)
peer_dx_finalize_kernel(PeerPtrs flags, uint32_t* state, int rank, int W, const float* slots, const float* __restrict__ x,
                        const float* __restrict__ inv_norm, float scale, int b, int d, float* __restrict__ out) {
    peer_entry_barrier(flags, state, rank, W);
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= b) return;
    const int nv = d >> 2;
    const float inv = inv_norm[row];
    float4 g[8], xv[8];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int z = 0; z < W; ++z) {
                const float* p = slots + (static_cast<size_t>(z) * b + row) * d + 4 * k;
                float4 q;
                asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p));
                a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
            }
            g[j] = a;
            float4 q = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d + 4 * k);
            q.x *= inv; q.y *= inv; q.z *= inv; q.w *= inv;
            xv[j] = q;
            dot += q.x * a.x + q.y * a.y + q.z * a.z + q.w * a.w;
        }
    }
    dot = warp_sum_p(dot);
    const float m = scale * inv;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;
        if (k < nv) {
            float4 a = g[j];
            a.x = (a.x - xv[j].x * dot) * m; a.y = (a.y - xv[j].y * dot) * m;
            a.z = (a.z - xv[j].z * dot) * m; a.w = (a.w - xv[j].w * dot) * m;
            *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * d + 4 * k) = a;
        }
    }
}

static int fill_peers(PeerPtrs* pp, void* const* ptrs, int W) {
    if (W < 1 || W > MAX_PEERS || ptrs == nullptr) return PFC_ERR_SHAPE;
    for (int i = 0; i < MAX_PEERS; ++i) pp->p[i] = i < W ? ptrs[i] : nullptr;
    return PFC_OK;
}
static inline int launched() { return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH; }

}  // namespace pfc

using namespace pfc;

extern "C" {

// Barrier timeout of the peer exchanges in milliseconds of SM time at 2 GHz (0 = never give up); see peer_wait_flag.
int pfc_peer_set_timeout_ms(double ms) {
    const long long cycles = ms <= 0 ? 0 : static_cast<long long>(ms * 2.0e6);
    return cudaMemcpyToSymbol(pfc::g_peer_timeout_cycles, &cycles, sizeof(cycles)) == cudaSuccess ? PFC_OK : PFC_ERR_CUDA;
}

int pfc_peer_max_ranks(void) { return MAX_PEERS; }

int pfc_peer_barrier(void* const* peer_flags, uint32_t* epoch_counter, int rank, int W, void* stream) {
    PeerPtrs f;
    int rc = fill_peers(&f, peer_flags, W);
    if (rc) return rc;
    launch_step_kernel(PDL_LABELS, peer_barrier_kernel, 1, 32, 0, (cudaStream_t)stream,
                       f, epoch_counter, rank, W);
    return launched();
}

int pfc_peer_l2norm_gather(const float* x, const int64_t* labels, int b, int d, int rank, int W,
                           void* const* peer_xn_all, void* const* peer_labels_all, float* inv_norm, int fp16_operands,
                           void* stream) {
    if (b <= 0 || d <= 0 || (d & 7) || d > 1024) return PFC_ERR_SHAPE;
    PeerPtrs xa, la;
    int rc = fill_peers(&xa, peer_xn_all, W);
    if (rc) return rc;
    rc = fill_peers(&la, peer_labels_all, W);
    if (rc) return rc;
    launch_step_kernel(PDL_NORMALISE, peer_l2norm_gather_kernel, (b + 7) / 8, 256, 0, (cudaStream_t)stream,
                       x, labels, b, d, rank, W, xa, la, inv_norm, fp16_operands);
    return launched();
}

int pfc_peer_row_stats(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                       int rank, int W, void* const* peer_slots, void* stream) {
    if (B <= 0 || n_tiles <= 0) return PFC_ERR_SHAPE;
    PeerPtrs s;
    int rc = fill_peers(&s, peer_slots, W);
    if (rc) return rc;
    const int B_pad = (B + 127) / 128 * 128;
    launch_step_kernel(PDL_STATS, peer_row_stats_kernel, (B + PRS_ROWS - 1) / PRS_ROWS, PRS_ROWS * PRS_GROUPS, 0,
                       (cudaStream_t)stream, part_sum, n_tiles, B, B_pad, labels_local, tgt_e, rank, W, s);
    return launched();
}

int pfc_peer_loss(void* const* peer_flags, uint32_t* barrier_state, int rank, const float* slots, int W, int B,
                  float* stats, float* row_L, float* loss, void* stream) {
    if (B <= 0 || W < 1) return PFC_ERR_SHAPE;
    PeerPtrs f = {};
    if (peer_flags) {
        int rc = fill_peers(&f, peer_flags, W);
        if (rc) return rc;
        if (!barrier_state) return PFC_ERR_SHAPE;
        launch_step_kernel(PDL_STATS, peer_loss_kernel<true>, 1, 1024, 0, (cudaStream_t)stream,
                       f, barrier_state, rank, slots, W, B, stats, row_L, loss);
    } else {
        launch_step_kernel(PDL_STATS, peer_loss_kernel<false>, 1, 1024, 0, (cudaStream_t)stream,
                       f, nullptr, rank, slots, W, B, stats, row_L, loss);
    }
    return launched();
}

int pfc_peer_loss_prepare(void* const* peer_flags, uint32_t* barrier_state, int rank, const float* slots, int W, int B,
                          float* stats, float* row_L, float* loss, unsigned int* ticket, const float* grad_loss, float s,
                          int d, const int32_t* labels_local, const float* tgt_raw, int margin_kind, float m2,
                          const void* xn, void* xs, float* coef, void* E, int n_pad, int fp16_operands, void* stream) {
    if (B <= 0 || W < 1 || W > 32 || !barrier_state || !ticket || d <= 0 || (d & 7) || d > 1024) return PFC_ERR_SHAPE;
    PeerPtrs f;
    int rc = fill_peers(&f, peer_flags, W);
    if (rc) return rc;
    const double pi = 3.14159265358979323846;
    PrepArgs a;
    a.grad_loss = grad_loss; a.s = s; a.B = B; a.d = d; a.labels = labels_local; a.tgt_raw = tgt_raw;
    a.margin_kind = margin_kind;
    a.cos_m = (float)cos((double)m2); a.sin_m = (float)sin((double)m2); a.theta = (float)cos(pi - (double)m2);
    a.xn = reinterpret_cast<const __nv_bfloat16*>(xn); a.xs = reinterpret_cast<__nv_bfloat16*>(xs);
    a.coef = coef; a.E = reinterpret_cast<__nv_bfloat16*>(E); a.n_pad = n_pad;
    a.xn_f16 = fp16_operands;
    launch_step_kernel(PDL_STATS, peer_loss_prepare_kernel, (B + PLP_WARPS - 1) / PLP_WARPS, PLP_WARPS * 32, 0,
                       (cudaStream_t)stream, f, barrier_state, rank, slots, W, B, stats, row_L, loss, ticket, a);
    return launched();
}

int pfc_peer_localize_labels(void* const* peer_flags, uint32_t* barrier_state, int rank, int W, const int64_t* labels,
                             int B, int64_t class_start, int num_local, int32_t* labels_local, void* stream) {
    if (B <= 0 || !barrier_state) return PFC_ERR_SHAPE;
    PeerPtrs f;
    int rc = fill_peers(&f, peer_flags, W);
    if (rc) return rc;
    launch_step_kernel(PDL_LABELS, peer_localize_labels_kernel, (B + 255) / 256, 256, 0, (cudaStream_t)stream,
                       f, barrier_state, rank, W, labels, B, class_start, num_local, labels_local);
    return launched();
}

int pfc_peer_dx_finalize(void* const* peer_flags, uint32_t* barrier_state, int rank, int W, const float* dx_slots,
                         const float* x, const float* inv_norm, float scale, int b, int d, float* out, void* stream) {
    if (b <= 0 || d <= 0 || (d & 7) || d > 1024 || !barrier_state || !x || !inv_norm) return PFC_ERR_SHAPE;
    PeerPtrs f;
    int rc = fill_peers(&f, peer_flags, W);
    if (rc) return rc;
    launch_step_kernel(PDL_DX_FINAL, peer_dx_finalize_kernel, (b + 7) / 8, 256, 0, (cudaStream_t)stream,
                       f, barrier_state, rank, W, dx_slots, x, inv_norm, scale, b, d, out);
    return launched();
}

int pfc_peer_dx_scatter(const float* partial, int splits, const float* coef, int B, int b, int d, int rank, int W,
                        void* const* peer_dx_slots, void* stream) {
    if (B <= 0 || b <= 0 || B != b * W || d <= 0 || (d & 7) || splits <= 0) return PFC_ERR_SHAPE;
    PeerPtrs s;
    int rc = fill_peers(&s, peer_dx_slots, W);
    if (rc) return rc;
    launch_step_kernel(PDL_DX_FINAL, peer_dx_scatter_kernel, (B + 7) / 8, 256, 0, (cudaStream_t)stream,
                       partial, splits, static_cast<size_t>(B) * d, coef, B, b, d, rank, s);
    return launched();
}

}  // extern "C"
