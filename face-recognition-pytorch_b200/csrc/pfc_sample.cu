// PartialFC negative-class sampling (reference: nets/PartialFC.py:92-121) as a radix select.
//
// Reference semantics: perm = rand(num_local); perm[positive] = 2.0; index = sort(topk(perm, num_sample).indices);
// if there are more positives than num_sample the index list is the sorted positives.  Both cases are
// "the k_eff = max(num_sample, n_pos) largest keys, emitted in ascending index order", because the forced
// 2.0 is larger than any draw.  Keys are the IEEE bit patterns mapped to an order-preserving uint32; the
// k_eff-th largest key T is found with three MSB-first histogram passes (11 + 11 + 10 bits); elements
// with key > T are taken, elements with key == T are taken lowest-index-first until k_eff is reached
// (the tie rule is ours: torch.topk leaves it implementation-defined), and an ordered stream compaction writes
// the ascending index list and, for positives, their slot (= searchsorted(index, label), :118).
//
// Launches: one memset + six kernels.  Every "pick" / "scan" / "remap" step that needs the result of a whole grid is run by
// the LAST CTA of the kernel that produces it (atomic ticket + fences), not by a kernel of its own: at the shard sizes of
// BASELINE configs[2] / [3] (45 k / 250 k classes per rank) each of those kernels was ~2 us of work behind ~3-6 us of launch
// latency (round 1: 1 + 11 launches, 72 us).
//
// The draw itself stays an input (the reference draws on the CPU generator, :110), so the selected set can be
// compared bit-for-bit with the reference given the same draw.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "pfc_internal.h"

namespace pfc {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;
__host__ __device__ constexpr int radix_bits(int pass) { return pass == 2 ? 10 : 11; }
__host__ __device__ constexpr int radix_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
constexpr int MAX_BINS = 2048;

struct SelState {
    uint32_t prefix;      // bits of T decided so far
    uint32_t k_rem;       // how many still to take among keys matching the prefix
    uint32_t n_pos;       // number of distinct positive classes
    uint32_t k_eff;       // max(num_sample, n_pos)
    uint32_t ticket[5];   // CTAs that have finished: hist passes 0..2, count, compact
    uint32_t hist[3][MAX_BINS];
};

// true in exactly one CTA of the grid: the one whose arrival completes the kernel's global writes (which it may then read
// through L2).  All threads of the CTA must call it.
__device__ __forceinline__ bool last_cta_done(uint32_t* ticket) {
    __shared__ uint32_t last_flag;
    __threadfence();                       // this thread's global writes / atomics are visible before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) last_flag = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    const bool last = last_flag != 0;
    if (last) __threadfence();
    return last;
}

__device__ __forceinline__ uint32_t sortable(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ uint32_t key_of(const float* perm, const uint8_t* flags, int i) {
    return flags[i] ? sortable(2.0f) : sortable(perm[i]);
}

__global__ void mark_positive_kernel(const int32_t* __restrict__ labels, int B, uint8_t* __restrict__ flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < B && labels[j] >= 0) flags[labels[j]] = 1;
}

template <int PASS>
__device__ void pick_digit(SelState* st, int num_sample, int nl);

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS)
hist_kernel(const float* __restrict__ perm, const uint8_t* __restrict__ flags, int nl, SelState* st, int num_sample) {
    __shared__ uint32_t h[MAX_BINS];
    __shared__ uint32_t npos_s;
    for (int b = threadIdx.x; b < MAX_BINS; b += SEL_THREADS) h[b] = 0;
    if (threadIdx.x == 0) npos_s = 0;
    __syncthreads();
    const uint32_t prefix = PASS ? st->prefix : 0;
    constexpr uint32_t hi_mask = PASS == 0 ? 0u : (PASS == 1 ? 0xFFE00000u : 0xFFFFFC00u);
    constexpr uint32_t dmask = (1u << radix_bits(PASS)) - 1;
    uint32_t np = 0;
    for (int i = blockIdx.x * SEL_THREADS + threadIdx.x; i < nl; i += gridDim.x * SEL_THREADS) {
        const uint32_t k = key_of(perm, flags, i);
        if (PASS == 0) np += flags[i];
        if ((k & hi_mask) == prefix) atomicAdd(&h[(k >> radix_shift(PASS)) & dmask], 1u);
    }
    if (PASS == 0 && np) atomicAdd(&npos_s, np);
    __syncthreads();
    for (int b = threadIdx.x; b < MAX_BINS; b += SEL_THREADS)
        if (h[b]) atomicAdd(&st->hist[PASS][b], h[b]);
    if (PASS == 0 && threadIdx.x == 0 && npos_s) atomicAdd(&st->n_pos, npos_s);
    if (last_cta_done(&st->ticket[PASS])) pick_digit<PASS>(st, num_sample, nl);     // the histogram is complete
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t& total);

// Pick the digit of this pass: walking the bins from the top down until the running count reaches k_rem stops at
// b* = max{ b >= 1 : I(b) >= k_rem } (0 if there is none), I(b) = sum of hist[j] over j >= b, and leaves
// k_rem - (I(b*) - hist[b*]).  I is an inclusive scan over the bins in reversed order; the first reversed position that
// reaches k_rem is taken with an atomicMin.  (A single-thread walk, one dependent L2 load per bin, cost 50-100 us of a
// 190 us sampler in round 1; this is one CTA-wide scan.)
template <int PASS>
__device__ void pick_digit(SelState* st, int num_sample, int nl) {
    constexpr int bins = 1 << radix_bits(PASS);
    constexpr int per = (bins + 1023) / 1024;
    __shared__ uint32_t wt[33];
    __shared__ uint32_t incl[MAX_BINS];
    __shared__ uint32_t hs[MAX_BINS];
    __shared__ uint32_t rstar, rem_s;
    if (threadIdx.x == 0) {
        if (PASS == 0) {
            const uint32_t np = __ldcg(&st->n_pos);
            uint32_t k = np > (uint32_t)num_sample ? np : (uint32_t)num_sample;
            if (k > (uint32_t)nl) k = nl;
            st->k_eff = k;
            st->k_rem = k;
            st->prefix = 0;
            rem_s = k;
        } else {
            rem_s = st->k_rem;
        }
        rstar = bins - 1;                     // reversed position of bin 0: the walk's default
    }
    __syncthreads();
    const uint32_t rem = rem_s;
    if (rem == 0) {                           // nothing to select: threshold above every key (CTA-uniform branch)
        if (threadIdx.x == 0) st->prefix = 0xFFFFFFFFu;
        return;
    }
    uint32_t v[per], sum = 0;
#pragma unroll
    for (int u = 0; u < per; ++u) {
        const int r = threadIdx.x * per + u;  // reversed bin index: r = 0 is the top bin
        const uint32_t c = r < bins ? __ldcg(&st->hist[PASS][bins - 1 - r]) : 0u;
        v[u] = c;
        sum += c;
        if (r < bins) hs[r] = c;
    }
    uint32_t tot;
    uint32_t run = block_excl_scan(sum, wt, tot);
#pragma unroll
    for (int u = 0; u < per; ++u) {
        const int r = threadIdx.x * per + u;
        run += v[u];
        if (r < bins) {
            incl[r] = run;
            if (run >= rem && r < bins - 1) atomicMin(&rstar, static_cast<uint32_t>(r));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t r = rstar;
        st->prefix |= static_cast<uint32_t>(bins - 1 - r) << radix_shift(PASS);
        st->k_rem = rem - (incl[r] - hs[r]);
    }
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t& total) {
    // exclusive scan over the 1024 threads of the CTA (thread order)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t t = warp_tot[lane];
        uint32_t ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += n;
        }
        warp_tot[lane] = ti - t;           // exclusive warp offsets
        if (lane == 31) warp_tot[32] = ti;  // grand total
    }
    __syncthreads();
    total = warp_tot[32];
    const uint32_t r = warp_tot[w] + inc - v;
    __syncthreads();
    return r;
}

// per-tile counts of (key > T) and (key == T)
__global__ void __launch_bounds__(SEL_THREADS)
count_kernel(const float* __restrict__ perm, const uint8_t* __restrict__ flags, int nl, SelState* st,
             uint32_t* gt_cnt, uint32_t* eq_cnt, int tiles) {
    __shared__ uint32_t sg, se;
    if (threadIdx.x == 0) { sg = 0; se = 0; }
    __syncthreads();
    const uint32_t T = st->prefix;
    const bool none = st->k_eff == 0;
    uint32_t g = 0, e = 0;
    const int base = blockIdx.x * SEL_TILE + threadIdx.x * SEL_ITEMS;
#pragma unroll
    for (int u = 0; u < SEL_ITEMS; ++u) {
        const int i = base + u;
        if (i < nl && !none) {
            const uint32_t k = key_of(perm, flags, i);
            g += k > T;
            e += k == T;
        }
    }
    g = __reduce_add_sync(0xffffffffu, g);
    e = __reduce_add_sync(0xffffffffu, e);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sg, g); atomicAdd(&se, e); }
    __syncthreads();
    if (threadIdx.x == 0) { gt_cnt[blockIdx.x] = sg; eq_cnt[blockIdx.x] = se; }
    if (!last_cta_done(&st->ticket[3])) return;
    // last CTA: exclusive scan of the per-tile counts (tiles <= a few hundred: serial chunks of 1024)
    __shared__ uint32_t wt[33];
    uint32_t carry_g = 0, carry_e = 0;
    for (int b0 = 0; b0 < tiles; b0 += SEL_THREADS) {
        const int i = b0 + threadIdx.x;
        const uint32_t g = i < tiles ? __ldcg(gt_cnt + i) : 0, e = i < tiles ? __ldcg(eq_cnt + i) : 0;
        uint32_t tg, te;
        const uint32_t xg = block_excl_scan(g, wt, tg);
        const uint32_t xe = block_excl_scan(e, wt, te);
        if (i < tiles) { gt_cnt[i] = carry_g + xg; eq_cnt[i] = carry_e + xe; }
        carry_g += tg;
        carry_e += te;
    }
}

// ordered compaction: ascending index list + slot of every positive class
__global__ void __launch_bounds__(SEL_THREADS)
compact_kernel(const float* __restrict__ perm, const uint8_t* __restrict__ flags, int nl, SelState* st,
               const uint32_t* __restrict__ gt_off, const uint32_t* __restrict__ eq_off,
               int64_t* __restrict__ index_out, int32_t* slot_of, int32_t* __restrict__ n_out,
               const int32_t* __restrict__ labels, int B, int32_t* __restrict__ labels_out) {
    __shared__ uint32_t wt[33];
    const uint32_t T = st->prefix;
    const uint32_t need_eq = st->k_rem;
    const bool none = st->k_eff == 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) n_out[0] = static_cast<int32_t>(st->k_eff);
    const int base = blockIdx.x * SEL_TILE + threadIdx.x * SEL_ITEMS;
    uint32_t isgt = 0, iseq = 0;   // bit u = item u
    uint32_t ng = 0, ne = 0;
#pragma unroll
    for (int u = 0; u < SEL_ITEMS; ++u) {
        const int i = base + u;
        if (i < nl && !none) {
            const uint32_t k = key_of(perm, flags, i);
            if (k > T) { isgt |= 1u << u; ++ng; }
            if (k == T) { iseq |= 1u << u; ++ne; }
        }
    }
    uint32_t tot;
    const uint32_t eq_before_blk = eq_off[blockIdx.x];
    uint32_t eq_rank = eq_before_blk + block_excl_scan(ne, wt, tot);
    // selected flags now known per item
    uint32_t sel = isgt, nsel = ng;
#pragma unroll
    for (int u = 0; u < SEL_ITEMS; ++u) {
        if (iseq & (1u << u)) {
            if (eq_rank < need_eq) { sel |= 1u << u; ++nsel; }
            ++eq_rank;
        }
    }
    const uint32_t sel_before_blk = gt_off[blockIdx.x] + min(eq_before_blk, need_eq);
    uint32_t pos = sel_before_blk + block_excl_scan(nsel, wt, tot);
#pragma unroll
    for (int u = 0; u < SEL_ITEMS; ++u) {
        if (sel & (1u << u)) {
            const int i = base + u;
            index_out[pos] = i;
            if (flags[i]) slot_of[i] = static_cast<int32_t>(pos);
            ++pos;
        }
    }
    if (!last_cta_done(&st->ticket[4])) return;
    // last CTA: every positive class has its slot -- labels -> position in the index list (= searchsorted, :118)
    for (int j = threadIdx.x; j < B; j += SEL_THREADS) {
        const int l = labels[j];
        labels_out[j] = l >= 0 ? __ldcg(slot_of + l) : -1;
    }
}

struct SelLayout {
    size_t flags, state, gt, eq, slot, total;
    int tiles;
};
static SelLayout sel_layout(int nl) {
    SelLayout L;
    L.tiles = (nl + SEL_TILE - 1) / SEL_TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    L.flags = take(static_cast<size_t>(nl));
    L.state = take(sizeof(SelState));
    L.gt = take(sizeof(uint32_t) * L.tiles);
    L.eq = take(sizeof(uint32_t) * L.tiles);
    L.slot = take(sizeof(int32_t) * static_cast<size_t>(nl));
    L.total = off;
    return L;
}

}  // namespace pfc

using namespace pfc;

extern "C" {

size_t pfc_sample_workspace_bytes(int num_local) { return num_local > 0 ? sel_layout(num_local).total : 0; }

int pfc_sample(const float* perm, const int32_t* labels_local, int B, int num_local, int num_sample,
               int64_t* index_out, int32_t* n_out, int32_t* labels_remapped, void* workspace,
               size_t workspace_bytes, void* stream_) {
    if (B <= 0 || num_local <= 0 || num_sample < 0) return PFC_ERR_SHAPE;
    const SelLayout L = sel_layout(num_local);
    if (workspace_bytes < L.total) return PFC_ERR_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)stream_;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint8_t* flags = ws + L.flags;
    SelState* st = reinterpret_cast<SelState*>(ws + L.state);
    uint32_t* gt = reinterpret_cast<uint32_t*>(ws + L.gt);
    uint32_t* eq = reinterpret_cast<uint32_t*>(ws + L.eq);
    int32_t* slot = reinterpret_cast<int32_t*>(ws + L.slot);
    // flags and the selection state are contiguous at the front of the workspace
    if (cudaMemsetAsync(ws, 0, L.gt, stream) != cudaSuccess) return PFC_ERR_CUDA;
    mark_positive_kernel<<<(B + 255) / 256, 256, 0, stream>>>(labels_local, B, flags);
    int hb = L.tiles;   // one CTA per SEL_TILE keys keeps every SM busy for the big shards, 1 CTA for small ones
    hist_kernel<0><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);     // + pick of digit 0
    hist_kernel<1><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);
    hist_kernel<2><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);
    count_kernel<<<L.tiles, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, gt, eq, L.tiles);   // + tile scan
    compact_kernel<<<L.tiles, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, gt, eq, index_out, slot, n_out,
                                                        labels_local, B, labels_remapped);          // + label remap
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

}  // extern "C"
