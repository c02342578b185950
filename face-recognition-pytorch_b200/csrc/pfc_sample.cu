// PartialFC negative-class sampling (reference: nets/PartialFC.py:92-121) as a radix select -- ONE launch.
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
// One thread-block cluster of 8 CTAs x 1024 threads runs the whole selection: CTA c owns the contiguous slice
// [c S, (c+1) S) of the shard.  The positive-class marks are a BITMAP in shared memory (no global flag array, no memset),
// the slice's keys are staged in shared memory once (up to 45 k classes per CTA), the digit histograms live in shared
// memory and are merged through distributed shared memory (CTA c sums bins [c bins/8, (c+1) bins/8) of the eight
// histograms and stores the sums into every CTA's merged copy; each CTA then picks the digit itself -- same inputs, same
// result), the per-CTA counts of (key > T) / (key == T) are exchanged the same way, the compaction is warp-ordered
// (ballot ranks, no CTA-wide scan per tile), and the cluster barrier (release / acquire) orders the slot table in global
// memory before the labels are remapped.  Rounds 1-2 ran this as 1 memset + 11, then 6, dependent launches
// (72 -> 54 us at the shard sizes of BASELINE configs[2] / [3], all of it launch latency: each kernel had ~2 us of work).
//
// The draw itself stays an input (the reference draws on the CPU generator, :110), so the selected set can be
// compared bit-for-bit with the reference given the same draw.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "pfc_internal.h"

namespace cg = cooperative_groups;

namespace pfc {

#ifdef PFC_SAMPLE_STAMPS      // tools/probe/sample_phases.cu: globaltimer stamps of CTA 0, thread 0 after every phase
__device__ unsigned long long g_stamps[16];
#define STAMP(i)                                                                              \
    do {                                                                                      \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                                            \
            unsigned long long t_;                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                            \
            g_stamps[i] = t_;                                                                 \
        }                                                                                     \
    } while (0)
#else
#define STAMP(i)
#endif

constexpr int SEL_THREADS = 1024;
constexpr int SEL_CLUSTER_MAX = 16;     // 8 = the portable maximum; 16 needs the non-portable opt-in (see pick_cluster)
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;
__host__ __device__ constexpr int radix_bits(int pass) { return pass == 2 ? 10 : 11; }
__host__ __device__ constexpr int radix_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
constexpr int MAX_BINS = 2048;

struct SelShared {
    uint32_t hist[3][MAX_BINS];   // this CTA's digit histograms, one array per pass (read remotely after each pass)
    uint32_t merged[MAX_BINS];    // cluster-wide histogram of the current pass (written by the CTAs that own the bins)
    uint32_t wg[32], we[32];      // per warp: keys > T, keys == T in the warp's part of the slice
    uint32_t n_pos, n_gt, n_eq;   // this CTA's slice: positives, keys > T, keys == T (read remotely)
    uint32_t prefix, k_rem, k_eff;   // selection state; every CTA derives the same values
    uint32_t rem_s, rstar;
    uint32_t wt[33];
};

__device__ __forceinline__ uint32_t sortable(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// exclusive scan over the 1024 threads of the CTA (thread order)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t& total) {
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

// Keys of this CTA's slice from the staged copy in shared memory, eight per thread and round (slot o = thread + 1024 (8 round
// + u)).  Slots past the end of the slice come back as invalid (bit u of the mask clear).
__device__ __forceinline__ uint32_t load_keys8(const uint32_t* keys, int len, int round, uint32_t (&k)[8]) {
    uint32_t valid = 0;
    const int o0 = threadIdx.x + SEL_THREADS * 8 * round;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int o = o0 + SEL_THREADS * u;
        k[u] = o < len ? keys[o] : 0u;
        valid |= (o < len ? 1u : 0u) << u;
    }
    return valid;
}

// One histogram pass over this CTA's slice.  Keys whose decided bits differ from the prefix are skipped.  In pass 0 equal
// digits inside a warp are combined first (uniform draws in [0, 1) put half of the keys into four bins).
template <int PASS>
__device__ __forceinline__ void hist_pass(const uint32_t* keys, int len, int rounds, SelShared& sh) {
    constexpr uint32_t hi_mask = PASS == 0 ? 0u : (PASS == 1 ? 0xFFE00000u : 0xFFFFFC00u);
    constexpr uint32_t dmask = (1u << radix_bits(PASS)) - 1;
    const uint32_t prefix = PASS ? sh.prefix : 0u;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < rounds; ++r) {                            // warp-uniform trip count
        uint32_t k[8];
        const uint32_t valid = load_keys8(keys, len, r, k);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = ((valid >> u) & 1u) && (k[u] & hi_mask) == prefix;
            const uint32_t dg = (k[u] >> radix_shift(PASS)) & dmask;
            if (PASS == 0) {
                const uint32_t act = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const uint32_t same = __match_any_sync(act, dg);
                    if (lane == __ffs(same) - 1) atomicAdd(&sh.hist[PASS][dg], static_cast<uint32_t>(__popc(same)));
                }
            } else if (in) {
                atomicAdd(&sh.hist[PASS][dg], 1u);
            }
        }
    }
}

// Merge the eight histograms of this pass and pick its digit: walking the bins from the top down until the running count
// reaches k_rem stops at b* = max{ b >= 1 : I(b) >= k_rem } (0 if there is none), I(b) = sum of hist[j] over j >= b, and
// leaves k_rem - (I(b*) - hist[b*]).  I is an inclusive scan over the bins in reversed order; the one reversed position
// where it crosses k_rem is the pick.  Every CTA of the cluster runs the pick on the same merged histogram.
// (Merging by letting every CTA read all eight histograms cost 5 us per pass: 64 KB through DSMEM per CTA.)
template <int PASS, int SEL_CLUSTER>
__device__ __forceinline__ void merge_and_pick(cg::cluster_group& cluster, SelShared& sh, int rank, int num_sample, int nl) {
    constexpr int bins = 1 << radix_bits(PASS);
    constexpr int per = (bins + SEL_THREADS - 1) / SEL_THREADS;
    constexpr int own = bins / SEL_CLUSTER;
    if (threadIdx.x < own) {
        const int bin = rank * own + threadIdx.x;
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < SEL_CLUSTER; ++q) c += *cluster.map_shared_rank(&sh.hist[PASS][bin], q);
#pragma unroll
        for (int q = 0; q < SEL_CLUSTER; ++q) *cluster.map_shared_rank(&sh.merged[bin], q) = c;
    }
    if (threadIdx.x == SEL_THREADS - 1) {
        if (PASS == 0) {
            uint32_t np = 0;
#pragma unroll
            for (int q = 0; q < SEL_CLUSTER; ++q) np += *cluster.map_shared_rank(&sh.n_pos, q);
            uint32_t k = np > static_cast<uint32_t>(num_sample) ? np : static_cast<uint32_t>(num_sample);
            if (k > static_cast<uint32_t>(nl)) k = nl;
            sh.k_eff = k;
            sh.k_rem = k;
            sh.prefix = 0;
            sh.rem_s = k;
        } else {
            sh.rem_s = sh.k_rem;
        }
        sh.rstar = bins - 1;                  // reversed position of bin 0: the walk's default
    }
    cluster.sync();
    const uint32_t rem = sh.rem_s;
    if (rem == 0) {                           // nothing to select: threshold above every key (CTA-uniform branch)
        if (threadIdx.x == 0) sh.prefix = 0xFFFFFFFFu;
        __syncthreads();
        return;
    }
    uint32_t v[per], sum = 0;
#pragma unroll
    for (int u = 0; u < per; ++u) {
        const int r = threadIdx.x * per + u;  // reversed bin index: r = 0 is the top bin
        const uint32_t c = r < bins ? sh.merged[bins - 1 - r] : 0u;
        v[u] = c;
        sum += c;
    }
    uint32_t tot;
    uint32_t run = block_excl_scan(sum, sh.wt, tot);
    uint32_t in_r = 0, c_r = 0;
    bool mine = false;
#pragma unroll
    for (int u = 0; u < per; ++u) {
        const int r = threadIdx.x * per + u;
        const uint32_t before = run;
        run += v[u];
        // the crossing bin (before < rem <= run) or, if the count never reaches rem above bin 0, bin 0 itself
        if (r < bins && ((r < bins - 1 && before < rem && run >= rem) || (r == bins - 1 && before < rem))) {
            mine = true;
            sh.rstar = r;
            in_r = run;
            c_r = v[u];
        }
    }
    if (mine) {
        sh.prefix |= static_cast<uint32_t>(bins - 1 - sh.rstar) << radix_shift(PASS);
        sh.k_rem = rem - (in_r - c_r);
    }
    __syncthreads();
}

template <int SEL_CLUSTER>                // cluster size: set by the launch attribute
__global__ void __launch_bounds__(SEL_THREADS)
sample_cluster_kernel(const float* __restrict__ perm, const int32_t* __restrict__ labels, int B, int nl, int num_sample,
                      int S, int64_t* __restrict__ index_out, int32_t* slot_of, int32_t* __restrict__ n_out,
                      int32_t* __restrict__ labels_out) {
    extern __shared__ uint32_t dyn[];         // bitmap [S / 32] (bit o = class lo + o is positive), then keys [S]
    __shared__ SelShared sh;
    uint32_t* bitmap = dyn;
    uint32_t* keys = dyn + S / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int lo = min(rank * S, nl), hi = min(lo + S, nl), len = hi - lo;
    const int rounds = (len + SEL_TILE - 1) / SEL_TILE;
    const int lane = threadIdx.x & 31;
    const uint32_t two = sortable(2.0f);

    STAMP(0);
    for (int b = threadIdx.x; b < 3 * MAX_BINS; b += SEL_THREADS) (&sh.hist[0][0])[b] = 0;
    for (int w = threadIdx.x; w < S / 32; w += SEL_THREADS) bitmap[w] = 0;
    if (threadIdx.x == 0) { sh.n_pos = 0; sh.n_gt = 0; sh.n_eq = 0; }
    for (int r = 0; r < rounds; ++r) {                            // stage the slice's keys: eight loads in flight
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int o = threadIdx.x + SEL_THREADS * (8 * r + u);
            f[u] = o < len ? __ldg(perm + lo + o) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int o = threadIdx.x + SEL_THREADS * (8 * r + u);
            if (o < len) keys[o] = sortable(f[u]);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < B; j += SEL_THREADS) {
        const int l = labels[j];
        if (l >= lo && l < hi) {
            atomicOr(&bitmap[(l - lo) >> 5], 1u << ((l - lo) & 31));
            keys[l - lo] = two;                                   // perm[positive] = 2.0 (:111)
        }
    }
    __syncthreads();
    {
        uint32_t np = 0;
        for (int w = threadIdx.x; w < S / 32; w += SEL_THREADS) np += __popc(bitmap[w]);
        np = __reduce_add_sync(0xffffffffu, np);
        if (lane == 0 && np) atomicAdd(&sh.n_pos, np);
    }
    STAMP(1);
    hist_pass<0>(keys, len, rounds, sh);
    STAMP(2);
    cluster.sync();
    STAMP(3);
    merge_and_pick<0, SEL_CLUSTER>(cluster, sh, rank, num_sample, nl);
    STAMP(4);
    hist_pass<1>(keys, len, rounds, sh);
    cluster.sync();
    STAMP(5);
    merge_and_pick<1, SEL_CLUSTER>(cluster, sh, rank, num_sample, nl);
    STAMP(6);
    hist_pass<2>(keys, len, rounds, sh);
    cluster.sync();
    STAMP(7);
    merge_and_pick<2, SEL_CLUSTER>(cluster, sh, rank, num_sample, nl);
    STAMP(8);

    const uint32_t T = sh.prefix;
    const uint32_t need_eq = sh.k_rem;
    const bool none = sh.k_eff == 0;
    // Count and compaction are warp-ordered: warp w owns slots [w S/32, (w+1) S/32) of the slice, lane l the slots
    // l, l + 32, ... of it, so ranks inside a warp come from ballots and only the 32 warp totals need a prefix sum.
    const int wlen = S / 32, warp = threadIdx.x >> 5, wbase = warp * wlen;
    const uint32_t lt_mask = (1u << lane) - 1u;
    {
        uint32_t g = 0, e = 0;
        if (!none) {
            for (int j = 0; j < wlen; j += 128) {                 // four independent loads in flight
                uint32_t k[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int o = wbase + j + 32 * u + lane;
                    k[u] = (j + 32 * u < wlen && o < len) ? keys[o] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int o = wbase + j + 32 * u + lane;
                    const bool ok = j + 32 * u < wlen && o < len;
                    g += ok && k[u] > T;
                    e += ok && k[u] == T;
                }
            }
        }
        g = __reduce_add_sync(0xffffffffu, g);
        e = __reduce_add_sync(0xffffffffu, e);
        if (lane == 0) {
            sh.wg[warp] = g;
            sh.we[warp] = e;
            if (g) atomicAdd(&sh.n_gt, g);
            if (e) atomicAdd(&sh.n_eq, e);
        }
    }
    cluster.sync();
    STAMP(9);
    uint32_t gt_before = 0, eq_before = 0;                        // keys > T / == T before this warp's first slot
#pragma unroll
    for (int q = 0; q < SEL_CLUSTER; ++q) {
        const uint32_t g = *cluster.map_shared_rank(&sh.n_gt, q), e = *cluster.map_shared_rank(&sh.n_eq, q);
        if (q < rank) { gt_before += g; eq_before += e; }
    }
    {
        const uint32_t g = lane < warp ? sh.wg[lane] : 0u, e = lane < warp ? sh.we[lane] : 0u;
        gt_before += __reduce_add_sync(0xffffffffu, g);
        eq_before += __reduce_add_sync(0xffffffffu, e);
    }
    if (rank == 0 && threadIdx.x == 0) n_out[0] = static_cast<int32_t>(sh.k_eff);

    // ordered compaction: ascending index list + slot of every positive
    if (!none) {
        uint32_t eq_run = eq_before;                              // keys == T before the current 32 slots
        uint32_t sel_run = gt_before + min(eq_before, need_eq);   // selected keys before them
        for (int j = 0; j < wlen; j += 128) {
            uint32_t k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = wbase + j + 32 * u + lane;
                k[u] = (j + 32 * u < wlen && o < len) ? keys[o] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = wbase + j + 32 * u + lane;
                const bool ok = j + 32 * u < wlen && o < len;
                const bool eq = ok && k[u] == T;
                const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
                const bool sel = ok && (k[u] > T || (eq && eq_run + __popc(m_eq & lt_mask) < need_eq));
                const uint32_t m_sel = __ballot_sync(0xffffffffu, sel);
                if (sel) {
                    const uint32_t pos = sel_run + __popc(m_sel & lt_mask);
                    index_out[pos] = lo + o;
                    if ((bitmap[o >> 5] >> (o & 31)) & 1u) slot_of[lo + o] = static_cast<int32_t>(pos);
                }
                eq_run += __popc(m_eq);
                sel_run += __popc(m_sel);
            }
        }
    }
    // every positive class has its slot (the barrier's release / acquire orders the global stores of the other CTAs, and no
    // CTA leaves while a peer may still read its shared memory): labels -> position in the index list (= searchsorted, :118)
    STAMP(10);
    cluster.sync();
    STAMP(11);
    const int per_cta = (B + SEL_CLUSTER - 1) / SEL_CLUSTER;
    for (int j = rank * per_cta + threadIdx.x; j < min(B, (rank + 1) * per_cta); j += SEL_THREADS) {
        const int l = labels[j];
        labels_out[j] = l >= 0 ? __ldcg(slot_of + l) : -1;
    }
    STAMP(12);
}

// ---------------------------------------------------------------------------------------------------------------------
// Shards too large for the cluster kernel's shared memory (> 376 k classes per rank, e.g. BASELINE configs[3] on ONE GPU):
// the same selection as one memset + six launches over as many CTAs as the shard has 8192-key tiles.  Every "pick" /
// "scan" / "remap" step that needs the result of a whole grid is run by the LAST CTA of the kernel that produces it
// (atomic ticket + fences).
namespace big {

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

}  // namespace big

// slice per CTA: a multiple of 1024 (warp-uniform loops, whole bitmap words)
static int slice_of(int nl, int cluster) {
    const int s = (nl + cluster - 1) / cluster;
    return (s + SEL_THREADS - 1) / SEL_THREADS * SEL_THREADS;
}

}  // namespace pfc

using namespace pfc;

// workspace: the slot table [num_local] int32 (only the entries of positive classes are ever written or read); shards
// beyond the cluster kernel's reach add the flag array, the selection state and the per-tile counts of the tiled path
static constexpr size_t kMaxDyn = 190u * 1024u;     // + sizeof(SelShared) = 33 KB static: under the 227 KB per CTA
static size_t cluster_dyn_bytes(int num_local, int cluster) {
    const size_t S = static_cast<size_t>(slice_of(num_local, cluster));
    return S / 32 * sizeof(uint32_t) + S * sizeof(uint32_t);
}

// Cluster size for a shard: 16 CTAs (non-portable size, one GPC) when the device can co-schedule them and the shard is
// large enough to use them, else the portable 8; 0 = the shard does not fit the shared memory of the largest cluster.
// The per-key phases (histogram pass 0, compaction) run at ~1 key / clock / SM, so they halve with twice the SMs.
static int g_max_cluster = -1;       // probed once: 16, 8, or 0 (cluster launch unavailable)
static int probe_max_cluster() {
    if (g_max_cluster >= 0) return g_max_cluster;
    g_max_cluster = 0;
    if (cudaFuncSetAttribute(sample_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kMaxDyn)) == cudaSuccess)
        g_max_cluster = 8;
    if (g_max_cluster == 8 &&
        cudaFuncSetAttribute(sample_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kMaxDyn)) == cudaSuccess &&
        cudaFuncSetAttribute(sample_cluster_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(SEL_THREADS);
        cfg.dynamicSmemBytes = kMaxDyn;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 16; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, sample_cluster_kernel<16>, &cfg) == cudaSuccess && n >= 1) g_max_cluster = 16;
    }
    cudaGetLastError();
    return g_max_cluster;
}
static int g_force_cluster = 0;      // pfc_sample_debug_cluster: 0 auto, 8 / 16 forced, -1 tiled path
static int pick_cluster(int num_local) {
    if (g_force_cluster < 0) return 0;
    const int mx = probe_max_cluster();
    // measured (tools/probe/sample_phases.cu): 45 k classes 16.6 us with 8 CTAs / 18.9 with 16; 257 k classes 41.9 / 30.6
    int c = (mx >= 16 && num_local >= 131072) ? 16 : (mx >= 8 ? 8 : 0);
    if (g_force_cluster > 0 && g_force_cluster <= mx) c = g_force_cluster;
    if (c && cluster_dyn_bytes(num_local, c) > kMaxDyn) c = (c == 8 && mx >= 16) ? 16 : 0;
    if (c && cluster_dyn_bytes(num_local, c) > kMaxDyn) c = 0;
    return c;
}

template <int CL>
static int launch_cluster(const float* perm, const int32_t* labels_local, int B, int num_local, int num_sample,
                          int64_t* index_out, int32_t* n_out, int32_t* labels_remapped, int32_t* slot,
                          cudaStream_t stream) {
    int S = slice_of(num_local, CL);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL);
    cfg.blockDim = dim3(SEL_THREADS);
    cfg.dynamicSmemBytes = cluster_dyn_bytes(num_local, CL);
    cfg.stream = stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, sample_cluster_kernel<CL>, perm, labels_local, B, num_local, num_sample, S, index_out,
                              slot, n_out, labels_remapped) == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

extern "C" {

// the tiled path's layout (flags, state, per-tile counts, slot table); the cluster kernel uses the slot table only, so
// one workspace serves whichever path a shard takes
size_t pfc_sample_workspace_bytes(int num_local) { return num_local > 0 ? big::sel_layout(num_local).total : 0; }

int pfc_sample(const float* perm, const int32_t* labels_local, int B, int num_local, int num_sample,
               int64_t* index_out, int32_t* n_out, int32_t* labels_remapped, void* workspace,
               size_t workspace_bytes, void* stream_) {
    if (B <= 0 || num_local <= 0 || num_sample < 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cl = pick_cluster(num_local);
    if (cl) {
        if (workspace_bytes < big::sel_layout(num_local).total) return PFC_ERR_WORKSPACE;
        int32_t* slot = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(workspace) + big::sel_layout(num_local).slot);
        return cl == 16 ? launch_cluster<16>(perm, labels_local, B, num_local, num_sample, index_out, n_out,
                                             labels_remapped, slot, stream)
                        : launch_cluster<8>(perm, labels_local, B, num_local, num_sample, index_out, n_out,
                                            labels_remapped, slot, stream);
    }
    if (workspace_bytes < big::sel_layout(num_local).total) return PFC_ERR_WORKSPACE;
    using namespace big;
    const SelLayout L = sel_layout(num_local);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint8_t* flags = ws + L.flags;
    SelState* st = reinterpret_cast<SelState*>(ws + L.state);
    uint32_t* gt = reinterpret_cast<uint32_t*>(ws + L.gt);
    uint32_t* eq = reinterpret_cast<uint32_t*>(ws + L.eq);
    int32_t* slot = reinterpret_cast<int32_t*>(ws + L.slot);
    // flags and the selection state are contiguous at the front of the workspace
    if (cudaMemsetAsync(ws, 0, L.gt, stream) != cudaSuccess) return PFC_ERR_CUDA;
    mark_positive_kernel<<<(B + 255) / 256, 256, 0, stream>>>(labels_local, B, flags);
    const int hb = L.tiles;
    hist_kernel<0><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);     // + pick of digit 0
    hist_kernel<1><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);
    hist_kernel<2><<<hb, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, num_sample);
    count_kernel<<<L.tiles, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, gt, eq, L.tiles);   // + tile scan
    compact_kernel<<<L.tiles, SEL_THREADS, 0, stream>>>(perm, flags, num_local, st, gt, eq, index_out, slot, n_out,
                                                        labels_local, B, labels_remapped);          // + label remap
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

// kernels pfc_sample launches for a shard of this size (1, or 6 + a memset)
int pfc_sample_launches(int num_local) { return num_local > 0 && pick_cluster(num_local) == 0 ? 6 : 1; }

// tests / A-B: 0 = automatic, 8 or 16 = that cluster size where it fits, -1 = always the tiled path
int pfc_sample_debug_cluster(int mode) {
    g_force_cluster = mode;
    return PFC_OK;
}

}  // extern "C"
