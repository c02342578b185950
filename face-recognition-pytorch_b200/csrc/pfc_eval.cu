// Pair-verification scorer (reference: utils/eval.py).
//   pair_score      (:68-99)  score_i = 1 - sum_k (double)(float)(e1[i,k]-e2[i,k])^2 / 4, 100001-bin histograms
//   performance_roc (:7-51)   threshold sweep 100000..1 -> EER threshold, FRR @ FAR = 1e-k
//   performance_acc (:54-66)  accuracy at the EER threshold
//   kfold           standard LFW 10-fold protocol on squared distances (BASELINE cfg-5 wording; not in the reference)
// The subtraction is fp32 and the square / accumulation fp64, exactly what numba generates for
// math.pow(float32 - float32, 2) accumulated into a Python float; an all-fp32 kernel mis-bins ~0.1% of pairs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <cooperative_groups.h>
#include "pfc_internal.h"

namespace pfc {

constexpr int HIST_BINS = 100001;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per pair; lane l handles float4 #l, #l+32, ...
__global__ void __launch_bounds__(256)
pair_score_kernel(const float* __restrict__ e1, const float* __restrict__ e2, const uint8_t* __restrict__ labels, int N,
                  int d, double* __restrict__ scores, double* __restrict__ dist, unsigned long long* __restrict__ hist_g,
                  unsigned long long* __restrict__ hist_i) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const float* a = e1 + static_cast<size_t>(i) * d;
    const float* b = e2 + static_cast<size_t>(i) * d;
    double s = 0.0;
    if ((d & 3) == 0) {
        for (int k = lane; k < (d >> 2); k += 32) {
            const float4 x = *reinterpret_cast<const float4*>(a + 4 * k);
            const float4 y = *reinterpret_cast<const float4*>(b + 4 * k);
            const double d0 = (double)(x.x - y.x), d1 = (double)(x.y - y.y), d2 = (double)(x.z - y.z),
                         d3 = (double)(x.w - y.w);
            s += d0 * d0;
            s += d1 * d1;
            s += d2 * d2;
            s += d3 * d3;
        }
    } else {
        for (int k = lane; k < d; k += 32) {
            const double dd = (double)(a[k] - b[k]);
            s += dd * dd;
        }
    }
    s = warp_sum_d(s);
    if (lane == 0) {
        const double score = 1.0 - s / 4.0;
        scores[i] = score;
        if (dist) dist[i] = s;
        long long idx = (long long)((1e5 - 1.0) * score);   // int(): truncation toward zero
        if (idx < 0) idx += HIST_BINS;                       // numpy negative indexing
        if (idx >= 0 && idx < HIST_BINS) atomicAdd((labels[i] ? hist_g : hist_i) + idx, 1ull);
    }
}

// cross_score (utils/eval.py:102-137): every pair j < i of one embedding set, pair index l = i(i-1)/2 + j.
// One thread per pair inside a 32x32 tile; rows are staged through shared memory in 32-float chunks and each
// thread accumulates its pair in fp64 in k order, i.e. the reference's sequential loop, bit for bit.
__global__ void __launch_bounds__(1024)
cross_score_kernel(const float* __restrict__ e, const long long* __restrict__ labels, int N, int d,
                   double* __restrict__ scores, double* __restrict__ label_list, unsigned long long* __restrict__ hist_g,
                   unsigned long long* __restrict__ hist_i) {
    const int ti0 = blockIdx.y * 32, tj0 = blockIdx.x * 32;
    if (tj0 > ti0) return;                       // only tiles touching the lower triangle
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // tx -> j, ty -> i
    const int i = ti0 + ty, j = tj0 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < d; k0 += 32) {
        // row (ty) of each tile, column k0 + tx
        const int k = k0 + tx;
        sa[ty][tx] = (ti0 + ty < N && k < d) ? e[static_cast<size_t>(ti0 + ty) * d + k] : 0.f;
        sb[ty][tx] = (tj0 + ty < N && k < d) ? e[static_cast<size_t>(tj0 + ty) * d + k] : 0.f;
        __syncthreads();
        const int kn = min(32, d - k0);
        for (int kk = 0; kk < kn; ++kk) {
            const double dd = (double)(sb[tx][kk] - sa[ty][kk]);   // embeddings[j,k] - embeddings[i,k], fp32
            acc += dd * dd;
        }
        __syncthreads();
    }
    if (i < N && j < i) {
        const double score = 1.0 - acc / 4.0;
        const size_t l = static_cast<size_t>(i) * (i - 1) / 2 + j;
        scores[l] = score;
        const bool same = labels[j] == labels[i];
        label_list[l] = same ? 1.0 : 0.0;
        long long idx = (long long)((1e5 - 1.0) * score);
        if (idx < 0) idx += HIST_BINS;
        if (idx >= 0 && idx < HIST_BINS) atomicAdd((same ? hist_g : hist_i) + idx, 1ull);
    }
}

struct RocOut {
    int eer_threshold;
    int pad;
    double eer;
    double total_genuine, total_imposter;
    double frr_at[16];     // NaN when never recorded
    int th_at[16];         // -1 when never recorded
};

// key for "minimum value, first from the top": smaller value wins, then larger threshold; `aux` rides along (the EER
// value of the winning threshold)
struct Best {
    double v;
    double aux;
    int th;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    if (b.th < 0) return a;
    if (a.th < 0) return b;
    if (b.v < a.v || (b.v == a.v && b.th > a.th)) return b;
    return a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.aux = __shfl_xor_sync(0xffffffffu, x.aux, o);
        y.th = __shfl_xor_sync(0xffffffffu, x.th, o);
        x = better(x, y);
    }
    return x;
}

// performance_roc as ONE cluster of 8 CTAs (8 SMs): CTA c stages bins [c S, (c+1) S) of both histograms into shared
// memory with coalesced loads (200 KB), thread t owns 49 consecutive thresholds of the slice.  The counts above each
// thread's run come from a warp-shuffle suffix scan inside the CTA plus the CTA totals exchanged through distributed
// shared memory, so far / frr are the same integer ratios the reference forms, divided in fp64.  ONE sweep over the
// thread's bins serves the EER and every FAR level: far and frr are divided once per threshold (fp64 divisions are the
// cost of this kernel); for "minimum frr with far <= 1e-k" the candidates are compared by their integer numerators
// (frr = num / total_genuine is strictly monotone in num) and only the winner is divided.  The per-CTA winners go to
// CTA 0 through DSMEM.  (Round 1: one CTA, strided global loads, 1 + levels passes over global memory: 767 us.)
__constant__ double ROC_LIT[17] = {1e0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10, 1e-11, 1e-12,
                                   1e-13, 1e-14, 1e-15, 1e-16};   // the doubles float('1e-k') parses to
constexpr int ROC_CLUSTER = 8;
constexpr int ROC_THREADS = 256;
constexpr int ROC_WARPS = ROC_THREADS / 32;
constexpr int ROC_SLICE = (HIST_BINS + ROC_CLUSTER - 1) / ROC_CLUSTER;      // 12501 bins per CTA
constexpr int ROC_PER = (ROC_SLICE + ROC_THREADS - 1) / ROC_THREADS;        // 49 bins per thread (256 threads: the 16 level candidates stay in registers)
constexpr int ROC_SMEM = 2 * ROC_SLICE * 8;
constexpr int ROC_LEVELS = 16;
constexpr unsigned long long ROC_NONE = ~0ull;

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ unsigned long long warp_suffix_excl(unsigned long long v, int lane, unsigned long long* total) {
    // exclusive suffix sum over the warp (sum of the lanes ABOVE this one); *total = sum over the warp
    unsigned long long s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_down_sync(0xffffffffu, s, o);
        if (lane + o < 32) s += y;
    }
    *total = __shfl_sync(0xffffffffu, s, 0);
    return s - v;
}

// candidate for one FAR level: smallest genuine count at or below the threshold (= frr numerator), then larger threshold
struct LevelBest {
    unsigned long long num;     // ROC_NONE: none
    int th;
};
__device__ __forceinline__ LevelBest level_better(LevelBest a, LevelBest b) {
    if (b.num < a.num || (b.num == a.num && b.num != ROC_NONE && b.th > a.th)) return b;
    return a;
}
__device__ __forceinline__ LevelBest warp_level_best(LevelBest x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        LevelBest y;
        y.num = __shfl_xor_sync(0xffffffffu, x.num, o);
        y.th = __shfl_xor_sync(0xffffffffu, x.th, o);
        x = level_better(x, y);
    }
    return x;
}

__global__ void __cluster_dims__(ROC_CLUSTER, 1, 1) __launch_bounds__(ROC_THREADS)
roc_kernel(const unsigned long long* __restrict__ hist_g, const unsigned long long* __restrict__ hist_i, int min_level,
           int max_level, RocOut* __restrict__ out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char roc_smem[];
    unsigned long long* sg = reinterpret_cast<unsigned long long*>(roc_smem);
    unsigned long long* si = sg + ROC_SLICE;
    __shared__ unsigned long long wtot_g[32], wtot_i[32];
    __shared__ unsigned long long cta_g[ROC_CLUSTER], cta_i[ROC_CLUSTER];     // every CTA's totals (written by the peers)
    __shared__ Best wbest[ROC_WARPS];
    __shared__ LevelBest wlevel[ROC_LEVELS][ROC_WARPS];
    __shared__ Best cand_eer[ROC_CLUSTER];                                    // CTA 0: per-CTA winners
    __shared__ LevelBest cand_level[ROC_CLUSTER][ROC_LEVELS];

    const int T = threadIdx.x, lane = T & 31, warp = T >> 5;
    const int c = static_cast<int>(cluster.block_rank());
    const int base = c * ROC_SLICE;
    const int len = max(0, min(ROC_SLICE, HIST_BINS - base));
    // asynchronous 8-byte copies straight into shared memory: all ~100 copies of a thread are in flight at once (with
    // register-staged loads the 200 KB fill was a chain of dependent DRAM round trips: 40 of the kernel's 68 us)
    for (int k = T; k < len; k += ROC_THREADS) {
        cp_async_8(sg + k, hist_g + base + k);
        cp_async_8(si + k, hist_i + base + k);
    }
    if (T < 32) { wtot_g[T] = 0; wtot_i[T] = 0; }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int lo = min(T * ROC_PER, len), hi = min(lo + ROC_PER, len);        // local bins [lo, hi)
    unsigned long long tg = 0, ti = 0;
    for (int b = lo; b < hi; ++b) { tg += sg[b]; ti += si[b]; }
    unsigned long long wg, wi;
    const unsigned long long ag = warp_suffix_excl(tg, lane, &wg);            // in-warp counts above this thread's run
    const unsigned long long ai = warp_suffix_excl(ti, lane, &wi);
    if (lane == 0) { wtot_g[warp] = wg; wtot_i[warp] = wi; }
    __syncthreads();
    if (warp == 0) {
        unsigned long long tw_g, tw_i;
        const unsigned long long vg = wtot_g[lane], vi = wtot_i[lane];        // zero beyond the CTA's warps
        const unsigned long long eg = warp_suffix_excl(vg, lane, &tw_g);
        const unsigned long long ei = warp_suffix_excl(vi, lane, &tw_i);
        wtot_g[lane] = eg; wtot_i[lane] = ei;                                 // now: counts in the warps above
        if (lane < ROC_CLUSTER) {                                             // lane r publishes this CTA's totals to CTA r
            *cluster.map_shared_rank(&cta_g[c], lane) = tw_g;
            *cluster.map_shared_rank(&cta_i[c], lane) = tw_i;
        }
    }
    cluster.sync();
    unsigned long long tot_g = 0, tot_i = 0, up_g = 0, up_i = 0;
#pragma unroll
    for (int r = 0; r < ROC_CLUSTER; ++r) {
        tot_g += cta_g[r]; tot_i += cta_i[r];
        if (r > c) { up_g += cta_g[r]; up_i += cta_i[r]; }
    }
    unsigned long long cgv = up_g + wtot_g[warp] + ag;                        // counts in bins strictly above this run
    unsigned long long civ = up_i + wtot_i[warp] + ai;
    const double total_g = (double)(long long)tot_g, total_i = (double)(long long)tot_i;
    const int levels = max_level - min_level + 1;
    double lim[ROC_LEVELS];
#pragma unroll
    for (int l = 0; l < ROC_LEVELS; ++l) lim[l] = l < levels ? ROC_LIT[l + min_level] : -1.0;   // far >= 0: never met
    Best eer; eer.v = 0; eer.aux = 0; eer.th = -1;
    LevelBest lb[ROC_LEVELS];
#pragma unroll
    for (int l = 0; l < ROC_LEVELS; ++l) { lb[l].num = ROC_NONE; lb[l].th = -1; }
    for (int b = hi - 1; b >= lo; --b) {                                      // thresholds from the top down
        const int th = base + b;
        const unsigned long long hgb = sg[b], hib = si[b];
        if (th >= 1) {
            const unsigned long long num = tot_g - cgv;                      // genuine pairs at or below th
            const double far = (double)(long long)(civ + hib) / total_i;
            const double frr = (double)(long long)num / total_g;
            Best cnd; cnd.th = th; cnd.v = fabs(far - frr); cnd.aux = (far + frr) / 2;
            if (cnd.v < 1.0) eer = better(eer, cnd);                          // strict '<' from the top: first minimum
#pragma unroll
            for (int l = 0; l < ROC_LEVELS; ++l) {
                if (far <= lim[l]) {
                    LevelBest q; q.num = num; q.th = th;
                    lb[l] = level_better(lb[l], q);
                }
            }
        }
        cgv += hgb; civ += hib;
    }
    {
        const Best w = warp_best(eer);
        if (lane == 0) wbest[warp] = w;
#pragma unroll
        for (int l = 0; l < ROC_LEVELS; ++l) {
            const LevelBest x = warp_level_best(lb[l]);
            if (lane == 0) wlevel[l][warp] = x;
        }
    }
    __syncthreads();
    if (warp == 0) {
        Best x; x.v = 0; x.aux = 0; x.th = -1;
        if (lane < ROC_WARPS) x = wbest[lane];
        x = warp_best(x);
        if (lane == 0) *cluster.map_shared_rank(&cand_eer[c], 0) = x;
        for (int l = 0; l < levels; ++l) {
            LevelBest y; y.num = ROC_NONE; y.th = -1;
            if (lane < ROC_WARPS) y = wlevel[l][lane];
            y = warp_level_best(y);
            if (lane == 0) *cluster.map_shared_rank(&cand_level[c][l], 0) = y;
        }
    }
    cluster.sync();
    if (c == 0 && warp == 0) {
        if (lane == 0) {
            Best x = cand_eer[0];
            for (int r = 1; r < ROC_CLUSTER; ++r) x = better(x, cand_eer[r]);
            out->eer_threshold = x.th < 0 ? 100000 : x.th;
            out->pad = 0;
            out->eer = x.th < 0 ? nan("") : x.aux;
            out->total_genuine = total_g;
            out->total_imposter = total_i;
        }
        if (lane < ROC_LEVELS) {
            LevelBest y; y.num = ROC_NONE; y.th = -1;
            if (lane < levels) {
                y = cand_level[0][lane];
                for (int r = 1; r < ROC_CLUSTER; ++r) y = level_better(y, cand_level[r][lane]);
            }
            out->frr_at[lane] = y.num == ROC_NONE ? nan("") : (double)(long long)y.num / total_g;
            out->th_at[lane] = y.num == ROC_NONE ? -1 : y.th;
        }
    }
}

__global__ void __launch_bounds__(256)
acc_kernel(const double* __restrict__ scores, const uint8_t* __restrict__ labels, int N, double thd,
           unsigned long long* __restrict__ fr_fa) {
    unsigned int fr = 0, fa = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double s = scores[i];
        const bool l = labels[i] == 1;
        fr += (s <= thd) && l;
        fa += (s > thd) && (labels[i] == 0);
    }
    fr = __reduce_add_sync(0xffffffffu, fr);
    fa = __reduce_add_sync(0xffffffffu, fa);
    if ((threadIdx.x & 31) == 0) {
        if (fr) atomicAdd(&fr_fa[0], (unsigned long long)fr);
        if (fa) atomicAdd(&fr_fa[1], (unsigned long long)fa);
    }
}

// Standard LFW k-fold protocol in ONE single-CTA launch.  A pair is classified "same" by threshold t iff
// dist < t*step, so its verdict flips exactly once along the sweep: tmin = the first t with dist < t*step (n_thr when
// there is none).  A genuine pair is counted correct for t >= tmin, an imposter pair for t < tmin -- one shared-memory
// histogram entry per pair and fold instead of n_thr CTAs each re-reading every distance (round 1: 226 us):
//   correct[f][t] = sum_{u <= t} Hg[f][u] + sum_{u > t} Hi[f][u].
// The comparison that decides tmin is the very `dist < t*step` in fp64 (t*step is monotone in t), so the counts are
// the ones a threshold-by-threshold sweep would produce.
constexpr int KF_THREADS = 1024;

__global__ void __launch_bounds__(KF_THREADS)
kfold_kernel(const double* __restrict__ dist, const uint8_t* __restrict__ labels, int N, int folds, int n_thr, double step,
             unsigned int* __restrict__ correct, double* __restrict__ acc, int* __restrict__ best_idx) {
    extern __shared__ unsigned int kf_smem[];
    const int S = n_thr + 1;
    unsigned int* hg = kf_smem;                 // [folds][S]  -> inclusive prefix -> correct[f][t]
    unsigned int* hi = hg + folds * S;          // [folds][S]
    unsigned int* tot = hi + folds * S;         // [n_thr] correct pairs over all folds
    const int T = threadIdx.x, lane = T & 31, warp = T >> 5, warps = KF_THREADS / 32;
    for (int k = T; k < 2 * folds * S + n_thr; k += KF_THREADS) kf_smem[k] = 0;
    __syncthreads();
    const int base = N / folds, rem = N % folds;   // sklearn KFold: first `rem` folds get one extra
    const int cut = rem * (base + 1);
    for (int i = T; i < N; i += KF_THREADS) {
        const int f = i < cut ? i / (base + 1) : rem + (i - cut) / base;
        const double dv = dist[i];
        int t = n_thr;
        if (dv == dv) {
            const double q = floor(dv / step);
            t = q < 0.0 ? 0 : (q > (double)n_thr ? n_thr : (int)q);
            while (t > 0 && dv < (double)(t - 1) * step) --t;
            while (t < n_thr && !(dv < (double)t * step)) ++t;
        }
        atomicAdd((labels[i] != 0 ? hg : hi) + f * S + t, 1u);
    }
    __syncthreads();
    // per fold: inclusive prefix sums of both histograms (one warp per fold, 32 thresholds per trip)
    for (int f = warp; f < folds; f += warps) {
        unsigned int cg = 0, ci = 0;
        for (int t0 = 0; t0 < S; t0 += 32) {
            const int t = t0 + lane;
            unsigned int vg = t < S ? hg[f * S + t] : 0u, vi = t < S ? hi[f * S + t] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int yg = __shfl_up_sync(0xffffffffu, vg, o), yi = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) { vg += yg; vi += yi; }
            }
            vg += cg; vi += ci;
            if (t < S) { hg[f * S + t] = vg; hi[f * S + t] = vi; }
            cg = __shfl_sync(0xffffffffu, vg, 31);
            ci = __shfl_sync(0xffffffffu, vi, 31);
        }
        // cg / ci: pairs of the fold by label; correct[f][t] = prefix_g[t] + (ci - prefix_i[t])
        for (int t = lane; t < n_thr; t += 32) {
            const unsigned int cr = hg[f * S + t] + (ci - hi[f * S + t]);
            hg[f * S + t] = cr;
            if (correct) correct[f * n_thr + t] = cr;
            atomicAdd(&tot[t], cr);
        }
    }
    __syncthreads();
    // per fold: best training threshold (np.argmax: the first maximum), applied to the held-out fold
    for (int f = warp; f < folds; f += warps) {
        const int test_n = base + (f < rem ? 1 : 0);
        const int train_n = N - test_n;
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int t = lane; t < n_thr; t += 32) {
            const double a = (double)(tot[t] - hg[f * S + t]) / (double)train_n;
            if (a > best) { best = a; bi = t; }        // ascending t per lane: keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            if (bi == 0x7fffffff) bi = 0;
            best_idx[f] = bi;
            acc[f] = (double)hg[f * S + bi] / (double)test_n;
        }
    }
}

}  // namespace pfc

using namespace pfc;

// both histograms zeroed by one memset node when the caller laid them out back to back (eval.py does)
static int zero_histograms(unsigned long long* hist_g, unsigned long long* hist_i, cudaStream_t stream) {
    const size_t bytes = sizeof(unsigned long long) * HIST_BINS;
    if (hist_i == hist_g + HIST_BINS)
        return cudaMemsetAsync(hist_g, 0, 2 * bytes, stream) != cudaSuccess;
    return cudaMemsetAsync(hist_g, 0, bytes, stream) != cudaSuccess || cudaMemsetAsync(hist_i, 0, bytes, stream) != cudaSuccess;
}

extern "C" {

int pfc_eval_hist_bins(void) { return HIST_BINS; }

int fr_pair_score(const float* e1, const float* e2, const uint8_t* labels, int N, int d, double* scores, double* dist,
                  unsigned long long* hist_g, unsigned long long* hist_i, void* stream_) {
    if (N < 0 || d <= 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (zero_histograms(hist_g, hist_i, stream)) return PFC_ERR_CUDA;
    if (N == 0) return PFC_OK;
    pair_score_kernel<<<(N + 7) / 8, 256, 0, stream>>>(e1, e2, labels, N, d, scores, dist, hist_g, hist_i);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

// out: 8 + 8 + 16 + 16*8 + 16*4 bytes = struct RocOut (see include/pfc.h: fr_roc_out_t)
int fr_roc(const unsigned long long* hist_g, const unsigned long long* hist_i, int min_level, int max_level, void* out,
           void* stream_) {
    if (max_level < min_level || max_level - min_level + 1 > 16 || min_level < 0 || max_level > 16)
        return PFC_ERR_SHAPE;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(roc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ROC_SMEM) != cudaSuccess)
            return PFC_ERR_CUDA;
        attr_set = true;
    }
    // one cluster of ROC_CLUSTER CTAs (compile-time __cluster_dims__)
    roc_kernel<<<ROC_CLUSTER, ROC_THREADS, ROC_SMEM, (cudaStream_t)stream_>>>(hist_g, hist_i, min_level, max_level,
                                                                            reinterpret_cast<RocOut*>(out));
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

int fr_acc_counts(const double* scores, const uint8_t* labels, int N, double threshold, unsigned long long* fr_fa,
                  void* stream_) {
    if (N < 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (cudaMemsetAsync(fr_fa, 0, 2 * sizeof(unsigned long long), stream) != cudaSuccess) return PFC_ERR_CUDA;
    if (N == 0) return PFC_OK;
    int grid = (N + 255) / 256;
    if (grid > 592) grid = 592;
    acc_kernel<<<grid, 256, 0, stream>>>(scores, labels, N, threshold, fr_fa);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

int fr_kfold_acc(const double* dist, const uint8_t* labels, int N, int folds, int n_thr, double step,
                 unsigned int* correct_ws, double* acc, int* best_idx, void* stream_) {
    if (N <= 0 || folds < 2 || folds > 64 || n_thr <= 0 || N < folds) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t smem = (static_cast<size_t>(2) * folds * (n_thr + 1) + n_thr) * sizeof(unsigned int);
    if (smem > 227 * 1024) return PFC_ERR_SHAPE;          // folds x thresholds beyond one CTA's shared memory
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return PFC_ERR_CUDA;
        attr_set = true;
    }
    kfold_kernel<<<1, KF_THREADS, smem, stream>>>(dist, labels, N, folds, n_thr, step, correct_ws, acc, best_idx);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

int fr_cross_score(const float* e, const long long* labels, int N, int d, double* scores, double* label_list,
                   unsigned long long* hist_g, unsigned long long* hist_i, void* stream_) {
    if (N < 0 || d <= 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (zero_histograms(hist_g, hist_i, stream)) return PFC_ERR_CUDA;
    if (N < 2) return PFC_OK;
    const int t = (N + 31) / 32;
    cross_score_kernel<<<dim3(t, t), 1024, 0, stream>>>(e, labels, N, d, scores, label_list, hist_g, hist_i);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

}  // extern "C"
