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

// key for "minimum value, first from the top": smaller value wins, then larger threshold
struct Best {
    double v;
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
        y.th = __shfl_xor_sync(0xffffffffu, x.th, o);
        x = better(x, y);
    }
    return x;
}

// One CTA.  Thread t owns a contiguous block of thresholds; cumulative counts above each threshold come from a
// two-level suffix scan, so far/frr are the same integer ratios the reference forms, divided in fp64.
__global__ void __launch_bounds__(1024)
roc_kernel(const unsigned long long* __restrict__ hist_g, const unsigned long long* __restrict__ hist_i, int min_level,
           int max_level, RocOut* __restrict__ out) {
    __shared__ unsigned long long sg[1024], si[1024];
    __shared__ Best sb[32];
    const int T = threadIdx.x;
    constexpr int PER = (HIST_BINS + 1023) / 1024;   // 98
    const int lo = T * PER, hi = min(lo + PER, HIST_BINS);   // bins [lo, hi)
    unsigned long long tg = 0, ti = 0;
    for (int b = lo; b < hi; ++b) { tg += hist_g[b]; ti += hist_i[b]; }
    sg[T] = tg; si[T] = ti;
    __syncthreads();
    // suffix sums over threads (serial by thread 0: 1024 adds)
    __shared__ unsigned long long tot_g, tot_i;
    if (T == 0) {
        unsigned long long ag = 0, ai = 0;
        for (int k = 1023; k >= 0; --k) {
            const unsigned long long g = sg[k], i2 = si[k];
            sg[k] = ag; si[k] = ai;       // counts in bins strictly above thread k's block
            ag += g; ai += i2;
        }
        tot_g = ag; tot_i = ai;
    }
    __syncthreads();
    const double total_g = (double)(long long)tot_g, total_i = (double)(long long)tot_i;
    const int levels = max_level - min_level + 1;
    __shared__ double s_eer[1024];
    // ---- EER: minimum |far - frr|, first from the top (strict <, starting from 1)
    {
        Best best; best.v = 0; best.th = -1;
        double eer_val = 0;
        unsigned long long cg = sg[T], ci = si[T];
        for (int th = hi - 1; th >= lo; --th) {
            if (th >= 1) {
                const double far = (double)(long long)(ci + hist_i[th]) / total_i;
                const double frr = (double)(long long)(tot_g - cg) / total_g;
                const double diff = fabs(far - frr);
                if (diff < 1.0) {
                    Best c; c.v = diff; c.th = th;
                    const Best nb = better(best, c);
                    if (nb.th == th) eer_val = (far + frr) / 2;
                    best = nb;
                }
            }
            cg += hist_g[th];
            ci += hist_i[th];
        }
        s_eer[T] = eer_val;
        const Best w = warp_best(best);
        if ((T & 31) == 0) sb[T >> 5] = w;
        __syncthreads();
        if (T < 32) {
            const Best x = warp_best(sb[T]);
            if (T == 0) {
                out->eer_threshold = x.th < 0 ? 100000 : x.th;
                out->total_genuine = total_g;
                out->total_imposter = total_i;
                out->eer = x.th < 0 ? nan("") : s_eer[x.th / PER];   // owner thread of the winning threshold
            }
        }
    }
    // ---- FRR @ FAR <= 1e-k: minimum frr among thresholds with far <= 1e-k, first from the top
    const double lit[17] = {1e0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10, 1e-11, 1e-12,
                            1e-13, 1e-14, 1e-15, 1e-16};   // the doubles float('1e-k') parses to
    for (int l = 0; l < levels; ++l) {
        const double lim = lit[l + min_level];
        Best best; best.v = 0; best.th = -1;
        unsigned long long cg = sg[T], ci = si[T];
        for (int th = hi - 1; th >= lo; --th) {
            if (th >= 1) {
                const double far = (double)(long long)(ci + hist_i[th]) / total_i;
                if (far <= lim) {
                    Best c; c.v = (double)(long long)(tot_g - cg) / total_g; c.th = th;
                    best = better(best, c);
                }
            }
            cg += hist_g[th];
            ci += hist_i[th];
        }
        __syncthreads();
        const Best w = warp_best(best);
        if ((T & 31) == 0) sb[T >> 5] = w;
        __syncthreads();
        if (T < 32) {
            const Best x = warp_best(sb[T]);
            if (T == 0) {
                out->frr_at[l] = x.th < 0 ? nan("") : x.v;
                out->th_at[l] = x.th;
            }
        }
    }
    for (int l = levels; l < 16; ++l)
        if (T == 0) { out->frr_at[l] = nan(""); out->th_at[l] = -1; }
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

// correct[f][t] = #pairs of fold f classified correctly by "dist < t*step"
__global__ void __launch_bounds__(256)
kfold_count_kernel(const double* __restrict__ dist, const uint8_t* __restrict__ labels, int N, int folds, int n_thr,
                   double step, unsigned int* __restrict__ correct) {
    const int t = blockIdx.x;
    if (t >= n_thr) return;
    __shared__ unsigned int c[64];
    if (threadIdx.x < 64) c[threadIdx.x] = 0;
    __syncthreads();
    const double thr = t * step;          // np.arange(0, 4, 0.01)[t]
    const int base = N / folds, rem = N % folds;   // sklearn KFold: first `rem` folds get one extra
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int cut = rem * (base + 1);
        const int f = i < cut ? i / (base + 1) : rem + (i - cut) / base;
        const bool pred = dist[i] < thr;
        if (pred == (labels[i] != 0)) atomicAdd(&c[f], 1u);
    }
    __syncthreads();
    if (threadIdx.x < folds) correct[threadIdx.x * n_thr + t] = c[threadIdx.x];
}

__global__ void kfold_pick_kernel(const unsigned int* __restrict__ correct, int N, int folds, int n_thr,
                                  double* __restrict__ acc, int* __restrict__ best_idx) {
    const int f = threadIdx.x;
    if (f >= folds) return;
    const int base = N / folds, rem = N % folds;
    const int test_n = base + (f < rem ? 1 : 0);
    const int train_n = N - test_n;
    double best = -1.0;
    int bi = 0;
    for (int t = 0; t < n_thr; ++t) {
        unsigned int tot = 0;
        for (int g = 0; g < folds; ++g)
            if (g != f) tot += correct[g * n_thr + t];
        const double a = (double)tot / (double)train_n;
        if (a > best) { best = a; bi = t; }   // np.argmax: first maximum
    }
    best_idx[f] = bi;
    acc[f] = (double)correct[f * n_thr + bi] / (double)test_n;
}

}  // namespace pfc

using namespace pfc;

extern "C" {

int pfc_eval_hist_bins(void) { return HIST_BINS; }

int fr_pair_score(const float* e1, const float* e2, const uint8_t* labels, int N, int d, double* scores, double* dist,
                  unsigned long long* hist_g, unsigned long long* hist_i, void* stream_) {
    if (N < 0 || d <= 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (cudaMemsetAsync(hist_g, 0, sizeof(unsigned long long) * HIST_BINS, stream) != cudaSuccess) return PFC_ERR_CUDA;
    if (cudaMemsetAsync(hist_i, 0, sizeof(unsigned long long) * HIST_BINS, stream) != cudaSuccess) return PFC_ERR_CUDA;
    if (N == 0) return PFC_OK;
    pair_score_kernel<<<(N + 7) / 8, 256, 0, stream>>>(e1, e2, labels, N, d, scores, dist, hist_g, hist_i);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

// out: 8 + 8 + 16 + 16*8 + 16*4 bytes = struct RocOut (see include/pfc.h: fr_roc_out_t)
int fr_roc(const unsigned long long* hist_g, const unsigned long long* hist_i, int min_level, int max_level, void* out,
           void* stream_) {
    if (max_level < min_level || max_level - min_level + 1 > 16 || min_level < 0 || max_level > 16)
        return PFC_ERR_SHAPE;
    roc_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(hist_g, hist_i, min_level, max_level,
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
    kfold_count_kernel<<<n_thr, 256, 0, stream>>>(dist, labels, N, folds, n_thr, step, correct_ws);
    kfold_pick_kernel<<<1, 64, 0, stream>>>(correct_ws, N, folds, n_thr, acc, best_idx);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

int fr_cross_score(const float* e, const long long* labels, int N, int d, double* scores, double* label_list,
                   unsigned long long* hist_g, unsigned long long* hist_i, void* stream_) {
    if (N < 0 || d <= 0) return PFC_ERR_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (cudaMemsetAsync(hist_g, 0, sizeof(unsigned long long) * HIST_BINS, stream) != cudaSuccess) return PFC_ERR_CUDA;
    if (cudaMemsetAsync(hist_i, 0, sizeof(unsigned long long) * HIST_BINS, stream) != cudaSuccess) return PFC_ERR_CUDA;
    if (N < 2) return PFC_OK;
    const int t = (N + 31) / 32;
    cross_score_kernel<<<dim3(t, t), 1024, 0, stream>>>(e, labels, N, d, scores, label_list, hist_g, hist_i);
    return cudaGetLastError() == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

}  // extern "C"
