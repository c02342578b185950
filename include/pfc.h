/* libpfc_b200 -- C ABI of the B200-native margin-softmax head (ArcFace + PartialFC) and pair-verification scorer.
 *
 * This is the drop-in boundary: plain pointers and sizes, caller-owned device memory, an explicit CUDA stream,
 * no allocation, no host synchronisation, no exceptions.  Every function returns PFC_OK (0) or a negative
 * PFC_ERR_* code; nothing is launched when an error is returned.  All pointers are DEVICE pointers unless
 * stated otherwise; `stream` is a cudaStream_t (NULL = default stream).  Matrices are row-major with the
 * embedding dimension d contiguous; d must be a multiple of 8 and <= 1024; bf16 buffers are 16-byte aligned.
 *
 * fp16_operands (the entry points that produce or consume the normalised GEMM operands Xn / Wn): 0 = bf16, the default
 * of this library (bf16-in / fp32-accumulate); 1 = fp16, what the reference's AMP mode multiplies
 * (torch.cuda.amp.autocast around the logits, nets/PartialFC.py:198; conf.mixed_precision).  Same buffer sizes, same
 * kernels: only the element format of the `xn` / `wn` arguments (named *_bf16 below) and the operand formats of the
 * tcgen05 instruction change.  The spill E' and the scaled rows Xs are bf16 in both modes (their exponent range is
 * needed), and tcgen05 kind::f16 takes no mixed bf16 x fp16 pair, so pfc_backward_dx wants a bf16 copy of an fp16 shard:
 * pfc_cast_f16_to_bf16 (elems a multiple of 8, 16-byte aligned buffers).  All calls of one step must agree.
 *
 * The reference is pure Python (aanna0701/face-recognition-pytorch); each entry point names the reference
 * code it replaces as file:line.  The Python binding a maintainer would add is shown in INTEGRATION.md.
 */
#ifndef PFC_B200_H
#define PFC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFC_OK 0
#define PFC_ERR_CUDA (-1)        /* a CUDA runtime call failed */
#define PFC_ERR_LAUNCH (-2)      /* kernel launch failed */
#define PFC_ERR_SHAPE (-3)       /* invalid / unsupported shape argument */
#define PFC_ERR_ALIGNMENT (-4)   /* pointer or stride not 16-byte aligned */
#define PFC_ERR_DRIVER (-5)      /* cuTensorMapEncodeTiled not available from the driver */
#define PFC_ERR_TENSORMAP (-6)   /* tensor-map encoding rejected */
#define PFC_ERR_SCALE_RANGE (-7) /* logit scale s outside the fixed-shift exponent range (s <= ~65.8) */
#define PFC_ERR_WORKSPACE (-8)   /* workspace too small */

#define PFC_MARGIN_ARCFACE 0     /* nets/ArcFace.py:63-91 and CombinedMarginLoss with m1 == 1, m3 == 0 (:42-52) */
#define PFC_MARGIN_COSFACE 1     /* nets/ArcFace.py:94-106 and CombinedMarginLoss with m3 > 0 (:54-57) */

int pfc_version(void);
const char* pfc_error_string(int code);           /* host string */

/* ---- shape helpers (host only, no CUDA calls except pfc_dx_splits' SM-count query) ---------------------- */
int pfc_exp_top(void);                            /* exponent offset of the spilled e terms (see pfc_forward) */
int pfc_padded_classes(int n);                    /* row stride (elements) of the E' spill for n active classes */
int pfc_padded_batch(int B);                      /* row count the part_sum slabs are padded to */
int pfc_num_class_tiles(int n);                   /* number of part_sum slabs (column groups) = leading dim of part_sum */
int pfc_part_sum_cols(void);                      /* classes per part_sum slab (64) */
int pfc_dx_splits(int B, int n, int d);           /* class splits pfc_backward_dx will use for this shape */
int pfc_dx_max_splits(int B, int d);              /* upper bound of pfc_dx_splits over all n (sizes `partial`) */

/* ---- (1) fused L2 normalise: F.normalize of embeddings / of the classifier shard, nets/PartialFC.py:199-200.
 * xn[r,:] = bf16(x[src,:] / max(||x[src,:]||, 1e-12)), inv_norm[r] = 1/max(||.||, 1e-12), src = index ? index[r] : r
 * (the optional gather is the `self.weight[self.weight_index]` of nets/PartialFC.py:120 fused in). */
int pfc_cast_f16_to_bf16(const void* src_f16, void* dst_bf16, size_t elems, void* stream);
int pfc_l2norm_rows(const float* x, const int64_t* index, int rows, int d, void* xn_bf16, float* inv_norm, int fp16_operands,
                    void* stream);

/* ---- label localisation, nets/PartialFC.py:188-193: out[i] = labels[i]-class_start if owned by this rank else -1 */
int pfc_localize_labels(const int64_t* labels, int B, int64_t class_start, int num_local, int32_t* labels_local,
                        void* stream);

/* ---- (2) negative-class sampling, nets/PartialFC.py:92-121 (sample()).
 * perm: the rank's uniform draw [num_local] (the reference draws it with torch.rand on the CPU generator, :110).
 * index_out [max(num_sample, #positives)] ascending int64 (== self.weight_index), n_out[0] its length,
 * labels_remapped[i] = searchsorted(index_out, labels_local[i]) for owned rows, -1 otherwise (:118).
 * Ties at the k-th value are resolved lowest-index-first.  One launch (a thread-block cluster of 16 CTAs, 8 for
 * small shards or where 16 cannot be co-scheduled) for shards of up to ~750 k classes per rank; larger shards take
 * one memset + six launches (pfc_sample_launches tells which).  pfc_sample_debug_cluster: 0 = automatic, 8 / 16 =
 * that cluster size where it fits, -1 = always the tiled six-launch path (tests, A/B measurements). */
size_t pfc_sample_workspace_bytes(int num_local);
int pfc_sample_launches(int num_local);
int pfc_sample_debug_cluster(int mode);
int pfc_sample(const float* perm, const int32_t* labels_local, int B, int num_local, int num_sample,
               int64_t* index_out, int32_t* n_out, int32_t* labels_remapped, void* workspace,
               size_t workspace_bytes, void* stream);
/* HOST function (no CUDA call): the next n float32 values `torch.rand(n)` would draw from a CPU torch.Generator, produced
 * in bulk from the generator's state blob (torch.Generator.get_state(): MT19937, 5056 bytes), which is advanced in place
 * exactly as torch would leave it -- the sampling draw of nets/PartialFC.py:110 without torch's per-element generator
 * call.  state and out are HOST pointers (out may be pinned).  PFC_ERR_SHAPE: not a seeded CPU MT19937 state blob. */
size_t pfc_host_mt19937_state_bytes(void);
int pfc_host_mt19937_uniform(uint8_t* state, size_t state_bytes, float* out, size_t n);


/* rows of up to 3 fp32 matrices at once: dst[t][r,:] = src[t][index[r],:]  (nets/PartialFC.py:120-121)
 * and dst[t][index[r],:] = src[t][r,:] (update(), nets/PartialFC.py:133-143).  src/dst are HOST arrays of device
 * pointers. */
int pfc_gather_rows(const float* const* src, float* const* dst, int count, const int64_t* index, int rows, int d,
                    void* stream);
int pfc_scatter_rows(const float* const* src, float* const* dst, int count, const int64_t* index, int rows, int d,
                     void* stream);

/* ---- (3) forward: cosine-logit GEMM + clamp + margin + scale + softmax terms, nets/PartialFC.py:201-207,
 * nets/ArcFace.py:76-91 (or :100-105), nets/PartialFC.py:446-458.  tcgen05 GEMM Xn[B,d] . Wn[n,d]^T whose epilogue
 * never writes logits: for every (sample i, class c) it forms e_ic = 2^(log2e*(z_ic - s) + pfc_exp_top()) with
 * z = s*clamp(cos,-1,1) (margin applied on the target column), accumulates the per-row sum of the NON-target terms
 * per part_sum slab (pfc_part_sum_cols() classes) into part_sum[slab][i] and spills e (zeroed where the inter-class
 * filter fires; the clamp's gradient gate is applied on the target column by pfc_backward_prepare) as bf16 into the
 * CLASS-BLOCKED array E[(c / 64) * B + i][c % 64] (n_pad = pfc_padded_classes(n), a multiple of 64; B * n_pad
 * elements): every 64-class column block is contiguous over the samples, so the forward's stores and both gradient
 * GEMMs' loads are multi-KB contiguous pieces.  E is opaque to callers: only pfc_backward_* read it, with the same
 * B and n_pad.  For rows whose target class is local it also writes the raw target cosine, the target's e term
 * and the target logit.  labels_local: -1 = target on another rank. */
int pfc_forward(const void* xn_bf16, const void* wn_bf16, const int32_t* labels_local, int B, int n, int d, float s,
                int margin_kind, float m2, float m3, float interclass_filtering_threshold, void* E_bf16, int n_pad,
                float* part_sum, float* tgt_raw, float* tgt_e, float* tgt_z, int fp16_operands, void* stream);

/* ---- stand-alone margin module, nets/ArcFace.py:76-91 (ArcFace.forward), :100-105 (CosFace), :27-61 (Combined):
 * out[i,c] = s * margin(logits[i,c]) on the target column labels[i] (int64, -1 = none), s * logits elsewhere
 * (0 where the inter-class filter fires); gate (nullable) = d out / d logits for the backward. */
int pfc_margin_apply(const float* logits, const int64_t* labels, int B, int n, int margin_kind, float s, float m2,
                     float m3, float interclass_filtering_threshold, float* out, float* gate, void* stream);

/* ---- (4) row statistics and loss, nets/PartialFC.py:446-461 (DistCrossEntropyFunc.forward).
 * pfc_row_stats: stats[i] = { sum_tiles part_sum[.][i], target e or 0 } -- the [B,2] array ranks all-reduce (SUM),
 * replacing the reference's three all-reduces (:448, :453, :459).
 * pfc_loss: row_L[i] = stats[i][0]+stats[i][1]; loss[0] = -mean_i log(max(stats[i][1]/row_L[i], 1e-30)). */
int pfc_row_stats(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                  float* stats, void* stream);
int pfc_loss(const float* stats, int B, float* row_L, float* loss, void* stream);
/* One-rank shortcuts (no exchange between the two halves): pfc_row_stats + pfc_loss in one launch (ticket -> a uint32
 * zeroed once; the last CTA through forms the loss, same bits as the two calls), and pfc_l2norm_rows +
 * pfc_localize_labels of the same rows in one launch. */
int pfc_row_stats_loss(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                       float* stats, float* row_L, float* loss, unsigned int* ticket, void* stream);
int pfc_l2norm_rows_localize(const float* x, int rows, int d, void* xn_bf16, float* inv_norm, const int64_t* labels,
                             int64_t class_start, int num_local, int32_t* labels_local, int fp16_operands,
                             void* stream);
/* pfc_row_stats_loss + pfc_backward_prepare (below) in one launch, for a step driven without autograd (d loss known when
 * the forward ends; model/FR_PartialFC.py:175-184 as one call): every CTA forms the coefficients of the rows it summed. */
int pfc_row_stats_loss_prepare(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                               float* stats, float* row_L, float* loss, unsigned int* ticket, const float* grad_loss,
                               float s, int d, const float* tgt_raw, int margin_kind, float m2, const void* xn_bf16,
                               void* xs_bf16, float* coef, void* E_bf16, int n_pad, int fp16_operands, void* stream);

/* ---- (5) backward, nets/PartialFC.py:464-484 (DistCrossEntropyFunc.backward) + autograd of :199-206.
 * pfc_backward_prepare: coef[i] = g*s/(B*row_L[i]) (g = grad_loss[0], device scalar, NULL = 1), xs = bf16(coef*xn),
 *   and the target column of E is patched to -dm_i*mask_i*stats[i][0] (dm = margin derivative, mask = clamp gate).
 * pfc_backward_dx: partial[z] = E[:, split z] . Wn[split z, :]   (tcgen05, `splits` = pfc_dx_splits(B,n,d) slabs [B,d])
 * pfc_dx_finalize: out = scale * normalize_backward(coef * sum_z partial[z]); with x == NULL only the scaled sum is
 *   formed (the array ranks reduce-scatter, replacing AllGatherFunc.backward :505-522; `scale` = world size, :521).
 * pfc_backward_dw: dwn[n,d] = E^T . xs   (tcgen05)
 * pfc_dw_finalize: dw = normalize_backward(dwn) * inv_grad_scale (un-fused: hand dw to any torch optimizer)
 * pfc_dw_sgd / pfc_dw_adam: the same plus the optimizer step fused (torch.optim.SGD / Adam / AdamW semantics,
 *   driven from model/FR_PartialFC.py:182-188), also emitting the next step's bf16 normalised shard. */
int pfc_backward_prepare(const float* stats, const float* row_L, const float* grad_loss, float s, int B, int d,
                         const int32_t* labels_local, const float* tgt_raw, int margin_kind, float m2,
                         const void* xn_bf16, void* xs_bf16, float* coef, void* E_bf16, int n_pad, int fp16_operands,
                         void* stream);
int pfc_backward_dx(const void* E_bf16, int n_pad, const void* wn_bf16, int B, int n, int d, float* partial,
                    int splits, void* stream);
int pfc_dx_finalize(const float* partial, int splits, const float* coef, const float* x, const float* inv_norm,
                    float scale, int rows, int rows_total, int d, float* out, void* stream);
/* dwn_bf16 != 0: dwn is a bf16 [n,d] matrix (halves the spill that pfc_dw_sgd re-reads; fused-SGD mode only);
 * 1: stored with L2 evict_last hints for a pfc_dw_sgd that runs right behind it, 2: plain stores. */
int pfc_backward_dw(const void* E_bf16, int n_pad, const void* xs_bf16, int B, int n, int d, void* dwn, int dwn_bf16,
                    void* stream);
/* inv_grad_scale must be 1: the un-fused gradient keeps the loss scale for GradScaler.unscale_ (model/FR_PartialFC.py:180) */
int pfc_dw_finalize(const float* dwn, const float* w, const float* inv_norm_w, int rows, int d, float inv_grad_scale,
                    float* dw, void* stream);
/* grad_scale: device scalar with the loss scale the gradient carries (g = d loss of pfc_backward_prepare), divided out
 * before the step; NULL = 1.  step_dev (Adam): int32 device scalar, the update is step step_dev[0] + 1 and `step` is
 * ignored (CUDA-graph replay; the caller increments it); NULL: `step` (host, >= 1).
 * index (sampled shards, nets/PartialFC.py:120-121 + :142-143 without the copies): NULL, or the ascending list of active
 * classes -- row r of dwn / inv_norm_w (/ wn_next) then belongs to row index[r] of w and of the optimizer state, which
 * are the FULL [num_local, d] arrays and are updated in place: no gather of the active rows before the step and no
 * scatter back after it.
 * wn_next_copy_bf16 (fp16_operands only, else NULL): a second [rows, d] matrix that receives the same next-step rows as
 * bf16 -- bit for bit what pfc_cast_f16_to_bf16(wn_next) would produce -- so that pfc_backward_dx of the next step finds
 * its bf16 operand without a separate cast pass over the shard. */
int pfc_dw_sgd(const void* dwn, int dwn_bf16, float* w, float* momentum_buf, const float* inv_norm_w, int rows, int d,
               float lr, float momentum, float weight_decay, const float* grad_scale, void* wn_next_bf16,
               float* inv_norm_next, const int64_t* index, int fp16_operands, void* wn_next_copy_bf16, void* stream);
int pfc_dw_adam(const float* dwn, float* w, float* exp_avg, float* exp_avg_sq, const float* inv_norm_w, int rows,
                int d, float lr, float beta1, float beta2, float eps, float weight_decay, int step, int decoupled,
                const float* grad_scale, void* wn_next_bf16, float* inv_norm_next, const int* step_dev,
                const int64_t* index, int fp16_operands, void* wn_next_copy_bf16, void* stream);
/* ---- (5b) the three exchanges of the step over peer memory (NVLink / NVSwitch), fused into the producing kernels.
 * They replace all_gather (nets/PartialFC.py:182-186), the softmax all_reduces (:448, :453, :459) and the dX
 * reduce (:505-522).  peer_* arguments are HOST arrays of W device pointers: entry q is rank q's symmetric buffer as
 * mapped into this process (torch.distributed._symmetric_memory); W <= pfc_peer_max_ranks().
 * pfc_peer_barrier: every store issued by any rank before it is visible to every rank after it.  peer_flags[q] ->
 *   rank q's uint32[W] flag array (zeroed once), epoch_counter -> this rank's private uint32 (zeroed once).
 * pfc_peer_l2norm_gather: pfc_l2norm_rows of the local batch, written as rows [rank*b, rank*b+b) of every rank's
 *   xn_all [W*b, d] bf16, plus the local labels into every rank's labels_all [W*b] int64.
 * pfc_peer_row_stats: pfc_row_stats, written into slot `rank` of every rank's slots [W][B][2] fp32.
 * pfc_peer_loss: stats = sum over the W slots (rank order, so every rank gets identical bits) + pfc_loss.
 * pfc_peer_localize_labels / pfc_peer_loss (peer_flags != NULL) / pfc_peer_dx_finalize take the barrier at their own
 *   start instead of behind a separate pfc_peer_barrier launch: barrier_state -> this rank's uint32[2] {epoch, ticket}
 *   (zeroed once; the same epoch word pfc_peer_barrier uses).  pfc_peer_dx_finalize = barrier + pfc_dx_finalize over
 *   this rank's dx_slots (splits = W, no coefficients).
 * pfc_peer_dx_scatter: coef[i] * sum_z partial[z][i,:] of global row i -> slot `rank` of rank i/b's dx_slots
 *   [W][b][d] fp32; the owner then runs pfc_dx_finalize(dx_slots, splits = W, coef = NULL, ...). */
int pfc_peer_max_ranks(void);
/* How long a rank waits for its peers at a flag barrier before it prints and traps (a CUDA error on this rank instead of
 * a silent hang): milliseconds of SM clock at 2 GHz, default 600 000 (NCCL's watchdog default), 0 = wait forever. */
int pfc_peer_set_timeout_ms(double ms);
int pfc_peer_barrier(void* const* peer_flags, uint32_t* epoch_counter, int rank, int W, void* stream);
int pfc_peer_l2norm_gather(const float* x, const int64_t* labels, int b, int d, int rank, int W,
                           void* const* peer_xn_all, void* const* peer_labels_all, float* inv_norm, int fp16_operands,
                           void* stream);
int pfc_peer_row_stats(const float* part_sum, int n_tiles, int B, const int32_t* labels_local, const float* tgt_e,
                       int rank, int W, void* const* peer_slots, void* stream);
int pfc_peer_loss(void* const* peer_flags, uint32_t* barrier_state, int rank, const float* slots, int W, int B,
                  float* stats, float* row_L, float* loss, void* stream);
/* barrier + pfc_peer_loss + pfc_backward_prepare in one multi-CTA launch (no-autograd step): one warp per row sums the W
 * slots in rank order and forms the row's coefficients; the last CTA (ticket: uint32, zeroed once) reduces the loss. */
int pfc_peer_loss_prepare(void* const* peer_flags, uint32_t* barrier_state, int rank, const float* slots, int W, int B,
                          float* stats, float* row_L, float* loss, unsigned int* ticket, const float* grad_loss, float s,
                          int d, const int32_t* labels_local, const float* tgt_raw, int margin_kind, float m2,
                          const void* xn_bf16, void* xs_bf16, float* coef, void* E_bf16, int n_pad, int fp16_operands,
                          void* stream);
int pfc_peer_localize_labels(void* const* peer_flags, uint32_t* barrier_state, int rank, int W, const int64_t* labels,
                             int B, int64_t class_start, int num_local, int32_t* labels_local, void* stream);
int pfc_peer_dx_finalize(void* const* peer_flags, uint32_t* barrier_state, int rank, int W, const float* dx_slots,
                         const float* x, const float* inv_norm, float scale, int b, int d, float* out, void* stream);
int pfc_peer_dx_scatter(const float* partial, int splits, const float* coef, int B, int b, int d, int rank, int W,
                        void* const* peer_dx_slots, void* stream);
/* ---- (6) pair verification, utils/eval.py.
 * fr_pair_score  (:68-99): scores[i] = 1 - ||e1_i - e2_i||^2/4 (fp32 difference, fp64 accumulation), optional
 *   dist[i] = ||.||^2, and the 100001-bin genuine / imposter histograms (uint64 counts; zeroed by the call).
 * fr_roc         (:7-51, :140-144): EER threshold + FRR at FAR = 10^-k for k in [min_level, max_level].
 * fr_acc_counts  (:54-66): fr_fa = { #(score <= threshold & label == 1), #(score > threshold & label == 0) }.
 * fr_kfold_acc: standard LFW protocol -- `folds` contiguous folds, thresholds t*step on dist, best-train threshold
 *   applied to the held-out fold (not in the reference; BASELINE.json config 5). correct_ws: folds*n_thr uint32. */
typedef struct {
    int32_t eer_threshold;
    int32_t pad;
    double eer;
    double total_genuine, total_imposter;
    double frr_at[16];   /* NaN where the reference would leave None */
    int32_t th_at[16];   /* -1 where the reference would leave None */
} fr_roc_out_t;

int pfc_eval_hist_bins(void);
int fr_pair_score(const float* e1, const float* e2, const uint8_t* labels, int N, int d, double* scores, double* dist,
                  unsigned long long* hist_genuine, unsigned long long* hist_imposter, void* stream);
/* fr_cross_score (utils/eval.py:102-137): all pairs j < i of ONE embedding set; scores / label_list have
 * N(N-1)/2 entries in the reference's order l = i(i-1)/2 + j; genuine when labels[j] == labels[i] (int64 ids). */
int fr_cross_score(const float* e, const long long* labels, int N, int d, double* scores, double* label_list,
                   unsigned long long* hist_genuine, unsigned long long* hist_imposter, void* stream);
int fr_roc(const unsigned long long* hist_genuine, const unsigned long long* hist_imposter, int min_level,
           int max_level, void* roc_out /* fr_roc_out_t, device */, void* stream);
int fr_acc_counts(const double* scores, const uint8_t* labels, int N, double threshold, unsigned long long* fr_fa,
                  void* stream);
int fr_kfold_acc(const double* dist, const uint8_t* labels, int N, int folds, int n_thr, double step,
                 unsigned int* correct_ws, double* acc, int* best_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFC_B200_H */
