// The three dense contractions of the margin-softmax head on tcgen05 (see pfc_umma.cuh for the pipeline)
// and their C-ABI launchers (declared in include/pfc.h).
//
// Numerics of the forward epilogue (reference: nets/PartialFC.py:198-207, nets/ArcFace.py:76-91,
// nets/PartialFC.py:442-461).  All logits are bounded, |z| <= s, so instead of an online row maximum the
// epilogue uses one fixed shift:   e = 2^(log2e*(z - s) + PFC_EXP_TOP).   With PFC_EXP_TOP = 64 and s <= 64
// every term is a normal fp32/bf16 number (2^-121 .. 2^64) and a row sum over <= 2^21 classes cannot
// overflow, so no term is ever flushed and the softmax is exact up to fp32 rounding.  e is therefore final
// the moment it is produced; it is spilled once as bf16 (E') and both gradient GEMMs read it back:
//   dXn_i = c_i * sum_c E'_ic Wn_c,    dWn_c = sum_i E'_ic (c_i Xn_i),    c_i = g*s / (B * L_i)
// where L_i is the global row sum and the target column of E' is patched to -dm_i*mask_i*Lothers_i
// (pfc_backward_prepare) so that the one-hot term needs no separate pass.
#include <stdlib.h>
#include "pfc_umma.cuh"
#include "pfc_umma2.cuh"
#include "pfc_internal.h"

namespace pfc {

// ============================================================================ G1: forward
struct FwdPolicy {
    // 8 epilogue warps x 128 columns, 4-stage operand ring.  Measured alternative (EPI_WARPS_ = 16, 64 columns each,
    // which costs one operand stage for the extra staging buffers and caps the kernel at 96 registers): 123 us
    // against 102 us -- the deeper ring matters more than the extra warps for this K = 512 contraction.
    static constexpr int EPI_WARPS_ = 8;
    static constexpr int STAGES_ = 4;
    static constexpr int PAIR_STAGES_ = 6;
    static constexpr int COLS = BN / (EPI_WARPS_ / 4);
    static constexpr bool A_MN = false;
    static constexpr bool B_MN = false;
    static constexpr bool A_BLOCKED = false;
    static constexpr bool SHARE_B = true;    // a CTA pair = two sample tiles sharing one class (W) stage
    struct Params {
        int num_tiles;
        int B, n, n_pad, B_pad;
        int m_tiles, k_stages;
        const int32_t* labels;   // [B] shard-local class id (after sampling remap) or -1
        float k1, k2;            // e = exp2(clamp(cos) * k1 - k2)
        float k1x2, k12;         // same through h = (clamp+1)/2: e = exp2(h * k1x2 - k12)
        float cos_m, sin_m, theta, sinmm, m3;
        int margin_kind;         // 0 = ArcFace-style (cos(theta+m)), 1 = CosFace-style (t - m3)
        float filter_thr;        // CombinedMarginLoss.interclass_filtering_threshold (0 = off)
        __nv_bfloat16* E;        // [B, n_pad]
        float* part_sum;         // [4 * n_tiles_n, B_pad] sum of non-target e over each 64-class quarter tile
        float* tgt_raw;          // [B] raw (unclamped) target cosine, written by the owning tile only
        float* tgt_e;            // [B] e of the margin-adjusted target logit
        float* tgt_z;            // [B] margin-adjusted target logit (already * s)
        float s;
        uint32_t idesc_xor;      // flips the operand formats of the bf16 instruction descriptor (fp16 operands)
    };
    __device__ static __forceinline__ DescCfg desc(const Params&) { return default_desc_cfg(false, false); }
    __device__ static __forceinline__ TileCoord tile(const Params& p, int t) {
        TileCoord tc;
        const int ct = t / p.m_tiles;
        tc.m0 = (t - ct * p.m_tiles) * BM;   // sample tile fastest: CTAs running together share the W tile in L2
        tc.n0 = ct * BN;
        tc.k0 = 0;
        tc.k1 = p.k_stages;
        tc.aux = ct;
        return tc;
    }

    // 32 accumulator columns -> 16 packed bf16x2 words o[kOff .. kOff+16); returns the sum of the non-target terms
    // (four independent partial sums: no 32-long dependent FADD chain)
    template <bool kHasTarget, bool kTail, bool kFilter>
    __device__ static __forceinline__ float chunk(const Params& p, const uint32_t (&v)[32], uint32_t (&o)[16],
                                                  int col_base, int jt) {
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            float e2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float raw = __uint_as_float(v[j + u]);
                // clamp(raw,-1,1) on the FMA pipe: h = sat(raw/2 + 1/2) = (clamp+1)/2, e = 2^(2 k1 h - (k1 + k2)).
                // The clamp's gradient gate (|raw| <= 1) is applied on the target column only (backward_prepare):
                // for a non-target class it could only fire if a foreign class centre coincided with the sample,
                // where the reference's own fp32 rounding decides the gate at random.
                float h = __saturatef(fmaf(raw, 0.5f, 0.5f));
                bool keep = true;
                if (kFilter) {
                    if (fmaf(h, 2.f, -1.f) > p.filter_thr) { h = 0.5f; keep = false; }
                }
                float e = fast_exp2(fmaf(h, p.k1x2, -p.k12));
                if (kTail) {
                    if (col_base + j + u >= p.n) e = 0.f;
                }
                if (kHasTarget) {
                    if (j + u == jt) e = 0.f;            // the target column is accounted separately
                }
                sum[((j >> 1) & 1) * 2 + u] += e;
                e2[u] = keep ? e : 0.f;
            }
            o[j >> 1] = pack_bf16x2(e2[0], e2[1]);
        }
        return (sum[0] + sum[1]) + (sum[2] + sum[3]);
    }

    // v = 32 accumulator columns already in registers
    __device__ static __forceinline__ float chunk32(const Params& p, const uint32_t (&v)[32], uint32_t (&o)[16],
                                                    int col_base, int row, int tgt_off_in_tile, int tile_col) {
        // tile_col = column of this chunk inside the 256-wide tile; tgt_off_in_tile = target column inside the tile or -1
        const bool has_t = (tgt_off_in_tile >= tile_col) && (tgt_off_in_tile < tile_col + 32);
        const bool tail = col_base + 32 > p.n;
        const bool filt = p.filter_thr > 0.f;
        const int jt = has_t ? (tgt_off_in_tile - tile_col) : -1;
        float sum;
        if (!has_t && !tail && !filt) {
            sum = chunk<false, false, false>(p, v, o, col_base, jt);
        } else if (!filt) {
            sum = chunk<true, true, false>(p, v, o, col_base, jt);
        } else {
            sum = chunk<true, true, true>(p, v, o, col_base, jt);
        }
        if (has_t) {
            float raw = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) raw = (j == jt) ? __uint_as_float(v[j]) : raw;
            const float t = fminf(fmaxf(raw, -1.f), 1.f);
            float fin;
            if (p.margin_kind == 0) {
                const float sin_t = sqrtf(fmaxf(1.f - t * t, 0.f));
                const float ctm = t * p.cos_m - sin_t * p.sin_m;
                fin = (t > p.theta) ? ctm : (t - p.sinmm);
            } else {
                fin = t - p.m3;
            }
            p.tgt_raw[row] = raw;
            p.tgt_e[row] = fast_exp2(fmaf(fin, p.k1, -p.k2));
            p.tgt_z[row] = fin * p.s;
        }
        return sum;
    }

    // The warp owns 32 rows x 128 columns: four 32-column TMEM loads, software-pipelined so that the next chunk is in
    // flight while the current one goes through the FMA / MUFU pipes.
    __device__ static __forceinline__ void epilogue(const Params& p, const TileCoord& tc, uint32_t taddr, int quarter,
                                                    int half, int lane, uint32_t stage, const CUtensorMap* tmc) {
        const int row0 = tc.m0 + quarter * 32;
        const int row = row0 + lane;
        const bool row_ok = row < p.B;
        const int lbl = row_ok ? p.labels[row] : -1;
        const int tgt_off = (lbl >= tc.n0 && lbl < tc.n0 + BN) ? (lbl - tc.n0) : -1;
        float sum = 0.f;
        uint32_t va[32], vb[32];
        tmem_ld_32x32(taddr, va);
#pragma unroll
        for (int cc = 0; cc < COLS / 64; ++cc) {
            const int tile_col = half * COLS + cc * 64;
            const int col64 = tc.n0 + tile_col;
            if (col64 >= p.n) break;                     // warp-uniform: nothing valid from here on
            uint32_t o[16];                              // 32 columns = 64 bytes of the row, staged half by half
            tmem_ld_wait();                              // va = columns [cc*64, cc*64+32)
            tmem_ld_32x32(taddr + cc * 64 + 32, vb);     // next chunk in flight during the math below
            sum += chunk32(p, va, o, col64, row, tgt_off, tile_col);
            warp_tma_store_begin(lane);                  // the previous box of this warp has left the staging buffer
            warp_tma_store_half<0>(stage, lane, o);
            tmem_ld_wait();                              // vb ready
            if (cc + 1 < COLS / 64) tmem_ld_32x32(taddr + (cc + 1) * 64, va);
            if (col64 + 32 < p.n) {
                sum += chunk32(p, vb, o, col64 + 32, row, tgt_off, tile_col + 32);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = 0u;
            }
            warp_tma_store_half<1>(stage, lane, o);
            // E'[col64 / 64][row0 .. row0+32)[0..64): one contiguous 4 KB piece of the class-blocked spill; the store
            // map clips rows >= B
            warp_tma_store_commit(stage, lane, tmc, 0, row0, col64 >> 6);
        }
        tmem_ld_wait();                                  // nothing may be outstanding when the accumulator is released
        if (row_ok) p.part_sum[static_cast<size_t>(tc.aux * (BN / COLS) + half) * p.B_pad + row] = sum;
    }
};

// ============================================================================ G2 / G3: fp32 tile store
// out[aux][row][col] = accumulator, rows < rows_valid, cols < cols_valid.
struct StoreParams {
    int num_tiles;
    int m_tiles, n_tiles, splits;
    int k_stages_total, k_stages_per_split;
    int n_fastest;
    int m_reverse;          // 1: walk the M tiles from the last to the first (most recently written rows of A first)
    int rows_valid, cols_valid;
    int ld;                 // row stride of out (elements)
    size_t split_stride;    // elements between split slabs
    float* out;
    int out_bf16;           // 1: `out` is a bf16 matrix (same ld in elements); used for the dWn spill in fused mode
    DescCfg dc;             // descriptor geometry (runtime so that tools/gpu_probe.py can try alternatives)
    uint32_t idesc_xor;     // flips BOTH operand formats of the bf16 instruction descriptor (0 for the gradient GEMMs)
};

// kKeep (dW GEMM only, PFC_L2_GRAD): the bf16 gradient tiles are stored with an L2 evict_last hint so that the update
// kernel, which runs next, reads them from L2 (and then discards them) instead of HBM.
template <bool kAMN, bool kKeep = false>
struct StorePolicy {
    static constexpr int EPI_WARPS_ = 8;
    static constexpr int STAGES_ = 4;
    static constexpr int PAIR_STAGES_ = 6;
    static constexpr int EPI_COLS = BN / (EPI_WARPS_ / 4);
    static constexpr bool A_MN = kAMN;
    static constexpr bool B_MN = true;
    static constexpr bool A_BLOCKED = true;  // A is the spill E' (dX: K-major, dW: MN-major), stored class-blocked
    // dX (A K-major): a CTA pair = two sample tiles sharing the Wn stage.
    // dW (A MN-major): a CTA pair = the two D halves of one class tile sharing the E'^T stage.
    static constexpr bool SHARE_B = !kAMN;
    using Params = StoreParams;
    __device__ static __forceinline__ DescCfg desc(const Params& p) { return p.dc; }
    __device__ static __forceinline__ TileCoord tile(const Params& p, int t) {
        // n_fastest=0: m fastest (CTAs running together share the B stage in L2);
        // n_fastest=1: the N tiles of one M tile run side by side (they share the A stage in L2);
        // n_fastest=2: the same for the CTA-pair kernel (see below)
        TileCoord tc;
        const int z = t / (p.m_tiles * p.n_tiles);
        const int r = t - z * (p.m_tiles * p.n_tiles);
        int mt, nt;
        if (p.n_fastest == 2) {          // CTA-pair kernel: tiles 2P, 2P+1 = the two M tiles of pair P; the N tiles of one
            const int pr = r >> 1;       // M pair run on neighbouring clusters (they share the A stage in L2)
            const int mp = pr / p.n_tiles;
            nt = pr - mp * p.n_tiles;
            mt = 2 * mp + (r & 1);
        } else if (p.n_fastest) { mt = r / p.n_tiles; nt = r - mt * p.n_tiles; }
        else                    { nt = r / p.m_tiles; mt = r - nt * p.m_tiles; }
        if (p.m_reverse) mt = p.m_tiles - 1 - mt;
        tc.m0 = mt * BM;
        tc.n0 = nt * BN;
        tc.k0 = z * p.k_stages_per_split;
        tc.k1 = min(tc.k0 + p.k_stages_per_split, p.k_stages_total);
        tc.aux = z;
        return tc;
    }
    // Output goes through the store tensor map tmc = {cols_valid, rows_valid, splits} (fp32, boxes of 32 columns) or
    // {cols_valid, rows_valid, 1} (bf16, boxes of 64 columns); rows / columns beyond the tensor are clipped by TMA.
    __device__ static __forceinline__ void epilogue(const Params& p, const TileCoord& tc, uint32_t taddr, int quarter,
                                                    int half, int lane, uint32_t stage, const CUtensorMap* tmc) {
        const int row0 = tc.m0 + quarter * 32;
        if (row0 >= p.rows_valid) return;                // warp-uniform: an all-padding row block
        if (p.out_bf16) {
            // bf16 spill: two 32-column chunks make one 128-byte line per row
#pragma unroll 1
            for (int cc = 0; cc < EPI_COLS / 64; ++cc) {
                const int col64 = tc.n0 + half * EPI_COLS + cc * 64;
                if (col64 >= p.cols_valid) break;        // warp-uniform
                uint32_t o[32];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + cc * 64 + h2 * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        o[h2 * 16 + j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                }
                warp_tma_store_rows<kKeep>(stage, lane, o, tmc, col64, row0, 0);
            }
            return;
        }
#pragma unroll 1
        for (int c = 0; c < EPI_COLS / 32; ++c) {
            const int col_base = tc.n0 + half * EPI_COLS + c * 32;
            if (col_base >= p.cols_valid) break;         // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait();
            warp_tma_store_rows(stage, lane, v, tmc, col_base, row0, tc.aux);
        }
    }
};

// ============================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// bf16 row-major [outer, inner] tensor, box = box_inner x box_outer, 128-byte swizzle, OOB reads give zero.
static int make_tmap(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                     uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return PFC_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * 2) % 16) return PFC_ERR_ALIGNMENT;
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstr[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PFC_OK : PFC_ERR_TENSORMAP;
}

// Store target: row-major [slabs][outer][inner] tensor of bf16 or fp32, boxes of 128 bytes x 32 rows x 1 slab in the
// SWIZZLE_128B layout warp_tma_store_rows() stages.
static int make_store_tmap(CUtensorMap* map, void* ptr, bool is_bf16, uint64_t inner, uint64_t outer, uint64_t slabs,
                           uint64_t row_stride_elems, uint64_t slab_stride_elems) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return PFC_ERR_DRIVER;
    const uint64_t es = is_bf16 ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * es) % 16 || (slab_stride_elems * es) % 16)
        return PFC_ERR_ALIGNMENT;
    cuuint64_t gdim[3] = {inner, outer, slabs};
    cuuint64_t gstr[2] = {row_stride_elems * es, (slabs > 1 ? slab_stride_elems : row_stride_elems * outer) * es};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / es), 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, gdim, gstr,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PFC_OK : PFC_ERR_TENSORMAP;
}

// Load map of the class-blocked spill E'[n_pad/64][B][64] (bf16): dims {64, B, n_pad/64}, boxes {64, box_rows, 1} in
// SWIZZLE_128B -- the same shared-memory image as a {64, box_rows} box of the row-major matrix, but every box is ONE
// contiguous piece of global memory (box_rows x 128 B) instead of box_rows pieces 2*n_pad bytes apart.
static int make_blocked_tmap(CUtensorMap* map, const void* ptr, uint64_t B, uint64_t n_pad, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return PFC_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || n_pad % 64) return PFC_ERR_ALIGNMENT;
    cuuint64_t gdim[3] = {64, B, n_pad / 64};
    cuuint64_t gstr[2] = {128, B * 128};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PFC_OK : PFC_ERR_TENSORMAP;
}

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
    }
    return sms;
}

// Debug hook (tools/gpu_probe.py): override the MN-major descriptor geometry. 0 = keep default.
static uint32_t g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0, g_dbg_mn_kstep = 0;
static DescCfg store_desc_cfg(bool a_mn) {
    DescCfg dc = default_desc_cfg(a_mn, true);
    if (g_dbg_mn_lbo) { dc.b_lbo = g_dbg_mn_lbo; if (a_mn) dc.a_lbo = g_dbg_mn_lbo; }
    if (g_dbg_mn_sbo) { dc.b_sbo = g_dbg_mn_sbo; if (a_mn) dc.a_sbo = g_dbg_mn_sbo; }
    if (g_dbg_mn_kstep) { dc.b_kstep = g_dbg_mn_kstep; if (a_mn) dc.a_kstep = g_dbg_mn_kstep; }
    return dc;
}

// GEMM launch mode (pfc_debug_cluster): 0 = auto, 1 = single-CTA kernel, 2 = CTA pair sharing one operand stage by
// TMA multicast, 3 = cta_group::2 pair kernel (pfc_umma2.cuh).  Measured on B200 at cfg-2 (event-timed in the bench
// step, us; single / pair): forward 92.9 / 80.5, dX 76.3 / 69.7, dW 85.1 / 88.7 -- auto takes the pair kernel for the
// forward and dX and the single-CTA kernel for dW.  Two fixes made the pair kernel win: its accumulator hand-back is
// a RELAXED cluster arrive (the release form cost a MEMBAR.ALL per warp and tile, 28 % of the stall samples), and the
// MMA warp walks its loops warp-uniformly with one elected lane issuing (one thread then feeds two SMs' tensor cores
// with 4 back-to-back UTCHMMA per stage instead of ~20 dependent instructions per UTCHMMA).
static int g_gemm_mode = 0;
enum { MODE_SINGLE = 1, MODE_MCAST = 2, MODE_PAIR = 3 };

static int pick_mode(int m_tiles, int mcast_dim_tiles, int auto_mode) {
    int mode = g_gemm_mode == 0 ? auto_mode : g_gemm_mode;
    if (mode == MODE_PAIR && m_tiles < 2) mode = MODE_SINGLE;
    if (mode == MODE_MCAST && (mcast_dim_tiles % 2)) mode = MODE_SINGLE;
    return mode;
}

template <class Kern, class Params>
static int launch_cluster(int pdl_id, Kern kern, int cluster, int threads, int smem_bytes, const CUtensorMap& ta,
                          const CUtensorMap& tb, const CUtensorMap& tc, const Params& prm, cudaStream_t stream) {
    const int sms = num_sms();
    if (sms <= 0) return PFC_ERR_CUDA;
    if (prm.num_tiles <= 0) return PFC_OK;
    if (prm.num_tiles % cluster) return PFC_ERR_SHAPE;
    int grid = prm.num_tiles < sms ? prm.num_tiles : sms;
    grid = grid / cluster * cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    (void)pdl_id;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, prm);
    return e == cudaSuccess ? PFC_OK : PFC_ERR_LAUNCH;
}

template <class P>
static int launch_gemm(int pdl_id, int mode, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const typename P::Params& prm, cudaStream_t stream) {
    using C = GemmCfg<P>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(umma_gemm_kernel<P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(umma_gemm_kernel<P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(umma_gemm_pair_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::PAIR_SMEM) != cudaSuccess)
            return PFC_ERR_CUDA;
        attr_set = true;
    }
    if (mode == MODE_PAIR)
        return launch_cluster(pdl_id, umma_gemm_pair_kernel<P>, 2, C::THREADS, C::PAIR_SMEM, ta, tb, tc, prm, stream);
    if (mode == MODE_MCAST)
        return launch_cluster(pdl_id, umma_gemm_kernel<P, 2>, 2, C::THREADS, C::SMEM, ta, tb, tc, prm, stream);
    return launch_cluster(pdl_id, umma_gemm_kernel<P, 1>, 1, C::THREADS, C::SMEM, ta, tb, tc, prm, stream);
}

static int even_up(int v) { return (v + 1) / 2 * 2; }

static void fill_fwd_params(FwdPolicy::Params& p, int B, int n, int n_pad, int d, int m_tiles, float s, int margin_kind,
                            float m2, float m3, float filter_thr, const int32_t* labels_local, void* E, float* part_sum,
                            float* tgt_raw, float* tgt_e, float* tgt_z) {
    const float log2e = 1.4426950408889634f;
    p.B = B; p.n = n; p.n_pad = n_pad; p.B_pad = (B + BM - 1) / BM * BM;
    p.m_tiles = m_tiles;                    // an odd count gets one all-padding tile in the pair kernel
    p.k_stages = (d + BK - 1) / BK;
    p.num_tiles = p.m_tiles * ((n + BN - 1) / BN);
    p.labels = labels_local;
    p.k1 = s * log2e;
    p.k2 = s * log2e - (float)PFC_EXP_TOP;
    p.k1x2 = 2.f * p.k1;
    p.k12 = p.k1 + p.k2;
    {   // same double-precision constants the reference computes with math.cos/sin (nets/ArcFace.py:69-72)
        const double pi = 3.14159265358979323846;
        p.cos_m = (float)cos((double)m2); p.sin_m = (float)sin((double)m2);
        p.theta = (float)cos(pi - (double)m2);
        p.sinmm = (float)(sin(pi - (double)m2) * (double)m2);
    }
    p.m3 = m3; p.margin_kind = margin_kind; p.filter_thr = filter_thr;
    p.E = reinterpret_cast<__nv_bfloat16*>(E);
    p.part_sum = part_sum; p.tgt_raw = tgt_raw; p.tgt_e = tgt_e; p.tgt_z = tgt_z; p.s = s;
}

// Keep the bf16 gradient of the dW GEMM in L2 for the update kernel that runs right behind it (evict_last stores here,
// evict_first streams + discard.global.L2 in dw_sgd_rows_kernel, pfc_rows.cu).  Bit-identical results; measured on B200 at
// cfg-2: 0.3872 -> 0.3812 ms per step on one GPU, 0.2486 -> 0.2452 on two (profiles/r02a_experimental_n*.log), hence on by
// default.  PFC_L2_GRAD=0 / pfc_debug_l2_grad(0) switch it off (the A/B test does); dwn_bf16 = 2 stores without the hints
// (tools/exp_overlap.py: chunked pipelines in which the update does not run right behind the GEMM).
static int g_l2_grad = -1;

}  // namespace pfc

using namespace pfc;

extern "C" {

int pfc_exp_top(void) { return PFC_EXP_TOP; }

// not part of the public header: descriptor-geometry override for the MN-major operands (0 = default)
void pfc_debug_mn_desc(unsigned lbo, unsigned sbo, unsigned kstep) {
    g_dbg_mn_lbo = lbo; g_dbg_mn_sbo = sbo; g_dbg_mn_kstep = kstep;
}
// not part of the public header: GEMM launch mode, see g_gemm_mode
void pfc_debug_cluster(int mode) { g_gemm_mode = mode; }
// not part of the public header: L2-resident bf16 gradient between pfc_backward_dw and pfc_dw_sgd (see g_l2_grad)
void pfc_debug_l2_grad(int on) { g_l2_grad = on ? 1 : 0; }
int pfc_l2_grad_enabled(void) {
    if (g_l2_grad < 0) {
        const char* e = getenv("PFC_L2_GRAD");
        g_l2_grad = e ? (atoi(e) != 0) : 1;
    }
    return g_l2_grad;
}

int pfc_padded_classes(int n) { return (n + 63) / 64 * 64; }
int pfc_num_class_tiles(int n) { return (BN / FwdPolicy::COLS) * ((n + BN - 1) / BN); }   // part_sum slabs
int pfc_part_sum_cols(void) { return FwdPolicy::COLS; }   // classes per part_sum slab
int pfc_padded_batch(int B) { return (B + BM - 1) / BM * BM; }

int pfc_forward(const void* xn, const void* wn, const int32_t* labels_local, int B, int n, int d, float s,
                int margin_kind, float m2, float m3, float filter_thr, void* E, int n_pad, float* part_sum,
                float* tgt_raw, float* tgt_e, float* tgt_z, int fp16_operands, void* stream) {
    if (B <= 0 || n <= 0 || d <= 0 || d % 8 || n_pad % 8 || n_pad < n) return PFC_ERR_SHAPE;
    const float log2e = 1.4426950408889634f;
    // every representable term must stay a normal bf16/fp32 number: 2*s*log2e <= TOP + 126
    if (!(s > 0.f) || 2.f * s * log2e > PFC_EXP_TOP + 126.f) return PFC_ERR_SCALE_RANGE;
    const int m_tiles = (B + BM - 1) / BM;
    const int mode = pick_mode(m_tiles, m_tiles, MODE_PAIR);   // pairs of sample tiles share the class (W) stage
    if (n_pad % 64) return PFC_ERR_SHAPE;              // the spill is stored in 64-column (128-byte) boxes
    CUtensorMap ta, tb, tc;
    int rc = make_tmap(&ta, xn, d, B, d, BK, BM);
    if (rc) return rc;
    // in both pair modes each CTA fetches half of the 256 class rows of a stage
    rc = make_tmap(&tb, wn, d, n, d, BK, mode == MODE_SINGLE ? BN : BN / 2);
    if (rc) return rc;
    rc = make_store_tmap(&tc, E, true, 64, B, n_pad / 64, 64, static_cast<uint64_t>(B) * 64);
    if (rc) return rc;
    FwdPolicy::Params p;
    fill_fwd_params(p, B, n, n_pad, d, mode == MODE_PAIR ? even_up(m_tiles) : m_tiles, s, margin_kind, m2, m3, filter_thr,
                    labels_local, E, part_sum, tgt_raw, tgt_e, tgt_z);
    p.idesc_xor = fp16_operands ? (UMMA_IDESC_A_BF16 | UMMA_IDESC_B_BF16) : 0u;     // Xn and Wn are fp16
    return launch_gemm<FwdPolicy>(PDL_FORWARD, mode, ta, tb, tc, p, reinterpret_cast<cudaStream_t>(stream));
}

// Number of class splits the dX contraction uses for a given shape (callers size `partial` with it).
int pfc_dx_splits(int B, int n, int d) {
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const int base = even_up((B + BM - 1) / BM) * ((d + BN - 1) / BN);
    const int k_total = (n + BK - 1) / BK;
    int splits = sms / base;
    if (splits < 1) splits = 1;
    if (splits > k_total) splits = k_total;
    const int per = (k_total + splits - 1) / splits;
    return (k_total + per - 1) / per;   // no empty split
}

// Upper bound of pfc_dx_splits over every n (callers size the partial buffer with it).
int pfc_dx_max_splits(int B, int d) {
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const int base = even_up((B + BM - 1) / BM) * ((d + BN - 1) / BN);
    return sms / base > 1 ? sms / base : 1;
}

// partial[z][B][d] (fp32) = E'[:, classes of split z] . Wn[classes of split z, :]
int pfc_backward_dx(const void* E, int n_pad, const void* wn, int B, int n, int d, float* partial, int splits,
                    void* stream) {
    if (B <= 0 || n <= 0 || d <= 0 || d % 8 || n_pad % 8 || splits <= 0) return PFC_ERR_SHAPE;
    CUtensorMap ta, tb;
    if (n_pad % 64) return PFC_ERR_SHAPE;
    int rc = make_blocked_tmap(&ta, E, B, n_pad, BM);      // A: E' K-major (K = classes), one 64-class block per stage
    if (rc) return rc;
    rc = make_tmap(&tb, wn, d, n, d, 64, BK);              // B: Wn [n(K), d(N)] MN-major boxes 64(N) x 64(K)
    if (rc) return rc;
    StoreParams p;
    const int m_tiles = (B + BM - 1) / BM;
    const int mode = pick_mode(m_tiles, m_tiles, MODE_PAIR);     // pairs of sample tiles share the Wn stage
    p.m_tiles = mode == MODE_PAIR ? even_up(m_tiles) : m_tiles;
    p.n_tiles = (d + BN - 1) / BN;
    p.k_stages_total = (n + BK - 1) / BK;
    p.k_stages_per_split = (p.k_stages_total + splits - 1) / splits;
    p.splits = (p.k_stages_total + p.k_stages_per_split - 1) / p.k_stages_per_split;
    if (p.splits != splits) return PFC_ERR_SHAPE;
    p.num_tiles = p.m_tiles * p.n_tiles * p.splits;
    p.n_fastest = 0;
    p.m_reverse = 0;
    p.rows_valid = B; p.cols_valid = d; p.ld = d;
    p.split_stride = static_cast<size_t>(B) * d;
    p.out = partial;
    p.out_bf16 = 0;
    p.dc = store_desc_cfg(false);
    // both operands bf16 in every mode: E' needs the bf16 exponent range and tcgen05 kind::f16 rejects mixed operand
    // formats (bf16 x fp16 raises an illegal-instruction error on sm_100a), so an fp16 shard is cast first (pfc_cast_...)
    p.idesc_xor = 0;
    CUtensorMap tc;
    rc = make_store_tmap(&tc, partial, false, d, B, p.splits, d, static_cast<uint64_t>(B) * d);
    if (rc) return rc;
    return launch_gemm<StorePolicy<false>>(PDL_DX, mode, ta, tb, tc, p, reinterpret_cast<cudaStream_t>(stream));
}

// dwn[n][d] = E'^T . Xs,   Xs = c_i * Xn_i (bf16, [B, d]);  dwn fp32 or (dwn_bf16) bf16
int pfc_backward_dw(const void* E, int n_pad, const void* xs, int B, int n, int d, void* dwn, int dwn_bf16,
                    void* stream) {
    if (B <= 0 || n <= 0 || d <= 0 || d % 8 || n_pad % 8) return PFC_ERR_SHAPE;
    CUtensorMap ta, tb;
    if (n_pad % 64) return PFC_ERR_SHAPE;
    int rc = make_blocked_tmap(&ta, E, B, n_pad, BK);      // A: E'^T MN-major boxes 64(M = classes) x 64(K = samples)
    if (rc) return rc;
    rc = make_tmap(&tb, xs, d, B, d, 64, BK);              // B: Xs [B(K), d(N)] MN-major
    if (rc) return rc;
    StoreParams p;
    const int m_tiles = (n + BM - 1) / BM;
    p.n_tiles = (d + BN - 1) / BN;
    // cta_group::2: pairs of CLASS tiles share the Xs stage (m fastest); multicast mode: the two D halves of one
    // class tile share the E'^T stage (n fastest)
    const int mode = pick_mode(m_tiles, p.n_tiles, MODE_PAIR);
    p.m_tiles = mode == MODE_PAIR ? even_up(m_tiles) : m_tiles;
    p.splits = 1;
    p.k_stages_total = (B + BK - 1) / BK;
    p.k_stages_per_split = p.k_stages_total;
    p.num_tiles = p.m_tiles * p.n_tiles;
    p.n_fastest = mode == MODE_PAIR ? 2 : 1;
    // the forward wrote E' class tile by class tile: its last tiles are still in L2, so start with them; the
    // gradient rows written LAST are then the low ones, which is where the row-ordered update kernel starts
    p.m_reverse = 1;
    p.rows_valid = n; p.cols_valid = d; p.ld = d;
    p.split_stride = 0;
    p.out = reinterpret_cast<float*>(dwn);
    p.out_bf16 = dwn_bf16 ? 1 : 0;
    p.dc = store_desc_cfg(true);
    p.idesc_xor = 0;                                            // E'^T and Xs are bf16 in both modes
    CUtensorMap tc;
    rc = make_store_tmap(&tc, dwn, dwn_bf16 != 0, d, n, 1, d, 0);
    if (rc) return rc;
    if (dwn_bf16 == 1 && pfc_l2_grad_enabled())   // 2: bf16 without the L2 hints
        return launch_gemm<StorePolicy<true, true>>(PDL_DW, mode, ta, tb, tc, p, reinterpret_cast<cudaStream_t>(stream));
    return launch_gemm<StorePolicy<true>>(PDL_DW, mode, ta, tb, tc, p, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
