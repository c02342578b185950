// FX: the forward GEMM (G1, S = Xn . Wn^T with the margin / exp / row-sum / bf16-spill epilogue) and the dX GEMM
// (G2, dXn += E' . Wn) of one head step in ONE persistent cta_group::2 kernel, interleaved class tile by class tile.
//
// Why.  Run back to back, G1 writes the 191 MB spill E' to HBM and G2 reads it (and the 96 MB bf16 shard) back from
// HBM 80 us later, long after the 126 MB L2 has been overwritten: 287 MB of reads that exist only because of the
// kernel boundary.  Here a cluster consumes an E' tile (and the Wn tile that produced it) one pipeline step after it
// was written, i.e. from L2, and the two contractions share the tensor pipe: while the epilogue warps turn the S
// accumulator into E', the MMA warp keeps the pipe busy with dX work instead of waiting for the accumulator
// (G1 alone is epilogue-paced: tensor pipe 64 % active).  The dX GEMM needs neither the softmax denominator nor the
// target patch: dXn_i = c_i * (sum_c E'_ic Wn_c + patch_i Wn_{y_i}) with the forward leaving 0 in the target column;
// c_i and the rank-1 term are applied when the partials are summed (pfc_dx_finalize_patched).
//
// Work split (B rows, n classes, d columns; R = ceil(B/256) row blocks, H = ceil(d/256) column blocks of dX,
// CT = ceil(n/256) class tiles, G = class groups with R*H*G <= #SMs/2):
//   cluster (g, r, h) owns the dX partial tile  P[g][256 r .. +256][256 h .. +256]  for the whole kernel (TMEM columns
//   [0,256) of both CTAs) and the forward tiles (r, ct) of every H-th class tile ct of its group (TMEM columns
//   [256,512)).  The H clusters that share (g, r) therefore compute every E' tile of that row block and group exactly
//   once, exchange them through global memory (L2) under a per-tile counter, and each accumulates
//   E'[r, ct] . Wn[ct, 256 h .. +256] over ALL class tiles of the group.
// Sequence per cluster, p = 0 .. P (P = ceil(L/H), L = class tiles of the group):
//   F(p H + h)   -- 8 K stages (d = 512), accumulator handed to the epilogue warps
//   X((p-1) H + j), j = 0 .. H-1   -- 4 K stages (64 classes) each, lagging one round so that the siblings' E' tiles are
//                                     there when the producer asks for them
// The three roles (TMA producer, MMA issuer, epilogue warps) walk the same sequence; the operand ring, its barriers
// and the two epilogues are those of pfc_umma2.cuh / pfc_gemm.cu (FwdPolicy, StorePolicy<false>).
//
// Optional input dependency (lazy update): wn_ready[ct] counts the rows of class tile ct that the fused update kernel
// of the PREVIOUS step's gradient (dw_sgd_ordered_kernel, pfc_rows.cu -- co-resident on the same SMs, HBM-bound while
// this kernel is tensor-bound) has rewritten; the producer asks for a Wn tile only once all its rows are there.
//
// Deadlock freedom: a wait only ever targets work that is EARLIER in every cluster's sequence (F(k) needs nothing
// from other clusters; X of round p-1 needs F of round p-1), the grid is at most one CTA per SM, and the update kernel
// waits for nothing -- so every CTA becomes resident and every wait is eventually satisfied.  All waits are bounded
// (PFC_WATCHDOG): a protocol bug traps instead of hanging the GPU.
#pragma once
#include "pfc_umma2.cuh"

namespace pfc {

struct FxParams {
    FwdPolicy::Params f;      // forward epilogue parameters (labels, margin, spill, part_sum ...)
    StoreParams x;            // dX partial store parameters (rows_valid = B, cols_valid = d, out_bf16 = 0)
    int R, H, G, CT;          // row blocks, dX column blocks, class groups, class tiles
    int kf_stages;            // K stages of a forward tile = ceil(d / 64)
    int n_pad;
    const int* wn_ready;      // [CT] rows of class tile ct rewritten by the update kernel, or nullptr
    int* e_ready;             // [R * CT] epilogue warps (of both CTAs) that have landed their part of E' tile (r, ct)
    int n;                    // active classes
};

constexpr int FX_STAGES = 6;
constexpr int FX_SMEM = FX_STAGES * PAIR_STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES + 1024;
static_assert(FX_SMEM <= 232448, "FX shared memory budget");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// one thread: wait until *p >= target (written by other CTAs / another kernel with release semantics), then order the
// TMA (async-proxy) reads that follow behind the acquire
__device__ __forceinline__ void wait_counter(const int* p, int target) {
    if (ld_acquire_gpu(p) < target) {
#if PFC_WATCHDOG
        const long long t0 = clock64();
#endif
        while (ld_acquire_gpu(p) < target) {
            __nanosleep(40);
#if PFC_WATCHDOG
            if (clock64() - t0 > 4000000000LL) {
                printf("pfc: FX counter wait timed out (block %d, counter %p = %d, target %d)\n", (int)blockIdx.x,
                       (const void*)p, ld_acquire_gpu(p), target);
                __trap();
            }
#endif
        }
    }
    fence_proxy_async_all();
}

// Register cap 144 (not the 168 the forward epilogue would take): the 10 warps of this kernel sit three to a scheduler on
// two of the SM's four sub-partitions, whose register files (16 384 each) must still take one 72-register warp of the
// co-resident update kernel -- 3 x 144 x 32 + 72 x 32 = 16 128.  With 168 the update kernel's CTAs only found room on
// the 4 SMs this kernel leaves idle and the lazy step took 0.69 ms (profiles/r02b_bench_modes.txt).
__global__ void __maxnreg__(144)
fx_kernel(const __grid_constant__ CUtensorMap tma_xn,      // Xn  [B, d]   box {64, 128}          A of F (K-major)
          const __grid_constant__ CUtensorMap tma_wn_k,    // Wn  [n, d]   box {64, 128}          B of F (K-major, this CTA's half)
          const __grid_constant__ CUtensorMap tma_e_st,    // E'  blocked  box {64, 32, 1}        F epilogue stores
          const __grid_constant__ CUtensorMap tma_e_ld,    // E'  blocked  box {64, 128, 1}       A of X (K-major)
          const __grid_constant__ CUtensorMap tma_wn_mn,   // Wn  [n, d]   box {64 (d), 64 (n)}   B of X (MN-major)
          const __grid_constant__ CUtensorMap tma_dx_st,   // P   [G][B][d] fp32 box {32, 32, 1}  X epilogue stores
          const FxParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + FX_STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + FX_STAGES * PAIR_STAGE_BYTES;

    __shared__ __align__(8) uint64_t full_bar[FX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[FX_STAGES];
    __shared__ __align__(8) uint64_t f_full_bar, f_empty_bar, x_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = static_cast<int>(cluster_ctarank());
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_xn);
        tma_prefetch_desc(&tma_wn_k);
        tma_prefetch_desc(&tma_e_st);
        tma_prefetch_desc(&tma_e_ld);
        tma_prefetch_desc(&tma_wn_mn);
        tma_prefetch_desc(&tma_dx_st);
#pragma unroll
        for (int s = 0; s < FX_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&f_full_bar, 1);
        mbar_init(&f_empty_bar, 2 * EPI_WARPS);
        mbar_init(&x_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    // ---- this cluster's share
    const int cid = blockIdx.x >> 1;
    const int h = cid % prm.H;
    const int r = (cid / prm.H) % prm.R;
    const int g = cid / (prm.H * prm.R);
    const int ct0 = static_cast<int>((static_cast<long long>(g) * prm.CT) / prm.G);
    const int L = static_cast<int>((static_cast<long long>(g + 1) * prm.CT) / prm.G) - ct0;
    const int P = (L + prm.H - 1) / prm.H;
    const int m0 = r * (2 * BM) + crank * BM;           // this CTA's 128 rows of the 256-row block
    const int nx0 = h * BN;                             // first dX column of this cluster

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        uint32_t stage = 0, phase = 0;
        const uint32_t lbar0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        auto fill = [&](bool is_f, int kel, int n0) {
            // one K stage: A 128 rows x 64, B this CTA's half (128 of the 256 B rows / columns)
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one_sync()) {
                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * PAIR_STAGE_BYTES);
                const uint32_t lbar = lbar0 + stage * 8;
                uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
                uint8_t* b_dst = sB + stage * PAIR_B_STAGE_BYTES;
                if (is_f) {
                    tma_load_2d_pair(a_dst, &tma_xn, lbar, kel, m0);
                    tma_load_2d_pair(b_dst, &tma_wn_k, lbar, kel, n0 + crank * (BN / 2));
                } else {
                    tma_load_3d_pair(a_dst, &tma_e_ld, lbar, 0, m0, kel >> 6);          // kel = first class of the stage
#pragma unroll
                    for (int jj = 0; jj < BN / 128; ++jj)
                        tma_load_2d_pair(b_dst + jj * MN_BOX_BYTES, &tma_wn_mn, lbar, nx0 + (crank * (BN / 128) + jj) * 64,
                                         kel);
                }
            }
            __syncwarp();
            if (++stage == FX_STAGES) { stage = 0; phase ^= 1; }
        };
        for (int p = 0; p <= P; ++p) {
            const int kF = p * prm.H + h;
            if (p < P && kF < L) {
                const int ct = ct0 + kF;
                if (prm.wn_ready != nullptr) {
                    if (elect_one_sync()) wait_counter(prm.wn_ready + ct, min(BN, prm.n - ct * BN));
                    __syncwarp();
                }
                for (int kc = 0; kc < prm.kf_stages; ++kc) fill(true, kc * BK, ct * BN);
            }
            if (p >= 1) {
                for (int j = 0; j < prm.H; ++j) {
                    const int kX = (p - 1) * prm.H + j;
                    if (kX >= L) break;
                    const int ct = ct0 + kX;
                    if (elect_one_sync()) wait_counter(prm.e_ready + r * prm.CT + ct, 2 * EPI_WARPS);
                    __syncwarp();
                    const int c0 = ct * BN;
                    const int ks = min(BN / BK, (prm.n_pad - c0 + BK - 1) / BK);
                    for (int kc = 0; kc < ks; ++kc) fill(false, c0 + kc * BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            constexpr uint32_t idesc_f = umma_idesc_bf16(2 * BM, BN, 0, 0);
            constexpr uint32_t idesc_x = umma_idesc_bf16(2 * BM, BN, 0, 1);
            constexpr DescCfg dcf = default_desc_cfg(false, false);
            constexpr DescCfg dcx = default_desc_cfg(false, true);
            const uint64_t a_tmpl = umma_smem_desc_sw128(smem_u32(sA), dcf.a_lbo, dcf.a_sbo);      // K-major in both
            const uint64_t bf_tmpl = umma_smem_desc_sw128(smem_u32(sB), dcf.b_lbo, dcf.b_sbo);
            const uint64_t bx_tmpl = umma_smem_desc_sw128(smem_u32(sB), dcx.b_lbo, dcx.b_sbo);
            constexpr uint32_t a_kstep = dcf.a_kstep >> 4, bf_kstep = dcf.b_kstep >> 4, bx_kstep = dcx.b_kstep >> 4;
            const uint32_t d_x = tmem_base;             // dX partial accumulator: columns [0, 256)
            const uint32_t d_f = tmem_base + BN;        // forward accumulator:    columns [256, 512)
            uint32_t stage = 0, phase = 0;
            int fi = 0;                                 // forward tiles issued so far
            uint32_t x_acc = 0;                         // 0 until the first dX MMA has initialised the accumulator
            for (int p = 0; p <= P; ++p) {
                const int kF = p * prm.H + h;
                if (p < P && kF < L) {
                    mbar_wait(&f_empty_bar, (fi & 1) ^ 1);
                    tc_fence_after();
                    for (int kc = 0; kc < prm.kf_stages; ++kc) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t adesc = a_tmpl + stage * (A_STAGE_BYTES >> 4);
                            const uint64_t bdesc = bf_tmpl + stage * (PAIR_B_STAGE_BYTES >> 4);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_bf16_ss_pair(d_f, adesc + k * a_kstep, bdesc + k * bf_kstep, idesc_f,
                                                  (kc > 0 || k > 0) ? 1u : 0u);
                            umma_commit_pair(&empty_bar[stage], 0b11);
                        }
                        __syncwarp();
                        if (++stage == FX_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one_sync()) umma_commit_pair(&f_full_bar, 0b11);
                    __syncwarp();
                    ++fi;
                }
                if (p >= 1) {
                    for (int j = 0; j < prm.H; ++j) {
                        const int kX = (p - 1) * prm.H + j;
                        if (kX >= L) break;
                        const int c0 = (ct0 + kX) * BN;
                        const int ks = min(BN / BK, (prm.n_pad - c0 + BK - 1) / BK);
                        for (int kc = 0; kc < ks; ++kc) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            if (elect_one_sync()) {
                                const uint64_t adesc = a_tmpl + stage * (A_STAGE_BYTES >> 4);
                                const uint64_t bdesc = bx_tmpl + stage * (PAIR_B_STAGE_BYTES >> 4);
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k)
                                    umma_bf16_ss_pair(d_x, adesc + k * a_kstep, bdesc + k * bx_kstep, idesc_x,
                                                      (k > 0) ? 1u : x_acc);
                                umma_commit_pair(&empty_bar[stage], 0b11);
                            }
                            __syncwarp();
                            x_acc = 1u;
                            if (++stage == FX_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            if (elect_one_sync()) umma_commit_pair(&x_full_bar, 0b11);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps, both CTAs)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const uint32_t stage_buf = smem_u32(sEpi + ew * EPI_STAGE_BYTES);
        const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t f_empty_leader = mapa_u32(smem_u32(&f_empty_bar), 0);
        int fi = 0;
        for (int p = 0; p < P; ++p) {
            const int kF = p * prm.H + h;
            if (kF >= L) break;
            const int ct = ct0 + kF;
            TileCoord tc;
            tc.m0 = m0; tc.n0 = ct * BN; tc.k0 = 0; tc.k1 = prm.kf_stages; tc.aux = ct;
            mbar_wait(&f_full_bar, fi & 1);
            tc_fence_after();
            FwdPolicy::epilogue(prm.f, tc, tmem_base + BN + half * EPI_COLS + lane_bits, quarter, half, lane, stage_buf,
                                &tma_e_st);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster_relaxed(f_empty_leader);     // the MMA warp may overwrite the accumulator
                // this warp's boxes of E' tile (r, ct) have landed in global memory (L2): publish them to the clusters
                // that contract over this tile.  The wait falls into the time the warp would spend waiting for the next
                // accumulator anyway.
                tma_store_wait_all();
                fence_proxy_async_all();
                __threadfence();
                red_release_gpu_add(prm.e_ready + r * prm.CT + ct, 1);
            }
            __syncwarp();
            ++fi;
        }
        // the dX partial tile of this cluster
        {
            TileCoord tc;
            tc.m0 = m0; tc.n0 = nx0; tc.k0 = 0; tc.k1 = 0; tc.aux = g;
            mbar_wait(&x_full_bar, 0);
            tc_fence_after();
            StorePolicy<false>::epilogue(prm.x, tc, tmem_base + half * EPI_COLS + lane_bits, quarter, half, lane, stage_buf,
                                         &tma_dx_st);
            tc_fence_before();
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

}  // namespace pfc
