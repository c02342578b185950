// CTA-pair (cta_group::2) variant of the GEMM skeleton in pfc_umma.cuh.
//
// Two CTAs on the two SMs of a TPC form a cluster and compute a 256 x 256 tile with ONE tcgen05.mma.cta_group::2
// per K step: each CTA keeps its own 128 A rows (its own tile rows; its own TMEM holds its 128 x 256 accumulator)
// but only HALF of the B stage (128 of the 256 B rows) -- the tensor core reads the other half from the peer's
// shared memory.  Per CTA and K=64 stage that is 16 KB (A) + 16 KB (B half) of TMA fill instead of 16 + 32, and
// 4 + 4 KB of operand reads per MMA instead of 4 + 8: the shared-memory port, which caps the single-CTA kernel at
// ~2/3 of the tensor peak (TMA fill 94 B/clk + operand reads 96 B/clk against 128 B/clk), is no longer the limit,
// and the smaller stages allow a 6-deep ring.
//
// Roles per CTA: warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (LEADER CTA only) + TMEM owner,
// warps 2..9 = epilogue (both CTAs, on their own accumulator halves; policies are shared with pfc_umma.cuh).
// Barriers: full[s]  lives in the leader; both CTAs' TMA loads complete their bytes on it;
//           empty[s] in each CTA, released by the leader's multicast tcgen05.commit;
//           tmem_full[a] in each CTA (multicast commit); tmem_empty[a] in the leader, 2 x 8 epilogue warps arrive.
#pragma once
#include "pfc_umma.cuh"

namespace pfc {

constexpr int PAIR_B_STAGE_BYTES = B_STAGE_BYTES / 2;
constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + PAIR_B_STAGE_BYTES;
static_assert(PAIR_STAGE_BYTES == PAIR_STAGE_BYTES_, "pair stage size");

// Tiles 2T and 2T+1 (T = pair index) must share n0 and the K range and differ only in m0.
template <class P>
__global__ void __launch_bounds__(GemmCfg<P>::THREADS, 1)
umma_gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      const __grid_constant__ CUtensorMap tma_c, const typename P::Params prm) {
    using C = GemmCfg<P>;
    constexpr int PAIR_STAGES = C::PAIR_ST;
    constexpr int EPI_WARPS = C::EW;
    constexpr int EPI_COLS = C::COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + PAIR_STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + PAIR_STAGES * PAIR_STAGE_BYTES;

    __shared__ __align__(8) uint64_t full_bar[PAIR_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[PAIR_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = static_cast<int>(cluster_ctarank());
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_c);
#pragma unroll
        for (int s = 0; s < PAIR_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 2 * EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int first_tile = (blockIdx.x / 2) * 2 + crank;
    const int tile_stride = gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs; uniform loops,
        // one elected lane issues)
        uint32_t stage = 0, phase = 0;
        const uint32_t lbar0 = mapa_u32(smem_u32(&full_bar[0]), 0);          // the leader's full barriers
        for (int t = first_tile; t < prm.num_tiles; t += tile_stride) {
            const TileCoord tc = P::tile(prm, t);
            for (int kc = tc.k0; kc < tc.k1; ++kc) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one_sync()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * PAIR_STAGE_BYTES);
                    const uint32_t lbar = lbar0 + stage * 8;
                    uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
                    uint8_t* b_dst = sB + stage * PAIR_B_STAGE_BYTES;
                    const int kel = kc * BK;
                    if constexpr (P::A_MN) {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_a_pair<P::A_BLOCKED>(a_dst + j * MN_BOX_BYTES, &tma_a, lbar, tc.m0 + j * 64, kel);
                    } else {
                        tma_load_a_pair<P::A_BLOCKED>(a_dst, &tma_a, lbar, kel, tc.m0);
                    }
                    if constexpr (P::B_MN) {
#pragma unroll
                        for (int jj = 0; jj < BN / 128; ++jj) {     // this CTA's two 64-wide N blocks
                            const int j = crank * (BN / 128) + jj;
                            tma_load_2d_pair(b_dst + jj * MN_BOX_BYTES, &tma_b, lbar, tc.n0 + j * 64, kel);
                        }
                    } else {                                         // tensor map box = 128 B rows
                        tma_load_2d_pair(b_dst, &tma_b, lbar, kel, tc.n0 + crank * (BN / 2));
                    }
                }
                __syncwarp();
                if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only; the whole
        // warp walks the loops, one elected lane issues; descriptors = template + start address, see pfc_umma.cuh)
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, P::A_MN ? 1 : 0, P::B_MN ? 1 : 0) ^ prm.idesc_xor;
            const DescCfg dc = P::desc(prm);
            const uint64_t a_tmpl = umma_smem_desc_sw128(smem_u32(sA), dc.a_lbo, dc.a_sbo);
            const uint64_t b_tmpl = umma_smem_desc_sw128(smem_u32(sB), dc.b_lbo, dc.b_sbo);
            const uint32_t a_kstep = dc.a_kstep >> 4, b_kstep = dc.b_kstep >> 4;
            uint32_t stage = 0, phase = 0;
            int tl = 0;
            for (int t = first_tile; t < prm.num_tiles; t += tile_stride, ++tl) {
                const TileCoord tc = P::tile(prm, t);
                const int acc = tl & 1;
                const uint32_t acc_phase = (tl >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kc = tc.k0; kc < tc.k1; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t adesc = a_tmpl + stage * (A_STAGE_BYTES >> 4);
                        const uint64_t bdesc = b_tmpl + stage * (PAIR_B_STAGE_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16_ss_pair(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc,
                                              (kc > tc.k0 || k > 0) ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage], 0b11);      // frees the slot in BOTH CTAs
                    }
                    __syncwarp();
                    if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) umma_commit_pair(&tmem_full_bar[acc], 0b11);   // both CTAs' epilogues may start
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps, both CTAs)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const uint32_t stage_buf = smem_u32(sEpi + ew * EPI_STAGE_BYTES);
        int tl = 0;
        for (int t = first_tile; t < prm.num_tiles; t += tile_stride, ++tl) {
            const TileCoord tc = P::tile(prm, t);
            const int acc = tl & 1;
            const uint32_t acc_phase = (tl >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr =
                tmem_base + acc * BN + half * EPI_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
            P::epilogue(prm, tc, taddr, quarter, half, lane, stage_buf, &tma_c);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
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
