// Warp-specialised persistent tcgen05 GEMM skeleton shared by the three dense contractions of the head:
//   G1  logits tile  S  = Xn  . Wn^T      (A K-major,  B K-major,  epilogue = margin/exp/row-sum/bf16 spill)
//   G2  dXn          += E' . Wn           (A K-major,  B MN-major, split over classes, epilogue = fp32 partial)
//   G3  dWn           = E'^T . Xs         (A MN-major, B MN-major, epilogue = fp32 tile store)
// One CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
// warps 2..9 = epilogue: two warps per TMEM lane quarter, each owning one 128-column half of the accumulator.
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue, two 256-column accumulators),
// and the static persistent tile loop.
//
// Tile = 128 (TMEM lanes) x 256 (TMEM columns) fp32, K consumed in 64-element (128-byte, SWIZZLE_128B)
// stages, 4 stages x (16 KB A + 32 KB B) = 192 KB of shared memory, plus 4 KB per epilogue warp used to
// transpose "one thread = one row" register fragments into full 128-byte-line global stores.
#pragma once
#include "pfc_ptx.cuh"
#include "pfc_launch.cuh"

namespace pfc {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = BN * BK * 2;
constexpr int MN_BOX_BYTES = 64 * BK * 2;   // one 64(MN) x 64(K) box of an MN-major operand
constexpr int EPI_WARPS = 8;
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);   // columns owned by one epilogue warp (128)
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_STAGE_BYTES = 32 * 128;   // per epilogue warp: 32 rows x 128 B
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_WARPS * EPI_STAGE_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr int PAIR_STAGE_BYTES_ = A_STAGE_BYTES + B_STAGE_BYTES / 2;

// Per-policy launch geometry.  A policy picks the number of epilogue warps (8: two per TMEM lane quarter, 128 columns
// each; 16: four per quarter, 64 columns each -- for epilogues whose dependent chains need more warps per scheduler to
// hide latency) and the depth of the operand ring; 4 KB of staging per epilogue warp comes out of the same 227 KB.
template <class P>
struct GemmCfg {
    static constexpr int EW = P::EPI_WARPS_;
    static constexpr int COLS = BN / (EW / 4);
    static constexpr int ST = P::STAGES_;
    static constexpr int THREADS = 64 + 32 * EW;
    static constexpr int SMEM = ST * (A_STAGE_BYTES + B_STAGE_BYTES) + EW * EPI_STAGE_BYTES + 1024;
    static constexpr int PAIR_ST = P::PAIR_STAGES_;
    static constexpr int PAIR_SMEM = PAIR_ST * PAIR_STAGE_BYTES_ + EW * EPI_STAGE_BYTES + 1024;
    static_assert(EW == 8 || EW == 16, "epilogue warps: 8 or 16");
    static_assert(SMEM <= 232448 && PAIR_SMEM <= 232448, "shared memory budget");
};

struct TileCoord {
    int m0;    // first row of the A-side (TMEM lane) dimension
    int n0;    // first row of the B-side (TMEM column) dimension
    int k0;    // first K stage (units of BK)
    int k1;    // one past the last K stage
    int aux;   // policy-defined (e.g. split index / class-tile index)
};

// Shared-memory descriptor geometry of one operand stage (bytes).
struct DescCfg {
    uint32_t a_lbo, a_sbo, a_kstep;
    uint32_t b_lbo, b_sbo, b_kstep;
};
__host__ __device__ constexpr DescCfg default_desc_cfg(bool a_mn, bool b_mn) {
    return DescCfg{a_mn ? (uint32_t)MN_BOX_BYTES : 16u, 1024u, a_mn ? 2048u : 32u,
                   b_mn ? (uint32_t)MN_BOX_BYTES : 16u, 1024u, b_mn ? 2048u : 32u};
}

// Each lane holds 128 bytes of ITS row (v[0..31]).  The warp stages its 32 rows x 128 B in the SWIZZLE_128B layout
// of the store tensor map (16-byte piece j of row r at r*128 + ((j ^ (r & 7)) << 4); `stage` is 1024-byte aligned)
// and one lane hands the box to the TMA unit, which writes full lines and clips rows / columns beyond the tensor.
// Compared with per-thread LDS + STG this removes 16 memory instructions, their 64-bit address arithmetic and the
// bounds predicates per box from the epilogue warps, which pace the kernel when only two of them share a scheduler.
// (c0, c1, c2) = element coordinates of the box in the map's (inner, row, slab) dimensions.
// kKeep: the store carries an L2 evict_last hint (the consumer of this output runs next and should find it in L2).
template <bool kKeep = false>
__device__ __forceinline__ void warp_tma_store_rows(uint32_t stage, int lane, const uint32_t (&v)[32],
                                                    const CUtensorMap* map, int c0, int c1, int c2) {
    if (lane == 0) tma_store_wait_read<0>();   // the previous box of this warp has left the staging buffer
    __syncwarp();
    const uint32_t row = stage + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        sts128(row + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        if constexpr (kKeep) tma_store_3d_hint(map, stage, c0, c1, c2, l2_policy_evict_last());
        else tma_store_3d(map, stage, c0, c1, c2);
        tma_store_commit();
    }
}

// The same store in two halves, for epilogues that produce a row's 128 bytes as two 64-byte pieces and should not keep
// both in registers: rows_begin (staging buffer free), rows_half<0>, rows_half<1>, rows_commit.
__device__ __forceinline__ void warp_tma_store_begin(int lane) {
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
}
template <int kHalf>
__device__ __forceinline__ void warp_tma_store_half(uint32_t stage, int lane, const uint32_t (&v)[16]) {
    const uint32_t row = stage + lane * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        sts128(row + (((4 * kHalf + j) ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void warp_tma_store_commit(uint32_t stage, int lane, const CUtensorMap* map, int c0, int c1,
                                                      int c2) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_3d(map, stage, c0, c1, c2);
        tma_store_commit();
    }
}

// Policy contract:
//   static constexpr bool A_MN, B_MN;           operand major-ness in shared memory
//   static constexpr bool A_BLOCKED;            A is the class-blocked spill E'[n_pad/64][B][64] (3-D tensor map)
//   struct Params { int num_tiles; ... };
//   static DescCfg desc(const Params&);         smem descriptor geometry (default_desc_cfg(A_MN, B_MN))
//   static TileCoord tile(const Params&, int t);
//   static void epilogue(const Params&, const TileCoord&, uint32_t taddr, int quarter, int half, int lane,
//                        uint32_t stage, const CUtensorMap* tma_c);
//        stage = shared-memory address of this warp's 4 KB staging buffer, tma_c = the output's store tensor map;
//        COLS = GemmCfg<P>::COLS; `half` = index of the warp's column group;
//        taddr = TMEM address of this warp's lane quarter at column (half * COLS) of the tile's accumulator;
//        the warp owns rows [32*quarter, 32*quarter+32) x columns [half*COLS, (half+1)*COLS) of the tile.
//
// CL = 2 runs CTA pairs as a thread-block cluster working on two adjacent tiles that share one operand stage
// (P::SHARE_B: the pair differs in m0 and shares B; otherwise it differs in n0 and shares A).  Each CTA fetches
// half of the shared stage and TMA-multicasts it into both CTAs' shared memory, halving that operand's L2 traffic;
// a stage is recycled only after BOTH CTAs' MMAs have retired (multicast tcgen05.commit on the empty barrier).
template <class P, int CL>
__global__ void __launch_bounds__(GemmCfg<P>::THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const typename P::Params prm) {
    using C = GemmCfg<P>;
    constexpr int STAGES = C::ST;
    constexpr int EPI_WARPS = C::EW;
    constexpr int EPI_COLS = C::COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);

    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_c);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CL);       // one multicast commit from every CTA of the cluster
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], EPI_WARPS);   // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all();   // peer barriers must be initialised before any remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int crank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
    const int first_tile = (blockIdx.x / CL) * CL + crank;   // cluster c works on tiles CL*c .. CL*c+CL-1, then strides
    const int tile_stride = gridDim.x;                        // gridDim.x is a multiple of CL
    constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (uniform loops, elected lane)
        {
            uint32_t stage = 0, phase = 0;
            for (int t = first_tile; t < prm.num_tiles; t += tile_stride) {
                const TileCoord tc = P::tile(prm, t);
                for (int kc = tc.k0; kc < tc.k1; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                    uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
                    uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
                    const int kel = kc * BK;
                    constexpr bool kShareA = (CL > 1) && !P::SHARE_B;
                    constexpr bool kShareB = (CL > 1) && P::SHARE_B;
                    if constexpr (P::A_MN) {
                        constexpr int nb = BM / 64 / (kShareA ? CL : 1);      // boxes this CTA fetches
#pragma unroll
                        for (int jj = 0; jj < nb; ++jj) {
                            const int j = kShareA ? crank * nb + jj : jj;
                            if constexpr (kShareA)
                                tma_load_a_mcast<P::A_BLOCKED>(a_dst + j * MN_BOX_BYTES, &tma_a, &full_bar[stage], tc.m0 + j * 64, kel, kMask);
                            else
                                tma_load_a<P::A_BLOCKED>(a_dst + j * MN_BOX_BYTES, &tma_a, &full_bar[stage], tc.m0 + j * 64, kel);
                        }
                    } else {
                        if constexpr (kShareA)     // tensor map box = BM/CL rows
                            tma_load_a_mcast<P::A_BLOCKED>(a_dst + crank * (A_STAGE_BYTES / CL), &tma_a, &full_bar[stage], kel,
                                                           tc.m0 + crank * (BM / CL), kMask);
                        else
                            tma_load_a<P::A_BLOCKED>(a_dst, &tma_a, &full_bar[stage], kel, tc.m0);
                    }
                    if constexpr (P::B_MN) {
                        constexpr int nb = BN / 64 / (kShareB ? CL : 1);
#pragma unroll
                        for (int jj = 0; jj < nb; ++jj) {
                            const int j = kShareB ? crank * nb + jj : jj;
                            if constexpr (kShareB)
                                tma_load_2d_mcast(b_dst + j * MN_BOX_BYTES, &tma_b, &full_bar[stage], tc.n0 + j * 64, kel, kMask);
                            else
                                tma_load_2d(b_dst + j * MN_BOX_BYTES, &tma_b, &full_bar[stage], tc.n0 + j * 64, kel);
                        }
                    } else {
                        if constexpr (kShareB)     // tensor map box = BN/CL rows
                            tma_load_2d_mcast(b_dst + crank * (B_STAGE_BYTES / CL), &tma_b, &full_bar[stage], kel,
                                              tc.n0 + crank * (BN / CL), kMask);
                        else
                            tma_load_2d(b_dst, &tma_b, &full_bar[stage], kel, tc.n0);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the tile / stage loops (uniform control flow), one elected lane issues.  The
        // descriptors are "template + start address": only the 14-bit address field changes, so a K step is one add.
        // K-major : rows of 128 B, 8-row swizzle atoms 1024 B apart; +32 B per 16-element K step.
        // MN-major: 64(MN) x 8(K) atoms of 1024 B; next 64-wide MN block one box (8 KB) further,
        //           next 8 K rows 1024 B further; +2048 B per 16-row K step.
        const uint32_t idesc = umma_idesc_bf16(BM, BN, P::A_MN ? 1 : 0, P::B_MN ? 1 : 0) ^ prm.idesc_xor;
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
                    const uint64_t bdesc = b_tmpl + stage * (B_STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16_ss(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc,
                                     (kc > tc.k0 || k > 0) ? 1u : 0u);
                    // smem slot reusable once these MMAs retire (in every CTA that multicasts into it)
                    if constexpr (CL > 1) umma_commit_mcast(&empty_bar[stage], kMask);
                    else umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tmem_full_bar[acc]);     // accumulator complete -> epilogue
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int ew = warp - 2;
        const int quarter = warp & 3;                  // TMEM lanes [32q, 32q+32) are the only ones this warp may read
        const int half = ew >> 2;                      // which 128-column half of the accumulator
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
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (lane == 0) tma_store_wait_all();     // this warp's last boxes are in global memory before the CTA exits
    }

    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all();   // no CTA may exit while its peer can still multicast into it
    else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace pfc
