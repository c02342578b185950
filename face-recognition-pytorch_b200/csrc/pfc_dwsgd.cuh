// dW GEMM with the normalise-backward + SGD/momentum update + next step's bf16 shard fused into its epilogue
// (reference: autograd of nets/PartialFC.py:200-201 for the weight, torch.optim.SGD.step driven from
// model/FR_PartialFC.py:182-188, and the F.normalize of :200 for the NEXT step).
//
//   dWn[c,:] = sum_i E'[i,c] * Xs[i,:]                     (tcgen05, K = global batch)
//   dw  = (dWn - wn (wn . dWn)) / ||w||;  g = dw + wd w;  mom = mu mom + g;  w -= lr mom;  wn' = bf16(w / ||w||)
//
// The un-normalised gradient never leaves the chip: compared with pfc_backward_dw + pfc_dw_sgd this removes the
// bf16 dWn spill (write + read, 4 B per weight) and runs the tensor work underneath the HBM-bound update.
//
// (1) CTA pair.  The normalise-backward needs a dot product over the FULL row (d = 512 columns), but one 128-row x
// 512-column fp32 accumulator is all of TMEM.  So d is split over a thread-block cluster of 2: CTA r owns columns
// [256 r, 256 r + 256) of the same 128-class tile, with two 256-column accumulators so that the MMAs of tile t+1 run
// underneath the update of tile t.  The E'^T operand stage is shared (each CTA fetches half and TMA-multicasts it).
// Row-wise partial sums are exchanged through distributed shared memory: every epilogue warp stores its partials
// into BOTH CTAs' exchange arrays and arrives (release.cluster) on BOTH CTAs' mbarriers; 4 contributors per row.
//
// (2) One exchange.  The update is linear in (w, mom, dWn):  w' = alpha w + beta mom + gamma dWn  with
// alpha = 1 - lr (wd - gs dot inv^2), beta = -lr mu, gamma = -lr gs  (dot = w . dWn, inv = 1/||w||, gs = inv / scale),
// so ||w'||^2 is a quadratic form in the six row dots {ww, mm, gg, wm, wg, mg}.  Pass A accumulates all six, ONE
// exchange makes dot AND the new norm known, and pass B writes w', mom' and the normalised bf16 row in one go.
//
// (3) TMA in, TMA out.  w and mom are streamed by TMA (boxes of 32 rows x 32 fp32 columns, SWIZZLE_128B) into a
// 3-box ring per epilogue warp; "one thread = one row" reads its 128 bytes conflict-free, overwrites the box with
// the result and the same box is TMA-stored back.  No register staging of global data: three 4 KB boxes per warp
// (96 KB per SM) are in flight independent of register pressure -- a first version with per-thread loads of one
// 32-column chunk at a time had ~32 KB per SM in flight and ran at 2.3 TB/s.  Rows beyond n are zero-filled on load
// and clipped on store by the tensor maps.
#pragma once
#include "pfc_umma.cuh"

namespace pfc {

constexpr int DWS_D = 2 * BN;   // the fused kernel is specialised for d = 512
constexpr int DWS_BK = 32;      // K (sample) elements per operand stage
constexpr int DWS_STAGES = 3;      // 3 x 24 KB operand ring: the kernel is HBM-bound and the update needs the rest of smem
constexpr int DWS_A_STAGE = BM * DWS_BK * 2;        // 8 KB: two 64-class x 32-sample boxes
constexpr int DWS_B_STAGE = BN * DWS_BK * 2;        // 16 KB: four 64-column x 32-sample boxes
constexpr int DWS_MN_BOX = 64 * DWS_BK * 2;         // 4 KB
constexpr int DWS_RING = 3;                         // state boxes in flight per epilogue warp
constexpr int DWS_BOX = 32 * 128;                   // 32 rows x 32 fp32
constexpr int DWS_WN_BOX = 32 * 64;                 // 32 rows x 32 bf16 (no swizzle)
constexpr int DWS_SMEM_BYTES = DWS_STAGES * (DWS_A_STAGE + DWS_B_STAGE) + EPI_WARPS * (DWS_RING * DWS_BOX + DWS_WN_BOX) + 1024;

struct DwSgdParams {
    int num_class_tiles;
    int n;                    // classes (rows of w)
    int k_stages;             // ceil(B / DWS_BK)
    const float* inv_w;       // [n] 1/||w|| of the CURRENT weights
    float* inv_next;          // [n] 1/||w_new||
    float lr, momentum, wd, inv_grad_scale;
    int prefetch;             // bit 0: next tile's E' operand boxes -> L2, bit 1: next tile's w / mom boxes -> L2
    DescCfg dc;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// Pull one box of the tensor into L2 (no shared-memory destination, no completion tracking): used one tile ahead so
// that the latency-critical loads of the shallow shared-memory rings hit L2 instead of a saturated HBM.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c_inner), "r"(c_outer)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_saddr(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner,
                                                  int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// One thread = one row of the warp's row quarter.  Publishes the six partial dots of this warp's 128 columns into
// slot `slot` of BOTH CTAs' exchange arrays with st.async: each 4-byte store completes its bytes on the destination
// CTA's mbarrier, so no fence and no separate arrive are needed (a first version with fence.acq_rel.cluster +
// release arrives spent a quarter of the epilogue's stall samples in MEMBAR).  The barrier of a row quarter expects
// 4 contributors x 32 rows x 6 values x 4 B per tile; the warp with half == 0 posts the expectation.  The arrays are
// double-buffered by tile parity: a contributor can be at most one tile ahead of the slowest reader.
constexpr uint32_t DWS_XCH_BYTES = 4 * 32 * 6 * 4;
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr),
                 "r"(__float_as_uint(v)), "r"(cluster_bar)
                 : "memory");
}
__device__ __forceinline__ void pair_row_allreduce6(float (*xch)[4][BM], uint64_t* bar, uint32_t parity, bool post,
                                                    int slot, int r, int lane, int crank, float (&vals)[6]) {
    if (post && lane == 0) mbar_arrive_expect_tx(bar, DWS_XCH_BYTES);
    const uint32_t b_own = mapa_u32(smem_u32(bar), crank), b_peer = mapa_u32(smem_u32(bar), crank ^ 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const uint32_t my = smem_u32(&xch[k][slot][r]);
        st_async_f32(mapa_u32(my, crank), vals[k], b_own);
        st_async_f32(mapa_u32(my, crank ^ 1), vals[k], b_peer);
    }
    mbar_wait(bar, parity);
#pragma unroll
    for (int k = 0; k < 6; ++k) vals[k] = (xch[k][0][r] + xch[k][1][r]) + (xch[k][2][r] + xch[k][3][r]);
}

// Sequence of state boxes one epilogue warp streams: per class tile 16 boxes --
//   pass A: (w, mom) of column chunks 0..3, pass B: the same eight boxes again (L2 hits).
struct BoxCursor {
    int ct;       // class tile
    int k;        // 0..15 inside the tile: pass = k >> 3, chunk = (k & 7) >> 1, which = k & 1 (0 = w, 1 = mom)
    int box;      // ring slot
    __device__ __forceinline__ void advance(int ct_stride) {
        if (++k == 16) { k = 0; ct += ct_stride; }
        if (++box == DWS_RING) box = 0;
    }
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
dw_sgd_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_m,
                   const __grid_constant__ CUtensorMap tm_wn, const DwSgdParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + DWS_STAGES * DWS_A_STAGE;
    uint8_t* sRing = smem + DWS_STAGES * (DWS_A_STAGE + DWS_B_STAGE);
    uint8_t* sWn = sRing + EPI_WARPS * DWS_RING * DWS_BOX;

    __shared__ __align__(8) uint64_t full_bar[DWS_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[DWS_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t xch_bar[4][2];                // per row quarter and tile parity
    __shared__ __align__(8) uint64_t ring_bar[EPI_WARPS][DWS_RING];
    __shared__ float xch[2][6][4][BM];   // [tile parity][ww mm gg wm wg mg][contributor = 2*cta + half][row of the tile]
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_m);
        tma_prefetch_desc(&tm_wn);
#pragma unroll
        for (int s = 0; s < DWS_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 2);            // one multicast commit from each CTA of the pair
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], EPI_WARPS);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {               // one expect_tx arrive; the data arrives as tx bytes
            mbar_init(&xch_bar[q][0], 1);
            mbar_init(&xch_bar[q][1], 1);
        }
        for (int i = 0; i < EPI_WARPS * DWS_RING; ++i) mbar_init(&ring_bar[0][0] + i, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int crank = static_cast<int>(cluster_ctarank());
    const int first_ct = blockIdx.x >> 1;              // class tile of this pair
    const int ct_stride = gridDim.x >> 1;
    const int n0 = crank * BN;                         // this CTA's 256 columns of d
    constexpr uint16_t kMask = 0b11;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (GEMM operands)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int ct = first_ct; ct < prm.num_class_tiles; ct += ct_stride) {
                const int m0 = ct * BM;
                const int m0_next = (ct + ct_stride < prm.num_class_tiles) ? (ct + ct_stride) * BM : -1;
                for (int kc = 0; kc < prm.k_stages; ++kc) {
                    // E' comes from HBM: this CTA's half of the NEXT tile's stage goes to L2 a whole tile ahead
                    if (m0_next >= 0 && (prm.prefetch & 1)) tma_prefetch_3d(&tma_a, 0, kc * DWS_BK, (m0_next >> 6) + crank);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], DWS_A_STAGE + DWS_B_STAGE);
                    uint8_t* a_dst = sA + stage * DWS_A_STAGE;
                    uint8_t* b_dst = sB + stage * DWS_B_STAGE;
                    const int kel = kc * DWS_BK;
                    // E'^T stage (128 classes x 32 samples) = two 64-class boxes: this CTA fetches one, both get both
                    tma_load_a_mcast<true>(a_dst + crank * DWS_MN_BOX, &tma_a, &full_bar[stage], m0 + crank * 64, kel, kMask);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_2d(b_dst + j * DWS_MN_BOX, &tma_b, &full_bar[stage], n0 + j * 64, kel);
                    if (++stage == DWS_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (uniform loops, elected lane)
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 1, 1);
        const DescCfg dc = prm.dc;
        const uint64_t a_tmpl = umma_smem_desc_sw128(smem_u32(sA), dc.a_lbo, dc.a_sbo);
        const uint64_t b_tmpl = umma_smem_desc_sw128(smem_u32(sB), dc.b_lbo, dc.b_sbo);
        const uint32_t a_kstep = dc.a_kstep >> 4, b_kstep = dc.b_kstep >> 4;
        uint32_t stage = 0, phase = 0;
        int tl = 0;
        for (int ct = first_ct; ct < prm.num_class_tiles; ct += ct_stride, ++tl) {
            const int acc = tl & 1;
            const uint32_t acc_phase = (tl >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kc = 0; kc < prm.k_stages; ++kc) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t adesc = a_tmpl + stage * (DWS_A_STAGE >> 4);
                    const uint64_t bdesc = b_tmpl + stage * (DWS_B_STAGE >> 4);
#pragma unroll
                    for (int k = 0; k < DWS_BK / UMMA_K; ++k)
                        umma_bf16_ss(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                    umma_commit_mcast(&empty_bar[stage], kMask);   // the peer multicasts into this slot too
                }
                __syncwarp();
                if (++stage == DWS_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tmem_full_bar[acc]);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue: fused update (8 warps)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const int slot = crank * 2 + half;
        const int r = quarter * 32 + lane;                  // row inside the class tile
        const int col0 = n0 + half * EPI_COLS;              // first of this warp's 128 columns of d
        const uint32_t ring = smem_u32(sRing + ew * DWS_RING * DWS_BOX);
        const uint32_t wn_stage = smem_u32(sWn + ew * DWS_WN_BOX);
        uint64_t* rbar = &ring_bar[ew][0];
        const uint32_t my_row = lane * 128;                 // this lane's row inside a box
        const int sw = lane & 7;                            // SWIZZLE_128B: piece j of row r sits at j ^ (r & 7)
        const float lr = prm.lr, mu = prm.momentum, wd = prm.wd;

        // issue side of the ring (lane 0 only): `freed` boxes have just been released, top the ring up again
        BoxCursor iss = {first_ct, 0, 0};
        int in_flight = 0;
        auto refill = [&](int freed) {
            if (lane == 0) {
                in_flight -= freed;
                while (in_flight < DWS_RING && iss.ct < prm.num_class_tiles) {
                    const int c = (iss.k & 7) >> 1;
                    mbar_arrive_expect_tx(&rbar[iss.box], DWS_BOX);
                    tma_load_2d_saddr(ring + iss.box * DWS_BOX, (iss.k & 1) ? &tm_m : &tm_w, &rbar[iss.box], col0 + 32 * c,
                                      iss.ct * BM + quarter * 32);
                    iss.advance(ct_stride);
                    ++in_flight;
                }
            }
        };
        uint32_t cbox = 0, cpar = 0;     // consume side (whole warp): ring slot and its mbarrier parity
        refill(0);

        int tl = 0;
        for (int ct = first_ct; ct < prm.num_class_tiles; ct += ct_stride, ++tl) {
            const int acc = tl & 1;
            const uint32_t acc_phase = (tl >> 1) & 1;
            const uint32_t xpar = tl & 1;
            const int row = ct * BM + r;
            const int row_box = ct * BM + quarter * 32;
            const bool ok = row < prm.n;
            const float inv = ok ? prm.inv_w[row] : 0.f;
            if (lane < 8 && (prm.prefetch & 2) && ct + ct_stride < prm.num_class_tiles)     // next tile's w / mom boxes of this warp -> L2
                tma_prefetch_2d((lane & 1) ? &tm_m : &tm_w, col0 + 32 * (lane >> 1), (ct + ct_stride) * BM + quarter * 32);
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BN + half * EPI_COLS + (static_cast<uint32_t>(quarter * 32) << 16);

            // ---- pass A: the six row dots over this warp's 128 columns
            float dots[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // ww mm gg wm wg mg
#pragma unroll 1
            for (int c = 0; c < EPI_COLS / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                const uint32_t bw = ring + cbox * DWS_BOX + my_row;
                mbar_wait(&rbar[cbox], cpar);
                if (++cbox == DWS_RING) { cbox = 0; cpar ^= 1; }
                const uint32_t bm = ring + cbox * DWS_BOX + my_row;
                mbar_wait(&rbar[cbox], cpar);
                if (++cbox == DWS_RING) { cbox = 0; cpar ^= 1; }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 wq = lds128(bw + ((j ^ sw) << 4));
                    const uint4 mq = lds128(bm + ((j ^ sw) << 4));
                    const float w4[4] = {__uint_as_float(wq.x), __uint_as_float(wq.y), __uint_as_float(wq.z), __uint_as_float(wq.w)};
                    const float m4[4] = {__uint_as_float(mq.x), __uint_as_float(mq.y), __uint_as_float(mq.z), __uint_as_float(mq.w)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float g = __uint_as_float(v[4 * j + u]);
                        dots[0] = fmaf(w4[u], w4[u], dots[0]);
                        dots[1] = fmaf(m4[u], m4[u], dots[1]);
                        dots[2] = fmaf(g, g, dots[2]);
                        dots[3] = fmaf(w4[u], m4[u], dots[3]);
                        dots[4] = fmaf(w4[u], g, dots[4]);
                        dots[5] = fmaf(m4[u], g, dots[5]);
                    }
                }
                __syncwarp();                 // every lane has read both boxes: they may be refilled
                refill(2);
            }
            // barrier and arrays alternate with the tile parity, so bytes of tile t+1 from a fast contributor can never
            // be counted into a CTA's still-open phase of tile t
            pair_row_allreduce6(xch[xpar], &xch_bar[quarter][xpar], acc_phase, half == 0, slot, r, lane, crank, dots);
            const float wscale = dots[4] * inv * inv;          // (wn . dWn) wn = w * (dot * inv^2)
            const float gs = inv * prm.inv_grad_scale;
            const float alpha = 1.f - lr * (wd - gs * wscale), beta = -lr * mu, gamma = -lr * gs;
            const float n2 = alpha * alpha * dots[0] + beta * beta * dots[1] + gamma * gamma * dots[2] +
                             2.f * (alpha * beta * dots[3] + alpha * gamma * dots[4] + beta * gamma * dots[5]);
            const float rden = 1.f / fmaxf(sqrtf(fmaxf(n2, 0.f)), 1e-12f);

            // ---- pass B: momentum + weight update in place in the boxes, normalised bf16 row; all stored by TMA
#pragma unroll 1
            for (int c = 0; c < EPI_COLS / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                const uint32_t boxw = ring + cbox * DWS_BOX;
                mbar_wait(&rbar[cbox], cpar);
                if (++cbox == DWS_RING) { cbox = 0; cpar ^= 1; }
                const uint32_t boxm = ring + cbox * DWS_BOX;
                mbar_wait(&rbar[cbox], cpar);
                if (++cbox == DWS_RING) { cbox = 0; cpar ^= 1; }
                tmem_ld_wait();
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t aw = boxw + my_row + ((j ^ sw) << 4), am = boxm + my_row + ((j ^ sw) << 4);
                    const uint4 wq = lds128(aw);
                    const uint4 mq = lds128(am);
                    float q[4] = {__uint_as_float(wq.x), __uint_as_float(wq.y), __uint_as_float(wq.z), __uint_as_float(wq.w)};
                    float b[4] = {__uint_as_float(mq.x), __uint_as_float(mq.y), __uint_as_float(mq.z), __uint_as_float(mq.w)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float g = __uint_as_float(v[4 * j + u]);
                        b[u] = fmaf(mu, b[u], fmaf(fmaf(-q[u], wscale, g), gs, wd * q[u]));
                        q[u] = fmaf(-lr, b[u], q[u]);
                    }
                    sts128(am, __float_as_uint(b[0]), __float_as_uint(b[1]), __float_as_uint(b[2]), __float_as_uint(b[3]));
                    sts128(aw, __float_as_uint(q[0]), __float_as_uint(q[1]), __float_as_uint(q[2]), __float_as_uint(q[3]));
                    o[(j & 1) * 2] = pack_bf16x2(q[0] * rden, q[1] * rden);
                    o[(j & 1) * 2 + 1] = pack_bf16x2(q[2] * rden, q[3] * rden);
                    if (j & 1) sts128(wn_stage + lane * 64 + ((j >> 1) << 4), o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_w, boxw, col0 + 32 * c, row_box);
                    tma_store_2d(&tm_m, boxm, col0 + 32 * c, row_box);
                    tma_store_2d(&tm_wn, wn_stage, col0 + 32 * c, row_box);
                    tma_store_commit();
                    tma_store_wait_read<0>();     // the boxes and the bf16 staging buffer have been read out
                }
                refill(2);
                __syncwarp();                     // nobody writes the staging buffer before lane 0's wait returned
            }
            // the accumulator is no longer needed: the MMAs of the tile after next may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if (ok && slot == 0) prm.inv_next[row] = rden;
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();   // no CTA may exit while its peer can still multicast / store into it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace pfc
