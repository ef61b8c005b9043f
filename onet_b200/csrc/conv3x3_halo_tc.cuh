// 3x3-convolution specialisations of the tap-GEMM kernels that cut the L2 -> SM operand traffic.
//
// With a pixel tile that is exactly 8 pixels wide, one row of the tile is one 1024-byte swizzle atom (8 rows x
// 128 B) in shared memory.  A box fetched with a one-row halo above and below therefore serves the three vertical
// taps (kh = 0,1,2) of a filter column by moving the UMMA descriptor start address in steps of 1024 B — always
// atom-aligned, so the canonical 128-byte-swizzle layout stays valid.  Only the three horizontal shifts (kw) need
// their own TMA load:  3 loads of (TH+2) x 8 pixels instead of 9 loads of TH x 8 pixels.
//
//   conv3x3_halo_px_kernel : forward / data-gradient, tile = 16 x 8 pixels (M = 128), separate smem rings for the
//                            activation boxes (18 KB each) and the per-tap weight tiles.
//   wgrad3x3_halo_kernel   : weight gradient, pixel slab = 8 x 8 (K = 64), the shifted operand is the output
//                            gradient G; accumulator rows 0..63 / 64..127 may come from two different taps
//                            (Cout = 64) or two adjacent 64-channel blocks (Cout >= 128).
#pragma once
#include "tapgemm_tc.cuh"

namespace onet {

template <int BN>
struct HaloCfg {
    static constexpr int kABytes = 144 * 128;                 // (16 + 2) rows x 8 pixels x one 128-byte K chunk of channels
    static constexpr int kBBytes = BN * 128;
    static constexpr int kTapsPerB = (BN == 256) ? 1 : 3;     // vertical taps fetched per weight stage
    static constexpr int kSA = (BN == 64) ? 4 : 3;
    static constexpr int kSB = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
    static constexpr int kBStageBytes = kTapsPerB * kBBytes;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kAuxBytes = 1024 + 4 * 2 * BN * 4;
    static constexpr int kSmemBytes = kSA * kABytes + kSB * kBStageBytes + kAuxBytes + 1024;
};

template <int BN, class Op = OpBf16>
__global__ void __launch_bounds__(kPxThreads, 1)
conv3x3_halo_px_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PxParams p) {
    using Cfg = HaloCfg<BN>;
    constexpr int KC = Op::kKC;
    constexpr int SA = Cfg::kSA, SB = Cfg::kSB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t ringA = base, ringB = base + SA * Cfg::kABytes;
    constexpr int TPB = Cfg::kTapsPerB;
    const uint32_t aux = ringB + SB * Cfg::kBStageBytes;
    uint8_t* gen_aux = gen_base + SA * Cfg::kABytes + SB * Cfg::kBStageBytes;
    const uint32_t bar_fullA = aux, bar_emptyA = aux + 8 * SA, bar_fullB = aux + 16 * SA, bar_emptyB = bar_fullB + 8 * SB,
                   bar_tfull = bar_emptyB + 8 * SB, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * SA + 16 * SB + 32);
    float* s_part = reinterpret_cast<float*>(gen_aux + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(bar_fullA + 8 * s, 1); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, kPxEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0) {
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = fd_div(tile, p.fd_m), m_tile = tile - n_tile * p.num_m_tiles;
                int wt, ht, nt;
                px_tile_coord(p, m_tile, wt, ht, nt);
                const int w0 = wt * 8, h0 = ht * 16, co0 = n_tile * BN;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    for (int kw = 0; kw < 3; ++kw) {
                        mbar_wait(bar_emptyA + 8 * sa, pa ^ 1);
                        mbar_expect_tx(bar_fullA + 8 * sa, Cfg::kABytes);
                        tma_load_5d(ringA + sa * Cfg::kABytes, &tmA, bar_fullA + 8 * sa, kc * KC, w0 + kw - 1, 0, h0 - 1, nt);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                        for (int kh = 0; kh < 3; kh += TPB) {
                            mbar_wait(bar_emptyB + 8 * sb, pb ^ 1);
                            mbar_expect_tx(bar_fullB + 8 * sb, Cfg::kBStageBytes);
#pragma unroll
                            for (int j = 0; j < TPB; ++j)
                                tma_load_2d(ringB + sb * Cfg::kBStageBytes + j * Cfg::kBBytes, &tmB, bar_fullB + 8 * sb,
                                            ((kh + j) * 3 + kw) * p.cin + kc * KC, co0);
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // one elected thread runs the whole issue loop (no per-iteration warp convergence, no uniform-datapath loops)
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc<Op>(128, BN, 0, 0);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    for (int kw = 0; kw < 3; ++kw) {
                        mbar_wait(bar_fullA + 8 * sa, pa);
                        const uint32_t sA = ringA + sa * Cfg::kABytes;
                        for (int kh = 0; kh < 3; kh += TPB) {
                            mbar_wait(bar_fullB + 8 * sb, pb);
                            tc_fence_after();
#pragma unroll
                            for (int j = 0; j < TPB; ++j) {
                                const uint64_t da = umma_smem_desc(sA + (kh + j) * 1024, 16, 1024);
                                const uint64_t db = umma_smem_desc(ringB + sb * Cfg::kBStageBytes + j * Cfg::kBBytes, 16, 1024);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    umma<Op>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                            umma_commit(bar_emptyB + 8 * sb);
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        umma_commit(bar_emptyA + 8 * sa);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, ew = warp - 2;
        PxStatAcc sacc;
        sacc.reset(-1);
        sacc.reset_rows();
        sacc.grp = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n_tile = fd_div(tile, p.fd_m), m_tile = tile - n_tile * p.num_m_tiles;
            const PxRowCoord rc = px_row_coord(p, m_tile, q, lane);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            px_store_epilogue<BN, false, Op>(p, rc, n_tile, acc, tmem_base, q, ew, lane, s_part, bar_tempty + 8 * acc, false, sacc);
        }
        px_stat_flush<BN>(p, sacc, ew, lane, s_part);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// =====================================================================================================
// Weight-resident variant for Cin = 64 (one K chunk) and Cout = BN <= 128: the whole packed weight tensor (9 taps x BN
// rows x 128 B = 72 / 144 KB) is loaded into shared memory ONCE per persistent CTA and only the activation boxes
// stream.  For these layers (the full-resolution ones) the weight re-fetch was 21-30 % of the shared-memory-port
// traffic that bounds the kernel.
// =====================================================================================================
template <int BN>
struct HaloResCfg {
    static constexpr int kABytes = 144 * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kSA = (BN == 64) ? 7 : 4;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kAuxBytes = 1024 + 4 * 2 * BN * 4;
    static constexpr int kSmemBytes = kSA * kABytes + 9 * kBBytes + kAuxBytes + 1024;
};

template <int BN, bool RED = false>
__global__ void __launch_bounds__(kPxThreads, 1)
conv3x3_halo_res_px_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PxParams p) {
    using Cfg = HaloResCfg<BN>;
    constexpr int SA = Cfg::kSA;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t ringA = base, resB = base + SA * Cfg::kABytes;
    const uint32_t aux = resB + 9 * Cfg::kBBytes;
    uint8_t* gen_aux = gen_base + SA * Cfg::kABytes + 9 * Cfg::kBBytes;
    const uint32_t bar_fullA = aux, bar_emptyA = aux + 8 * SA, bar_bres = aux + 16 * SA, bar_tfull = bar_bres + 8,
                   bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * SA + 8 + 32 + 8);
    float* s_part = reinterpret_cast<float*>(gen_aux + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(bar_fullA + 8 * s, 1); mbar_init(bar_emptyA + 8 * s, 1); }
        mbar_init(bar_bres, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, kPxEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int num_tiles = p.num_m_tiles;       // one n-tile, one K chunk

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar_bres, 9 * Cfg::kBBytes);
#pragma unroll
            for (int t = 0; t < 9; ++t) tma_load_2d(resB + t * Cfg::kBBytes, &tmB, bar_bres, t * p.cin, 0);
            int sa = 0;
            uint32_t pa = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int wt, ht, nt;
                px_tile_coord(p, tile, wt, ht, nt);
                const int w0 = wt * 8, h0 = ht * 16;
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_wait(bar_emptyA + 8 * sa, pa ^ 1);
                    mbar_expect_tx(bar_fullA + 8 * sa, Cfg::kABytes);
                    tma_load_5d(ringA + sa * Cfg::kABytes, &tmA, bar_fullA + 8 * sa, 0, w0 + kw - 1, 0, h0 - 1, nt);
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
            mbar_wait(bar_bres, 0);
            int sa = 0;
            uint32_t pa = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_wait(bar_fullA + 8 * sa, pa);
                    tc_fence_after();
                    const uint32_t sA = ringA + sa * Cfg::kABytes;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint64_t da = umma_smem_desc(sA + kh * 1024, 16, 1024);
                        const uint64_t db = umma_smem_desc(resB + (kh * 3 + kw) * Cfg::kBBytes, 16, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(bar_emptyA + 8 * sa);
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, ew = warp - 2;
        PxStatAcc sacc;
        sacc.reset(-1);
        sacc.reset_rows();
        sacc.grp = 0;
        int it = 0;
        if (RED) px_red_prefetch<BN>(p, blockIdx.x, 0, q, ew, lane);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            if (RED && tile + static_cast<int>(gridDim.x) < num_tiles) px_red_prefetch<BN>(p, tile + gridDim.x, 0, q, ew, lane);
            const PxRowCoord rc = px_row_coord(p, tile, q, lane);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            px_store_epilogue<BN, RED>(p, rc, 0, acc, tmem_base, q, ew, lane, s_part, bar_tempty + 8 * acc, false, sacc);
        }
        px_stat_flush<BN>(p, sacc, ew, lane, s_part);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// =====================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes TWO adjacent pixel tiles (M = 256) x BN couts per
// tcgen05.mma.  Each CTA stages its own activation boxes and only HALF of the weight tile (BN/2 rows), which cuts
// the shared-memory-port traffic per CTA (TMA fills + operand reads) by 21-37 % - the resource that bounds the
// single-CTA kernel.  The leader CTA (cluster rank 0) issues every MMA; TMA completions of both
// CTAs are signalled on the leader's "full" barriers, MMA completions are multicast to both CTAs' "empty" /
// "accumulator full" barriers, and both CTAs' epilogue warps arrive on the leader's "accumulator empty" barrier.
// =====================================================================================================
template <int BN>
struct Halo2Cfg {
    static constexpr int kABytes = 144 * 128;
    static constexpr int kBHalfBytes = (BN / 2) * 128;
    static constexpr int kTapsPerB = (BN == 256) ? 1 : 3;
    static constexpr int kBStageBytes = kTapsPerB * kBHalfBytes;
    static constexpr int kSA = (BN == 256) ? 4 : (BN == 128 ? 5 : 6);
    static constexpr int kSB = (BN == 256) ? 6 : (BN == 128 ? 4 : 6);
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kAuxBytes = 1024 + 4 * 2 * BN * 4;
    static constexpr int kSmemBytes = kSA * kABytes + kSB * kBStageBytes + kAuxBytes + 1024;
};

template <int BN, bool RED = false, class Op = OpBf16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPxThreads, 1)
conv3x3_halo2_px_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PxParams p) {
    using Cfg = Halo2Cfg<BN>;
    constexpr int KC = Op::kKC;
    constexpr int SA = Cfg::kSA, SB = Cfg::kSB, TPB = Cfg::kTapsPerB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t ringA = base, ringB = base + SA * Cfg::kABytes;
    const uint32_t aux = ringB + SB * Cfg::kBStageBytes;
    uint8_t* gen_aux = gen_base + SA * Cfg::kABytes + SB * Cfg::kBStageBytes;
    const uint32_t bar_fullA = aux, bar_emptyA = aux + 8 * SA, bar_fullB = aux + 16 * SA, bar_emptyB = bar_fullB + 8 * SB,
                   bar_tfull = bar_emptyB + 8 * SB, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * SA + 16 * SB + 32);
    float* s_part = reinterpret_cast<float*>(gen_aux + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(bar_fullA + 8 * s, 1); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * kPxEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised and its TMEM is allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_pair_m = (p.num_m_tiles + 1) >> 1;
    const int num_units = num_pair_m * p.num_n_tiles;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t lead_fullA = mapa_shared(bar_fullA, 0), lead_fullB = mapa_shared(bar_fullB, 0);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int unit = pair; unit < num_units; unit += npairs) {
                const int n_tile = fd_div(unit, p.fd_pm), m_tile = 2 * (unit - n_tile * num_pair_m) + static_cast<int>(rank);
                // an m_tile == num_m_tiles (odd tile count) decodes to image index N: every row is out of bounds -> zeros
                int wt, ht, nt;
                px_tile_coord(p, m_tile, wt, ht, nt);
                const int w0 = wt * 8, h0 = ht * 16, co0 = n_tile * BN + static_cast<int>(rank) * (BN / 2);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    for (int kw = 0; kw < 3; ++kw) {
                        mbar_wait(bar_emptyA + 8 * sa, pa ^ 1);
                        if (rank == 0) mbar_expect_tx(bar_fullA + 8 * sa, 2 * Cfg::kABytes);
                        tma_load_5d_2cta(ringA + sa * Cfg::kABytes, &tmA, lead_fullA + 8 * sa, kc * KC, w0 + kw - 1, 0, h0 - 1, nt);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                        for (int kh = 0; kh < 3; kh += TPB) {
                            mbar_wait(bar_emptyB + 8 * sb, pb ^ 1);
                            if (rank == 0) mbar_expect_tx(bar_fullB + 8 * sb, 2 * Cfg::kBStageBytes);
#pragma unroll
                            for (int j = 0; j < TPB; ++j)
                                tma_load_2d_2cta(ringB + sb * Cfg::kBStageBytes + j * Cfg::kBHalfBytes, &tmB, lead_fullB + 8 * sb,
                                                 ((kh + j) * 3 + kw) * p.cin + kc * KC, co0);
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc<Op>(256, BN, 0, 0);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int unit = pair; unit < num_units; unit += npairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    for (int kw = 0; kw < 3; ++kw) {
                        mbar_wait(bar_fullA + 8 * sa, pa);
                        const uint32_t sA = ringA + sa * Cfg::kABytes;
                        for (int kh = 0; kh < 3; kh += TPB) {
                            mbar_wait(bar_fullB + 8 * sb, pb);
                            tc_fence_after();
#pragma unroll
                            for (int j = 0; j < TPB; ++j) {
                                const uint64_t da = umma_smem_desc(sA + (kh + j) * 1024, 16, 1024);
                                const uint64_t db = umma_smem_desc(ringB + sb * Cfg::kBStageBytes + j * Cfg::kBHalfBytes, 16, 1024);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    umma_2cta<Op>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                            umma_commit_2cta(bar_emptyB + 8 * sb);
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        umma_commit_2cta(bar_emptyA + 8 * sa);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
                umma_commit_2cta(bar_tfull + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3, ew = warp - 2;
        const uint32_t lead_tempty = mapa_shared(bar_tempty, 0);
        PxStatAcc sacc;
        sacc.reset(-1);
        sacc.reset_rows();
        sacc.grp = 0;
        int it = 0;
        if (RED && pair < num_units)
            px_red_prefetch<BN>(p, 2 * (pair - fd_div(pair, p.fd_pm) * num_pair_m) + static_cast<int>(rank), fd_div(pair, p.fd_pm), q, ew, lane);
        for (int unit = pair; unit < num_units; unit += npairs, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n_tile = fd_div(unit, p.fd_pm), m_tile = 2 * (unit - n_tile * num_pair_m) + static_cast<int>(rank);
            if (RED && unit + npairs < num_units) {
                const int nn = fd_div(unit + npairs, p.fd_pm);
                px_red_prefetch<BN>(p, 2 * (unit + npairs - nn * num_pair_m) + static_cast<int>(rank), nn, q, ew, lane);
            }
            const PxRowCoord rc = px_row_coord(p, min(m_tile, p.num_m_tiles - 1), q, lane);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            if (m_tile < p.num_m_tiles) {
                px_store_epilogue<BN, RED, Op>(p, rc, n_tile, acc, tmem_base, q, ew, lane, s_part, lead_tempty + 8 * acc, true, sacc);
            } else {           // padding tile of an odd tile count: nothing to store
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * acc);
            }
        }
        px_stat_flush<BN>(p, sacc, ew, lane, s_part);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // no CTA frees tensor memory or exits while its peer may still signal / multicast to it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
    }
}

// =====================================================================================================
// weight gradient with H-halo boxes on the shifted (G) operand
// =====================================================================================================
constexpr int kWhMaxBox = 4;
constexpr int kWhMaxAcc = 5;
constexpr int kWhBoxBytes = 80 * 128;        // (8 + 2) rows x 8 pixels x 64 channels bf16

struct WhAcc {
    int start_off;            // byte offset of rows 0..63 of the accumulator's M operand inside the stage's M region
    int lbo;                  // byte distance to rows 64..127
    int tapA, chA, tapB, chB; // (tap, channel offset inside the m-tile) of the two row blocks; tapB < 0: unused
};
struct WhUnit {
    int nbox;
    int box_kw[kWhMaxBox], box_ch[kWhMaxBox];
    int nacc;
    WhAcc acc[kWhMaxAcc];
};
struct WhParams {
    int N, H, W;
    int tiles_w, tiles_h, num_px_tiles;       // 8x8 pixel slabs
    int ksplit, px_tiles_per_split;
    int ntypes, num_m_tiles, num_n_tiles, m_tile_channels;
    WhUnit types[3];
    int ks_slowest;                           // unit order: all (type, m, n) tiles of one pixel range run together (L2 reuse)
    float* out;                               // dW [m_total][n_total][9], accumulated with atomics
    int m_total, n_total;
};

// XS = 3 ("x-shifted" form, BNW = 64): the filter COLUMN kw is carried by the unshifted operand instead - its 64 channels are
// loaded three times, shifted by kw - 1 pixels in w, and form N = 3 x 64 = 192 columns (column block = kw), while G is only
// row-shifted (h-halo) and one accumulator holds two filter rows (M = 2 x 64):
//     dW[co][ci][kh][kw] = sum_px G[px - (kh-1, 0)][co] * In[px + (0, kw-1)][ci]
// A 64-channel G block then costs 8 MMAs of N = 192 per pixel slab (10 KB of operand reads per 96-clock MMA: tensor-bound)
// instead of 20 MMAs of N = 64 (6 KB per 32-clock MMA: bound by the shared-memory pipe at 2/3 of the tensor rate).
template <int BNW, int XS = 1>
struct WhCfg {
    static constexpr int kNCols = BNW * XS;
    static constexpr int kNBytes = (kNCols / 64) * 8192;
    static constexpr int kMBytes = (XS == 1 ? kWhMaxBox : 2) * kWhBoxBytes;     // 40 KB / 20 KB
    static constexpr int kStageBytes = kNBytes + kMBytes;
    static constexpr int kStages = XS == 1 ? 3 : 4;
    static constexpr int kTmemCols = 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
};

template <int BNW, int XS = 1>
__global__ void __launch_bounds__(192, 1)
wgrad3x3_halo_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI, const WhParams p) {
    static_assert(XS == 1 || (XS == 3 && BNW == 64), "x-shifted form: three 64-channel column blocks");
    using Cfg = WhCfg<BNW, XS>;
    constexpr int NCOLS = Cfg::kNCols;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t aux = base + STAGES * Cfg::kStageBytes;
    uint8_t* gen_aux = gen_base + STAGES * Cfg::kStageBytes;
    const uint32_t bar_full = aux, bar_empty = aux + 8 * STAGES, bar_tfull = aux + 16 * STAGES, bar_tempty = bar_tfull + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmI);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // unit = ((type * num_m_tiles + mt) * num_n_tiles + nt) * ksplit + ks
    const int num_units = p.ntypes * p.num_m_tiles * p.num_n_tiles * p.ksplit;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const int nt = (unit / p.ksplit) % p.num_n_tiles;
                const int mt = (unit / (p.ksplit * p.num_n_tiles)) % p.num_m_tiles;
                const WhUnit& u = p.types[unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles)];
                const int m0 = mt * p.m_tile_channels, n0 = nt * BNW;
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                const uint32_t tx = static_cast<uint32_t>(u.nbox) * kWhBoxBytes + Cfg::kNBytes;
                for (int pt = px_begin; pt < px_end; ++pt) {
                    const int wt = pt % p.tiles_w, ht = (pt / p.tiles_w) % p.tiles_h, n = pt / (p.tiles_w * p.tiles_h);
                    const int w0 = wt * 8, h0 = ht * 8;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    const uint32_t fb = bar_full + 8 * stage;
                    mbar_expect_tx(fb, tx);
                    if constexpr (XS == 1) {
#pragma unroll
                        for (int b = 0; b < BNW / 64; ++b) tma_load_5d(sN + b * 8192, &tmI, fb, n0 + b * 64, w0, 0, h0, n);
                    } else {
#pragma unroll
                        for (int b = 0; b < XS; ++b) tma_load_5d(sN + b * 8192, &tmI, fb, n0, w0 + b - 1, 0, h0, n);   // column block = kw
                    }
                    for (int b = 0; b < u.nbox; ++b)     // G shifted by -(kw-1) in w, one-row halo in h
                        tma_load_5d(sM + b * kWhBoxBytes, &tmG, fb, m0 + u.box_ch[b], w0 - (u.box_kw[b] - 1), 0, h0 - 1, n);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, NCOLS, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x, ++it) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const WhUnit& u = p.types[unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles)];
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                mbar_wait(bar_tempty, (it & 1) ^ 1);
                tc_fence_after();
                for (int pt = px_begin; pt < px_end; ++pt) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    const uint64_t db = umma_smem_desc(sN, 8192, 1024);
                    for (int a = 0; a < u.nacc; ++a) {
                        const uint64_t da = umma_smem_desc(sM + u.acc[a].start_off, u.acc[a].lbo, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + a * NCOLS, da + 128 * kk, db + 128 * kk, idesc, (pt > px_begin) || kk);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x, ++it) {
            const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
            const int nt = (unit / p.ksplit) % p.num_n_tiles;
            const int mt = (unit / (p.ksplit * p.num_n_tiles)) % p.num_m_tiles;
            const WhUnit& u = p.types[unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles)];
            const int m0 = mt * p.m_tile_channels, n0 = nt * BNW;
            mbar_wait(bar_tfull, it & 1);
            tc_fence_after();
            for (int a = 0; a < u.nacc; ++a) {
                const int tap = (row < 64) ? u.acc[a].tapA : u.acc[a].tapB;     // XS == 3: the tap of filter column 0 (kh * 3)
                const int m = m0 + ((row < 64) ? u.acc[a].chA + row : u.acc[a].chB + row - 64);
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * NCOLS;
#pragma unroll 1
                for (int ch = 0; ch < NCOLS / 32; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    if (tap >= 0) {
                        const int col = XS == 1 ? ch * 32 : (ch & 1) * 32;      // channel inside the n-tile
                        const int kw = XS == 1 ? 0 : (ch >> 1);                 // column block = filter column
                        float* dst = p.out + (static_cast<long long>(m) * p.n_total + n0 + col) * 9 + tap + kw;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j * 9, __uint_as_float(r[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// =====================================================================================================
// CTA-pair variant of the weight gradient (cta_group::2, M = 256): the two CTAs of a cluster own two different
// "M units" - (64-channel-pair block mt, filter column kw) combinations - over the SAME pixel slabs and the same
// 128 input channels.  Each CTA stages its own shifted G boxes and only HALF (64 channels) of the In tile.
// Requires Mc % 128 == 0 and BNW == 128.  M unit index = mt * 3 + kw; an odd count is padded with a dummy unit whose
// loads are out of bounds (zeros) and whose epilogue is skipped.
// =====================================================================================================
struct Wh2Params {
    int N, H, W;
    int tiles_w, tiles_h, num_px_tiles;
    int ksplit, px_tiles_per_split;
    int num_m_units, num_m_pairs, num_n_tiles;     // M units = (Mc / 128) * 3
    int ks_slowest;
    float* out;
    int m_total, n_total;
};

struct Wh2Cfg {
    static constexpr int kNBytes = 8192;                        // 64 px x 64 channels (this CTA's half of the N operand)
    static constexpr int kMBytes = 2 * kWhBoxBytes;             // two 64-channel G boxes with h-halo
    static constexpr int kStageBytes = kNBytes + kMBytes;       // 28 KB
    static constexpr int kStages = 6;
    static constexpr int kTmemCols = 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
wgrad3x3_halo2_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI, const Wh2Params p) {
    using Cfg = Wh2Cfg;
    constexpr int STAGES = Cfg::kStages;
    constexpr int BNW = 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t aux = base + STAGES * Cfg::kStageBytes;
    uint8_t* gen_aux = gen_base + STAGES * Cfg::kStageBytes;
    const uint32_t bar_full = aux, bar_empty = aux + 8 * STAGES, bar_tfull = aux + 16 * STAGES, bar_tempty = bar_tfull + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmI);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // unit = (m_pair * num_n_tiles + nt) * ksplit + ks
    const int num_units = p.num_m_pairs * p.num_n_tiles * p.ksplit;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t lead_full = mapa_shared(bar_full, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int unit_ = pair; unit_ < num_units; unit_ += npairs) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const int nt = (unit / p.ksplit) % p.num_n_tiles;
                const int mu = 2 * (unit / (p.ksplit * p.num_n_tiles)) + static_cast<int>(rank);
                const bool live = mu < p.num_m_units;
                const int mt = mu / 3, kw = mu - 3 * mt;
                const int m0 = mt * 128, n0 = nt * BNW + static_cast<int>(rank) * 64;
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                for (int pt = px_begin; pt < px_end; ++pt) {
                    const int wt = pt % p.tiles_w, ht = (pt / p.tiles_w) % p.tiles_h, n = pt / (p.tiles_w * p.tiles_h);
                    const int w0 = wt * 8, h0 = ht * 8;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * Cfg::kStageBytes);
                    const uint32_t fb = lead_full + 8 * stage;
                    tma_load_5d_2cta(sN, &tmI, fb, n0, w0, 0, h0, n);
                    // G shifted by -(kw-1) in w, one-row halo in h; a padding unit reads image index N (all zeros)
                    const int ng = live ? n : p.N;
                    tma_load_5d_2cta(sM, &tmG, fb, m0, w0 - (kw - 1), 0, h0 - 1, ng);
                    tma_load_5d_2cta(sM + kWhBoxBytes, &tmG, fb, m0 + 64, w0 - (kw - 1), 0, h0 - 1, ng);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BNW, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit_ = pair; unit_ < num_units; unit_ += npairs, ++it) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                mbar_wait(bar_tempty, (it & 1) ^ 1);
                tc_fence_after();
                for (int pt = px_begin; pt < px_end; ++pt) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    const uint64_t db = umma_smem_desc(sN, 8192, 1024);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint64_t da = umma_smem_desc(sM + (2 - kh) * 1024, kWhBoxBytes, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_2cta(tmem_base + kh * BNW, da + 128 * kk, db + 128 * kk, idesc, (pt > px_begin) || kk);
                    }
                    umma_commit_2cta(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(bar_tfull);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lead_tempty = mapa_shared(bar_tempty, 0);
        int it = 0;
        for (int unit_ = pair; unit_ < num_units; unit_ += npairs, ++it) {
            const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
            const int nt = (unit / p.ksplit) % p.num_n_tiles;
            const int mu = 2 * (unit / (p.ksplit * p.num_n_tiles)) + static_cast<int>(rank);
            const int mt = mu / 3, kw = mu - 3 * mt;
            const int m = mt * 128 + row, n0 = nt * BNW;
            mbar_wait(bar_tfull, it & 1);
            tc_fence_after();
            if (mu < p.num_m_units) {
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh) {
                    const int tap = kh * 3 + kw;
                    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kh * BNW;
#pragma unroll 1
                    for (int ch = 0; ch < BNW / 32; ++ch) {
                        uint32_t r[32];
                        tmem_ld_32x32(t_addr + ch * 32, r);
                        tmem_ld_wait();
                        float* dst = p.out + (static_cast<long long>(m) * p.n_total + n0 + ch * 32) * 9 + tap;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j * 9, __uint_as_float(r[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace onet
