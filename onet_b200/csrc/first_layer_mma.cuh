// First convolution of a 1-channel network in bf16 (Onet_vanilla_20240606.py:111, in_chns = 1) on the warp-level tensor-core path.
//
// first_layer.cuh recomputes the 64 outputs of a pixel from its 3 x 3 patch with 9 * 64 FMAs and, in the backward pass, adds
// another 9 * 64 FMAs per pixel for A[c][k] = sum_p dz[p][c] v_p[k]: both kernels are bound by instruction issue (~1000
// instructions per pixel), 3-4 x above their HBM time.  Both products are small GEMMs whose operands are EXACT in bf16 (image,
// packed weights and the stored gradient are bf16 in this mode), so a 16-pixel segment of an image row is handled by one warp with
// mma.sync.m16n8k16 (tcgen05 has no shape for K = 9 / N = 10; the kernels stay HBM-bound, the MMAs only take the FMAs away):
//
//   stage 1   Y^T[c][p] = W[c][k] . P^T[k][p]        M = 16 channels (x 4), N = 8 pixels (x 2), K = 16 taps (9 used)
//   stage 2   dz = relu'(bn(y)) g                    on the accumulator fragment, g loaded as 16-byte pieces
//   stage 3   A[c][k] += dz^T[c][p] . P[p][k]        M = 16 channels (x 4), N = 8 (taps 0..7) + 8 (tap 8, ONES, 6 unused), K = 16 pixels
//
// The accumulator fragment of stage 1 IS the A-operand fragment of stage 3 (row = channel, column = pixel), so dz never leaves
// registers.  The ones column gives s1[c] = sum dz, and s2[c] = sum dz (y - mu) invstd follows in closed form from A because y is
// linear in the patch: sum dz y = w_c . A[c][:]  (first_bwd_assemble_kernel, derive_s2).  The M index of an MMA row is mapped to
// the channel  r * 8 + 2 mt + half  (r = lane / 4), so that a thread owns 8 ADJACENT channels of a pixel: g is loaded and the
// activation is stored as one 16-byte access per pixel and a warp instruction covers 4 whole 128-byte lines.
//
// Forward (first_mma_fwd_kernel) and backward (first_mma_bwd_kernel) evaluate y with the same MMA (same operand roles, same tap
// positions), so the ReLU mask of the backward pass is the forward pass's; a pixel's value does not depend on its position in the
// segment or on the tile (halo-tiled inference relies on that).
#pragma once
#include "first_layer.cuh"

namespace onet {

constexpr int kFmRows = 16;                 // image rows per block
constexpr int kFmCols = 256;                // image columns per block
constexpr int kFmPitch = kFmCols + 2;       // 129 words: odd, so the three patch rows fall into different banks

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Zero-padded input tile rows [h0-1, h0+kFmRows], columns [w0-1, w0+kFmCols] of image n (bf16 bits).
__device__ __forceinline__ void fm_load_tile(const __nv_bfloat16* __restrict__ in, long long nb, int h0, int w0, int H, int W,
                                             unsigned short* __restrict__ s_x) {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(in);
    for (int rr = threadIdx.x >> 5; rr < kFmRows + 2; rr += static_cast<int>(blockDim.x >> 5)) {
        const int h = h0 - 1 + rr;
        const bool hv = h >= 0 && h < H;
        for (int cc = threadIdx.x & 31; cc < kFmPitch; cc += 32) {
            const int w = w0 - 1 + cc;
            s_x[rr * kFmPitch + cc] = (hv && w >= 0 && w < W) ? __ldg(src + nb + static_cast<long long>(h) * W + w) : static_cast<unsigned short>(0);
        }
    }
}

// Weight fragments of stage 1 (A operand, 16 channels x 16 taps per mt): row r -> channel r*8 + 2mt, row r+8 -> that + 1.
__device__ __forceinline__ void fm_weight_frags(const __nv_bfloat16* __restrict__ wp, int r, int q, uint32_t (&wa)[4][4]) {
    const unsigned short* w = reinterpret_cast<const unsigned short*>(wp);     // [64][9]
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int c0 = r * 8 + 2 * mt, c1 = c0 + 1;
        const int k = 2 * q;                                                    // taps 2q, 2q+1 (all < 9), and 8 for q == 0
        wa[mt][0] = static_cast<uint32_t>(w[c0 * 9 + k]) | (static_cast<uint32_t>(w[c0 * 9 + k + 1]) << 16);
        wa[mt][1] = static_cast<uint32_t>(w[c1 * 9 + k]) | (static_cast<uint32_t>(w[c1 * 9 + k + 1]) << 16);
        wa[mt][2] = q == 0 ? static_cast<uint32_t>(w[c0 * 9 + 8]) : 0u;
        wa[mt][3] = q == 0 ? static_cast<uint32_t>(w[c1 * 9 + 8]) : 0u;
    }
}

// Stage 1 for one 16-pixel segment whose first patch element is sb[0] (= tile row of the pixel's row - 1, column of pixel 0 - 1):
// y[mt][nt] = accumulator fragments: [0] = (c0, px 2q), [1] = (c0, px 2q+1), [2] = (c1, px 2q), [3] = (c1, px 2q+1), px += 8 nt.
struct FmTaps { int o0, o1; };          // offsets of taps 2q and 2q+1 inside the tile
__device__ __forceinline__ FmTaps fm_taps(int q) {
    FmTaps t;
    t.o0 = ((2 * q) / 3) * kFmPitch + (2 * q) % 3;
    t.o1 = ((2 * q + 1) / 3) * kFmPitch + (2 * q + 1) % 3;
    return t;
}

__device__ __forceinline__ void fm_stage1_b(const unsigned short* sb, int r, int q, FmTaps t, uint32_t (&b)[2][2]) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const unsigned short* p = sb + r + 8 * nt;
        b[nt][0] = static_cast<uint32_t>(p[t.o0]) | (static_cast<uint32_t>(p[t.o1]) << 16);
        b[nt][1] = q == 0 ? static_cast<uint32_t>(p[2 * kFmPitch + 2]) : 0u;
    }
}

struct FmGeom {
    int n, h0, w0, rows, segs;
};
__device__ __forceinline__ FmGeom fm_geom(int H, int W) {
    const int cb = (W + kFmCols - 1) / kFmCols, rb = (H + kFmRows - 1) / kFmRows;
    int b = blockIdx.x;
    FmGeom g;
    g.w0 = (b % cb) * kFmCols; b /= cb;
    g.h0 = (b % rb) * kFmRows;
    g.n = b / rb;
    g.rows = min(kFmRows, H - g.h0);
    g.segs = (min(kFmCols, W - g.w0) + 15) >> 4;
    return g;
}

// act = relu(bn(conv(x))) for in_chns = 1, bf16, y unrounded (the ROUND_Y = false form of first_conv_fwd_kernel<FIRST_APPLY>)
__global__ void __launch_bounds__(256, 2)
first_mma_fwd_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ scale,
                     const float* __restrict__ shift, int group_images, __nv_bfloat16* __restrict__ out) {
    __shared__ unsigned short s_x[(kFmRows + 2) * kFmPitch];
    const FmGeom gm = fm_geom(H, W);
    const long long nb = static_cast<long long>(gm.n) * H * W;
    fm_load_tile(in, nb, gm.h0, gm.w0, H, W, s_x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int grp = min(gm.n / group_images, 1);
    uint32_t wa[4][4];
    fm_weight_frags(wp, r, q, wa);
    float sc[8], sh[8];
    load8<float>(scale + grp * 64 + r * 8, sc);
    load8<float>(shift + grp * 64 + r * 8, sh);
    const FmTaps taps = fm_taps(q);
    __syncthreads();
    const int tiles = gm.rows * gm.segs;
    for (int t = warp; t < tiles; t += 8) {
        const int row = t / gm.segs, seg = t - row * gm.segs;
        uint32_t b[2][2];
        fm_stage1_b(s_x + row * kFmPitch + seg * 16, r, q, taps, b);
        uint32_t o[2][2][4];                      // [nt][j] -> 8 channels of pixel 2q + j + 8nt
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float y[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16_16816(y, wa[mt], b[nt][0], b[nt][1]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float v0 = fmaxf(fmaf(y[j], sc[2 * mt], sh[2 * mt]), 0.f);
                    const float v1 = fmaxf(fmaf(y[j + 2], sc[2 * mt + 1], sh[2 * mt + 1]), 0.f);
                    const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
                    o[nt][j][mt] = *reinterpret_cast<const uint32_t*>(&pk);
                }
            }
        const long long p0 = nb + static_cast<long long>(gm.h0 + row) * W + gm.w0 + seg * 16;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int px = 2 * q + j + 8 * nt;
                if (gm.w0 + seg * 16 + px < W)
                    *reinterpret_cast<uint4*>(out + (p0 + px) * 64 + r * 8) = make_uint4(o[nt][j][0], o[nt][j][1], o[nt][j][2], o[nt][j][3]);
            }
    }
}

// One backward pass over g: acc_a[g][c][k] += sum dz v[k]  (k = 0..8),  sums[g][0][c] += sum dz.   sums[g][1][c] is derived from
// A afterwards (first_bwd_assemble_kernel with derive_s2).
__global__ void __launch_bounds__(256, 2)
first_mma_bwd_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wp, const __nv_bfloat16* __restrict__ g, const FirstFusedArgs a) {
    __shared__ unsigned short s_x[(kFmRows + 2) * kFmPitch];
    __shared__ float s_acc[64 * 10];
    const int H = a.H, W = a.W;
    const FmGeom gm = fm_geom(H, W);
    const long long nb = static_cast<long long>(gm.n) * H * W;
    fm_load_tile(in, nb, gm.h0, gm.w0, H, W, s_x);
    for (int i = threadIdx.x; i < 64 * 10; i += 256) s_acc[i] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int grp = min(gm.n / a.group_images, 1);
    uint32_t wa[4][4];
    fm_weight_frags(wp, r, q, wa);
    float sc[8], sh[8];
    load8<float>(a.scale + grp * 64 + r * 8, sc);
    load8<float>(a.shift + grp * 64 + r * 8, sh);
    const FmTaps taps = fm_taps(q);
    // stage 3, B operand rows = pixels 2q, 2q+1 (+8), column = tap r: offset of tap r; second n-tile: r == 0 -> tap 8, r == 1 -> ones
    const int o3 = (r / 3) * kFmPitch + r % 3;
    const uint32_t ones = r == 1 ? 0x3f803f80u : 0u;
    float acc[4][2][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    __syncthreads();
    const int tiles = gm.rows * gm.segs;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int t = warp; t < tiles; t += 8) {
        // [nt][j]: 8 channels (r*8..) of pixel 2q + j + 8nt.  Neither prefetching the next segment's g one iteration ahead nor a
        // 4-deep cp.async ring per warp made the kernel faster (0.33 -> 0.36 ms with the ring): it is bound by the rate of the
        // legacy mma.sync path - 16 HMMAs per segment at ~32 clocks each per SM sub-partition account for 80 % of its cycles.
        uint4 gc[2][2];
        {
            const int row = t / gm.segs, seg = t - row * gm.segs;
            const long long p0 = nb + static_cast<long long>(gm.h0 + row) * W + gm.w0 + seg * 16;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int px = 2 * q + j + 8 * nt;
                    gc[nt][j] = (gm.w0 + seg * 16 + px < W) ? __ldg(reinterpret_cast<const uint4*>(g + (p0 + px) * 64 + r * 8)) : zero4;
                }
        }
        const int row = t / gm.segs, seg = t - row * gm.segs;
        const unsigned short* sb = s_x + row * kFmPitch + seg * 16;
        uint32_t b[2][2];
        fm_stage1_b(sb, r, q, taps, b);
        // stage-3 B fragments: P[px][k]
        uint32_t p3[2][2];
        {
            const unsigned short* p = sb + 2 * q + o3;
            p3[0][0] = static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 16);
            p3[0][1] = static_cast<uint32_t>(p[8]) | (static_cast<uint32_t>(p[9]) << 16);
            const unsigned short* p8 = sb + 2 * q + 2 * kFmPitch + 2;
            p3[1][0] = r == 0 ? (static_cast<uint32_t>(p8[0]) | (static_cast<uint32_t>(p8[1]) << 16)) : ones;
            p3[1][1] = r == 0 ? (static_cast<uint32_t>(p8[8]) | (static_cast<uint32_t>(p8[9]) << 16)) : ones;
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t af[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float y[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16_16816(y, wa[mt], b[nt][0], b[nt][1]);
                const uint32_t g0 = mt == 0 ? gc[nt][0].x : mt == 1 ? gc[nt][0].y : mt == 2 ? gc[nt][0].z : gc[nt][0].w;   // pixel 2q
                const uint32_t g1 = mt == 0 ? gc[nt][1].x : mt == 1 ? gc[nt][1].y : mt == 2 ? gc[nt][1].z : gc[nt][1].w;   // pixel 2q+1
                const uint32_t m_lo = (relu_open(y[0], sc[2 * mt], sh[2 * mt]) ? 0x0000ffffu : 0u) |
                                      (relu_open(y[1], sc[2 * mt], sh[2 * mt]) ? 0xffff0000u : 0u);
                const uint32_t m_hi = (relu_open(y[2], sc[2 * mt + 1], sh[2 * mt + 1]) ? 0x0000ffffu : 0u) |
                                      (relu_open(y[3], sc[2 * mt + 1], sh[2 * mt + 1]) ? 0xffff0000u : 0u);
                af[2 * nt] = __byte_perm(g0, g1, 0x5410) & m_lo;           // channel c0 of pixels 2q, 2q+1 (+8 nt)
                af[2 * nt + 1] = __byte_perm(g0, g1, 0x7632) & m_hi;       // channel c1
            }
            mma_bf16_16816(acc[mt][0], af, p3[0][0], p3[0][1]);
            mma_bf16_16816(acc[mt][1], af, p3[1][0], p3[1][1]);
        }
    }
    // accumulator fragment: [0] = (c0, k = 2q), [1] = (c0, 2q+1), [2] = (c1, 2q), [3] = (c1, 2q+1); second n-tile: k = 8 + 2q (+1),
    // of which k = 8 (tap 8) and k = 9 (s1) are used
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int c0 = r * 8 + 2 * mt;
        atomicAdd(&s_acc[c0 * 10 + 2 * q], acc[mt][0][0]);
        atomicAdd(&s_acc[c0 * 10 + 2 * q + 1], acc[mt][0][1]);
        atomicAdd(&s_acc[(c0 + 1) * 10 + 2 * q], acc[mt][0][2]);
        atomicAdd(&s_acc[(c0 + 1) * 10 + 2 * q + 1], acc[mt][0][3]);
        if (q == 0) {
            atomicAdd(&s_acc[c0 * 10 + 8], acc[mt][1][0]);
            atomicAdd(&s_acc[c0 * 10 + 9], acc[mt][1][1]);
            atomicAdd(&s_acc[(c0 + 1) * 10 + 8], acc[mt][1][2]);
            atomicAdd(&s_acc[(c0 + 1) * 10 + 9], acc[mt][1][3]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 10; i += 256) {
        const int c = i / 10, k = i - c * 10;
        if (k < 9) atomicAdd(a.acc_a + (static_cast<long long>(grp) * 64 + c) * 9 + k, s_acc[i]);
        else atomicAdd(a.sums + static_cast<long long>(grp) * 2 * 64 + c, static_cast<double>(s_acc[i]));
    }
}

}  // namespace onet
