// First convolution (1 -> 64 or 3 -> 64 channels) in bf16 (Onet_vanilla_20240606.py:111) on the warp-level tensor-core path.
//
// first_layer.cuh recomputes the 64 outputs of a pixel from its 3 x 3 patch with 9 * 64 FMAs and, in the backward pass, adds
// another 9 * 64 FMAs per pixel for A[c][k] = sum_p dz[p][c] v_p[k]: both kernels are bound by instruction issue (~1000
// instructions per pixel), 3-4 x above their HBM time.  Both products are small GEMMs whose operands are EXACT in bf16 (image,
// packed weights and the stored gradient are bf16 in this mode), so a 16-pixel segment of an image row is handled by one warp with
// mma.sync.m16n8k16 (tcgen05 has no shape for K = 9 / N = 10; the kernels stay HBM-bound, the MMAs only take the FMAs away):
//
//   stage 1   Y^T[c][p] = W[c][k] . P^T[k][p]        M = 16 channels (x 4), N = 8 pixels (x 2), K = 16 patch elements (x 1 or 2: 9 / 27 used)
//   stage 2   dz = relu'(bn(y)) g                    on the accumulator fragment, g loaded as 16-byte pieces
//   stage 3   A[c][k] += dz^T[c][p] . P[p][k]        M = 16 channels (x 4), N = 8 (x 2 or 4: patch elements, then a ONES column), K = 16 pixels
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

// Patch element k = tap * CIN + ci (tap = 3 ky + kx), the order of the packed weights.  In the NHWC input tile the 3 * CIN elements
// of one patch row are contiguous: offset(k) = (k / (3 CIN)) * pitch + k % (3 CIN) from the element of the pixel's (ky, kx) = (0, 0).
template <int CIN>
struct Fm {
    static constexpr int K = 9 * CIN;
    static constexpr int KS = (K + 15) / 16;            // k-steps of stage 1
    static constexpr int NT3 = (K + 1 + 7) / 8;         // n-tiles of stage 3: K patch elements + the ones column
    static constexpr int kPitch = (kFmCols + 2) * CIN;  // elements per tile row: 129 / 387 words, odd -> patch rows in different banks
    static constexpr int kTile = (kFmRows + 2) * kPitch;
    __device__ static __forceinline__ int off(int k) { return (k / (3 * CIN)) * kPitch + k % (3 * CIN); }
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Zero-padded input tile rows [h0-1, h0+kFmRows], columns [w0-1, w0+kFmCols] of image n (bf16 bits, NHWC).
template <int CIN>
__device__ __forceinline__ void fm_load_tile(const __nv_bfloat16* __restrict__ in, long long nb, int h0, int w0, int H, int W,
                                             unsigned short* __restrict__ s_x) {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(in);
    for (int rr = threadIdx.x >> 5; rr < kFmRows + 2; rr += static_cast<int>(blockDim.x >> 5)) {
        const int h = h0 - 1 + rr;
        const bool hv = h >= 0 && h < H;
        for (int cc = threadIdx.x & 31; cc < Fm<CIN>::kPitch; cc += 32) {
            const int w = w0 - 1 + cc / CIN;
            s_x[rr * Fm<CIN>::kPitch + cc] =
                (hv && w >= 0 && w < W) ? __ldg(src + (nb + static_cast<long long>(h) * W + w) * CIN + cc % CIN) : static_cast<unsigned short>(0);
        }
    }
}

// Two adjacent patch elements k, k + 1 of the pixel whose (0, 0) patch element is p[0], packed; elements >= K are zero.
template <int CIN>
__device__ __forceinline__ uint32_t fm_pair(const unsigned short* p, int k) {
    const uint32_t lo = k < Fm<CIN>::K ? static_cast<uint32_t>(p[Fm<CIN>::off(k < Fm<CIN>::K ? k : 0)]) : 0u;
    const uint32_t hi = k + 1 < Fm<CIN>::K ? static_cast<uint32_t>(p[Fm<CIN>::off(k + 1 < Fm<CIN>::K ? k + 1 : 0)]) : 0u;
    return lo | (hi << 16);
}

// Weight fragments of stage 1 (A operand, 16 channels x 16 patch elements per (mt, ks)): row r -> channel r*8 + 2mt, row r+8 -> that + 1.
template <int CIN>
__device__ __forceinline__ void fm_weight_frags(const __nv_bfloat16* __restrict__ wp, int r, int q, uint32_t (&wa)[4][Fm<CIN>::KS][4]) {
    constexpr int K = Fm<CIN>::K;
    const unsigned short* w = reinterpret_cast<const unsigned short*>(wp);     // [64][K]
    auto pair = [&](int c, int k) -> uint32_t {
        const uint32_t lo = k < K ? static_cast<uint32_t>(w[c * K + (k < K ? k : 0)]) : 0u;
        const uint32_t hi = k + 1 < K ? static_cast<uint32_t>(w[c * K + (k + 1 < K ? k + 1 : 0)]) : 0u;
        return lo | (hi << 16);
    };
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int c0 = r * 8 + 2 * mt, c1 = c0 + 1;
#pragma unroll
        for (int ks = 0; ks < Fm<CIN>::KS; ++ks) {
            const int k = 16 * ks + 2 * q;
            wa[mt][ks][0] = pair(c0, k);
            wa[mt][ks][1] = pair(c1, k);
            wa[mt][ks][2] = pair(c0, k + 8);
            wa[mt][ks][3] = pair(c1, k + 8);
        }
    }
}

// Stage-1 B fragments of one 16-pixel segment (sb = tile element of pixel 0's (0, 0) patch element): [n-tile][k-step][2]
template <int CIN>
__device__ __forceinline__ void fm_stage1_b(const unsigned short* sb, int r, int q, uint32_t (&b)[2][Fm<CIN>::KS][2]) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const unsigned short* p = sb + (r + 8 * nt) * CIN;
#pragma unroll
        for (int ks = 0; ks < Fm<CIN>::KS; ++ks) {
            b[nt][ks][0] = fm_pair<CIN>(p, 16 * ks + 2 * q);
            b[nt][ks][1] = fm_pair<CIN>(p, 16 * ks + 2 * q + 8);
        }
    }
}

// y of (channel block mt, pixel block nt): [0] = (c0, px 2q), [1] = (c0, px 2q+1), [2] = (c1, px 2q), [3] = (c1, px 2q+1), px += 8 nt
template <int CIN>
__device__ __forceinline__ void fm_stage1(float (&y)[4], const uint32_t (&wa)[Fm<CIN>::KS][4], const uint32_t (&b)[Fm<CIN>::KS][2]) {
    y[0] = y[1] = y[2] = y[3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < Fm<CIN>::KS; ++ks) mma_bf16_16816(y, wa[ks], b[ks][0], b[ks][1]);
}

// Stage-3 / moment B fragments: rows = pixels 2q, 2q+1 (b0) and 2q+8, 2q+9 (b1), column r of n-tile nt = patch element 8 nt + r;
// element K is the ONES column, elements beyond are zero.
template <int CIN>
__device__ __forceinline__ void fm_patch_b(const unsigned short* sb, int r, int q, uint32_t (&pb)[Fm<CIN>::NT3][2]) {
    constexpr int K = Fm<CIN>::K;
#pragma unroll
    for (int nt = 0; nt < Fm<CIN>::NT3; ++nt) {
        const int k = 8 * nt + r;
        if (k < K) {
            const unsigned short* p = sb + 2 * q * CIN + Fm<CIN>::off(k);
            pb[nt][0] = static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[CIN]) << 16);
            pb[nt][1] = static_cast<uint32_t>(p[8 * CIN]) | (static_cast<uint32_t>(p[9 * CIN]) << 16);
        } else {
            pb[nt][0] = pb[nt][1] = k == K ? 0x3f803f80u : 0u;
        }
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

// act = relu(bn(conv(x))), y unrounded (the ROUND_Y = false form of first_conv_fwd_kernel<FIRST_APPLY>)
template <int CIN>
__global__ void __launch_bounds__(256, 2)
first_mma_fwd_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ scale,
                     const float* __restrict__ shift, int group_images, __nv_bfloat16* __restrict__ out) {
    using F = Fm<CIN>;
    __shared__ unsigned short s_x[F::kTile];
    const FmGeom gm = fm_geom(H, W);
    const long long nb = static_cast<long long>(gm.n) * H * W;
    fm_load_tile<CIN>(in, nb, gm.h0, gm.w0, H, W, s_x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int grp = min(gm.n / group_images, 1);
    uint32_t wa[4][F::KS][4];
    fm_weight_frags<CIN>(wp, r, q, wa);
    float sc[8], sh[8];
    load8<float>(scale + grp * 64 + r * 8, sc);
    load8<float>(shift + grp * 64 + r * 8, sh);
    __syncthreads();
    const int tiles = gm.rows * gm.segs;
    for (int t = warp; t < tiles; t += 8) {
        const int row = t / gm.segs, seg = t - row * gm.segs;
        uint32_t b[2][F::KS][2];
        fm_stage1_b<CIN>(s_x + row * F::kPitch + seg * 16 * CIN, r, q, b);
        uint32_t o[2][2][4];                      // [nt][j] -> 8 channels of pixel 2q + j + 8nt
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float y[4];
                fm_stage1<CIN>(y, wa[mt], b[nt]);
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

// One backward pass over g: acc_a[g][c][k] += sum dz v[k]  (k < K),  sums[g][0][c] += sum dz.   sums[g][1][c] is derived from
// A afterwards (first_bwd_assemble_kernel with derive_s2).
template <int CIN>
__global__ void __launch_bounds__(256, CIN == 1 ? 2 : 1)
first_mma_bwd_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wp, const __nv_bfloat16* __restrict__ g, const FirstFusedArgs a) {
    using F = Fm<CIN>;
    constexpr int K = F::K;
    __shared__ unsigned short s_x[F::kTile];
    __shared__ float s_acc[64 * (K + 1)];
    const int H = a.H, W = a.W;
    const FmGeom gm = fm_geom(H, W);
    const long long nb = static_cast<long long>(gm.n) * H * W;
    fm_load_tile<CIN>(in, nb, gm.h0, gm.w0, H, W, s_x);
    for (int i = threadIdx.x; i < 64 * (K + 1); i += 256) s_acc[i] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int grp = min(gm.n / a.group_images, 1);
    uint32_t wa[4][F::KS][4];
    fm_weight_frags<CIN>(wp, r, q, wa);
    float sc[8], sh[8];
    load8<float>(a.scale + grp * 64 + r * 8, sc);
    load8<float>(a.shift + grp * 64 + r * 8, sh);
    float acc[4][F::NT3][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < F::NT3; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    __syncthreads();
    const int tiles = gm.rows * gm.segs;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int t = warp; t < tiles; t += 8) {
        // [nt][j]: 8 channels (r*8..) of pixel 2q + j + 8nt.  Neither prefetching the next segment's g one iteration ahead nor a
        // 4-deep cp.async ring per warp made the kernel faster (0.33 -> 0.36 ms with the ring): it is bound by the rate of the
        // legacy mma.sync path - 16 HMMAs per segment at ~32 clocks each per SM sub-partition account for 80 % of its cycles.
        const int row = t / gm.segs, seg = t - row * gm.segs;
        uint4 gc[2][2];
        {
            const long long p0 = nb + static_cast<long long>(gm.h0 + row) * W + gm.w0 + seg * 16;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int px = 2 * q + j + 8 * nt;
                    gc[nt][j] = (gm.w0 + seg * 16 + px < W) ? __ldg(reinterpret_cast<const uint4*>(g + (p0 + px) * 64 + r * 8)) : zero4;
                }
        }
        const unsigned short* sb = s_x + row * F::kPitch + seg * 16 * CIN;
        uint32_t b[2][F::KS][2];
        fm_stage1_b<CIN>(sb, r, q, b);
        uint32_t pb[F::NT3][2];
        fm_patch_b<CIN>(sb, r, q, pb);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t af[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float y[4];
                fm_stage1<CIN>(y, wa[mt], b[nt]);
                const uint32_t g0 = mt == 0 ? gc[nt][0].x : mt == 1 ? gc[nt][0].y : mt == 2 ? gc[nt][0].z : gc[nt][0].w;   // pixel 2q
                const uint32_t g1 = mt == 0 ? gc[nt][1].x : mt == 1 ? gc[nt][1].y : mt == 2 ? gc[nt][1].z : gc[nt][1].w;   // pixel 2q+1
                const uint32_t m_lo = (relu_open(y[0], sc[2 * mt], sh[2 * mt]) ? 0x0000ffffu : 0u) |
                                      (relu_open(y[1], sc[2 * mt], sh[2 * mt]) ? 0xffff0000u : 0u);
                const uint32_t m_hi = (relu_open(y[2], sc[2 * mt + 1], sh[2 * mt + 1]) ? 0x0000ffffu : 0u) |
                                      (relu_open(y[3], sc[2 * mt + 1], sh[2 * mt + 1]) ? 0xffff0000u : 0u);
                af[2 * nt] = __byte_perm(g0, g1, 0x5410) & m_lo;           // channel c0 of pixels 2q, 2q+1 (+8 nt)
                af[2 * nt + 1] = __byte_perm(g0, g1, 0x7632) & m_hi;       // channel c1
            }
#pragma unroll
            for (int nt = 0; nt < F::NT3; ++nt) mma_bf16_16816(acc[mt][nt], af, pb[nt][0], pb[nt][1]);
        }
    }
    // accumulator fragment of (mt, nt): [0] = (c0, k = 8nt + 2q), [1] = (c0, k + 1), [2] = (c1, k), [3] = (c1, k + 1); k = K: s1
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int c0 = r * 8 + 2 * mt;
#pragma unroll
        for (int nt = 0; nt < F::NT3; ++nt) {
            const int k = 8 * nt + 2 * q;
            if (k <= K) {
                atomicAdd(&s_acc[c0 * (K + 1) + k], acc[mt][nt][0]);
                atomicAdd(&s_acc[(c0 + 1) * (K + 1) + k], acc[mt][nt][2]);
            }
            if (k + 1 <= K) {
                atomicAdd(&s_acc[c0 * (K + 1) + k + 1], acc[mt][nt][1]);
                atomicAdd(&s_acc[(c0 + 1) * (K + 1) + k + 1], acc[mt][nt][3]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (K + 1); i += 256) {
        const int c = i / (K + 1), k = i - c * (K + 1);
        if (k < K) atomicAdd(a.acc_a + (static_cast<long long>(grp) * 64 + c) * K + k, s_acc[i]);
        else atomicAdd(a.sums + static_cast<long long>(grp) * 2 * 64 + c, static_cast<double>(s_acc[i]));
    }
}

// Patch moments of a statistics group on the same fragments: M[k][k'] = sum_p v_p[k] v_p[k'] with a ONES column k' = K, i.e.
// gram = S[K] then G[K][K] (first_layer.cuh).  Used for in_chns = 3 (K = 27: 405 distinct moments do not fit the registers of the
// CUDA-core first_gram_kernel); 2 x 4 MMAs per 16-pixel segment.
template <int CIN>
__global__ void __launch_bounds__(256, 2)
first_gram_mma_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, int group_images, double* __restrict__ gram) {
    using F = Fm<CIN>;
    constexpr int K = F::K, MT = F::KS;
    __shared__ unsigned short s_x[F::kTile];
    __shared__ float s_acc[K * (K + 1)];
    const FmGeom gm = fm_geom(H, W);
    const long long nb = static_cast<long long>(gm.n) * H * W;
    fm_load_tile<CIN>(in, nb, gm.h0, gm.w0, H, W, s_x);
    for (int i = threadIdx.x; i < K * (K + 1); i += 256) s_acc[i] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int grp = min(gm.n / group_images, 1);
    float acc[MT][F::NT3][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < F::NT3; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    __syncthreads();
    const int tiles = gm.rows * gm.segs;
    for (int t = warp; t < tiles; t += 8) {
        const int row = t / gm.segs, seg = t - row * gm.segs;
        const int wleft = W - (gm.w0 + seg * 16);          // pixels of the segment inside the image: the others must not count
        const unsigned short* sb = s_x + row * F::kPitch + seg * 16 * CIN;
        uint32_t pb[F::NT3][2];
        fm_patch_b<CIN>(sb, r, q, pb);
        // A operand: rows = patch elements 16 mt + r (+8), columns = pixels 2q, 2q+1 (+8); pixels beyond the image are zeroed here
        // (the tile holds their real left neighbours in its patch columns)
        const uint32_t m0 = (2 * q < wleft ? 0x0000ffffu : 0u) | (2 * q + 1 < wleft ? 0xffff0000u : 0u);
        const uint32_t m1 = (2 * q + 8 < wleft ? 0x0000ffffu : 0u) | (2 * q + 9 < wleft ? 0xffff0000u : 0u);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t af[4];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k = 16 * mt + r + 8 * hh;
                uint32_t lo = 0u, hi = 0u;
                if (k < K) {
                    const unsigned short* p = sb + 2 * q * CIN + F::off(k);
                    lo = static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[CIN]) << 16);
                    hi = static_cast<uint32_t>(p[8 * CIN]) | (static_cast<uint32_t>(p[9 * CIN]) << 16);
                }
                af[hh] = lo & m0;
                af[2 + hh] = hi & m1;
            }
#pragma unroll
            for (int nt = 0; nt < F::NT3; ++nt) mma_bf16_16816(acc[mt][nt], af, pb[nt][0], pb[nt][1]);
        }
    }
    // [0] = (k = 16mt + r, k' = 8nt + 2q), [1] = (k, k' + 1), [2] = (k + 8, k'), [3] = (k + 8, k' + 1); k' = K: S[k]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < F::NT3; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = 16 * mt + r + 8 * (i >> 1), k2 = 8 * nt + 2 * q + (i & 1);
                if (k < K && k2 <= K) atomicAdd(&s_acc[k * (K + 1) + k2], acc[mt][nt][i]);
            }
    __syncthreads();
    double* dst = gram + static_cast<long long>(grp) * (K + K * K);
    for (int i = threadIdx.x; i < K * (K + 1); i += 256) {
        const int k = i / (K + 1), k2 = i - k * (K + 1);
        if (k2 == K) atomicAdd(dst + k, static_cast<double>(s_acc[i]));
        else atomicAdd(dst + K + k * K + k2, static_cast<double>(s_acc[i]));
    }
}

}  // namespace onet
