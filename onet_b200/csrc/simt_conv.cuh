// CUDA-core (FP32 FMA) implicit-GEMM kernels, templated on the storage type T (float = FP32 verification
// mode, __nv_bfloat16 = throughput mode).  They serve three purposes: the FP32 verification mode of the whole
// path, the first convolution (K = 9*in_chns = 9 or 27 is not a tensor-core shape), and the ground truth the
// tcgen05 kernels are unit-tested against.  Activations NHWC, weights in PyTorch layouts.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace onet {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// 3x3 / pad 1 convolution, forward (also used for dgrad with flipped+transposed packed weights).
//   in  : [N,H,W,ldi] (+ci_off), Cin channels used
//   wp  : packed [Cout][9][Cin]  (K index = tap*Cin + ci), type T
//   out : [N,H,W,ldo] (+co_off), raw conv output, type T
//   stat_sum / stat_sq : double [groups][Cout] partial sums of the STORED values, or nullptr
// Tile: 64 pixels x 64 couts per 256-thread block, 4x4 outputs per thread, K chunks of 16.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const T* __restrict__ in, long long ldi, int ci_off, int N, int H, int W, int Cin,
                    const T* __restrict__ wp, int Cout, T* __restrict__ out, long long ldo, int co_off,
                    double* __restrict__ stat_sum, double* __restrict__ stat_sq, int group_images) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    // [sum|sq][group][pixel-row thread ty][channel]: per-thread partial sums, folded over ty in a FIXED order.  (Shared-memory
    // float atomics made the statistics - and through ReLU / max-pool switching the whole gradient of an ill-conditioned small
    // case - depend on the order in which the threads arrived: FP32 verification mode must be reproducible.)
    __shared__ float s_stat[2][2][16][64];
    const int tid = threadIdx.x;
    const long long M = static_cast<long long>(N) * H * W;
    const long long m0 = static_cast<long long>(blockIdx.x) * 64;
    const int co0 = blockIdx.y * 64;
    const int K = 9 * Cin;

    // A-loader role: thread loads k = tid%16, pixels tid/16 + 16*e
    const int a_kk = tid & 15;
    int a_n[4], a_h[4], a_w[4];
    bool a_ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long m = m0 + (tid >> 4) + 16 * e;
        a_ok[e] = m < M;
        const long long mm = a_ok[e] ? m : 0;
        a_w[e] = static_cast<int>(mm % W);
        a_h[e] = static_cast<int>((mm / W) % H);
        a_n[e] = static_cast<int>(mm / (static_cast<long long>(W) * H));
    }
    // B-loader role: thread loads k = tid%16, couts tid/16 + 16*e
    const int ty = tid >> 4, tx = tid & 15;   // compute role: pixels ty*4.., couts tx*4..
    float acc[4][4] = {};

    for (int k0 = 0; k0 < K; k0 += 16) {
        const int k = k0 + a_kk;
        int tap = k / Cin, ci = k - tap * Cin;
        const int kh = tap / 3 - 1, kw = tap % 3 - 1;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = 0.f;
            if (k < K && a_ok[e]) {
                const int hh = a_h[e] + kh, ww = a_w[e] + kw;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    v = to_f<T>(in[((static_cast<long long>(a_n[e]) * H + hh) * W + ww) * ldi + ci_off + ci]);
            }
            As[a_kk][(tid >> 4) + 16 * e] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int co = co0 + (tid >> 4) + 16 * e;
            float v = 0.f;
            if (k < K && co < Cout) v = to_f<T>(wp[static_cast<long long>(co) * K + k]);
            Bs[a_kk][(tid >> 4) + 16 * e] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool do_stats = stat_sum != nullptr;
    float ps[2][4] = {}, pq[2][4] = {};      // this thread's sums / sums of squares per statistics group and channel j
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int n = static_cast<int>(m / (static_cast<long long>(W) * H));
        const int grp = do_stats ? min(n / group_images, 1) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co >= Cout) continue;
            const T sv = from_f<T>(acc[i][j]);
            out[m * ldo + co_off + co] = sv;
            if (do_stats) {
                const float f = to_f<T>(sv);
                if (grp == 0) { ps[0][j] += f; pq[0][j] = fmaf(f, f, pq[0][j]); }
                else { ps[1][j] += f; pq[1][j] = fmaf(f, f, pq[1][j]); }
            }
        }
    }
    if (do_stats) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_stat[0][g][ty][tx * 4 + j] = ps[g][j];
                s_stat[1][g][ty][tx * 4 + j] = pq[g][j];
            }
        __syncthreads();
        for (int i = tid; i < 2 * 2 * 64; i += 256) {
            const int which = i >> 7, g = (i >> 6) & 1, c = i & 63;
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) t += s_stat[which][g][r][c];
            if (co0 + c < Cout && t != 0.f)
                atomicAdd((which == 0 ? stat_sum : stat_sq) + static_cast<long long>(g) * Cout + co0 + c, static_cast<double>(t));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// First convolution of the U-Net (in_chns = CIN <= 4, Cout = 64, W % 4 == 0): K = 9*CIN is not a tensor-core
// shape and the layer is purely bandwidth-bound (writes 128 B per pixel, reads 2*CIN B), so it gets a direct
// kernel: one thread = 4 horizontally adjacent pixels x 8 output channels (the 3 x 6 input patch is shared by the
// four pixels), weights in shared memory, BatchNorm partial sums fused.
// ------------------------------------------------------------------------------------------------
template <typename T, int CIN, int PX>
__device__ __forceinline__ void load_patch(const T* __restrict__ in, long long nb, int h, int w0, int H, int W,
                                           float (&x)[3][PX + 2][CIN]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
#pragma unroll
        for (int cidx = 0; cidx < PX + 2; ++cidx) {
            const int ww = w0 + cidx - 1;
            const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < CIN; ++c)
                x[r][cidx][c] = ok ? to_f<T>(in[(nb + static_cast<long long>(hh) * W + ww) * CIN + c]) : 0.f;
        }
    }
}

// One thread = PX horizontally adjacent pixels x 8 output channels; the 3 x (PX+2) input patch is shared by the PX
// pixels.  W % PX == 0.
template <typename T, int CIN, int PX>
__global__ void __launch_bounds__(256, PX == 4 ? 2 : 3)
conv_first_fwd_kernel(const T* __restrict__ in, int N, int H, int W, const T* __restrict__ wp, T* __restrict__ out,
                      double* __restrict__ stat_sum, double* __restrict__ stat_sq, int group_images) {
    constexpr int K = 9 * CIN;
    __shared__ float ws[K][64];
    __shared__ float s_red[16][256];
    for (int i = threadIdx.x; i < K * 64; i += 256) ws[i % K][i / K] = to_f<T>(wp[i]);   // wp is [64][K]
    __syncthreads();
    const int oc = threadIdx.x & 7, ln = threadIdx.x >> 3;
    const int g = blockIdx.y;
    const int WG = W / PX;
    const long long QW = static_cast<long long>(H) * WG;      // pixel groups per image
    const long long q_begin = static_cast<long long>(g) * group_images * QW;
    const long long q_end = static_cast<long long>(g == static_cast<int>(gridDim.y) - 1 ? N : (g + 1) * group_images) * QW;
    float a1[8] = {}, a2[8] = {};
    for (long long q = q_begin + blockIdx.x * 32LL + ln; q < q_end; q += gridDim.x * 32LL) {
        const int w0 = static_cast<int>(q % WG) * PX, h = static_cast<int>((q / WG) % H);
        const long long nb = (q / QW) * H * W;
        float x[3][PX + 2][CIN];
        load_patch<T, CIN, PX>(in, nb, h, w0, H, W, x);
        asm volatile("" ::: "memory");     // keep the weight reads inside the loop (hoisting 9*CIN*8 of them spills)
        float acc[PX][8] = {};
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                float wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = ws[t * CIN + c][oc * 8 + i];
#pragma unroll
                for (int j = 0; j < PX; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(x[t / 3][j + t % 3][c], wv[i], acc[j][i]);
            }
        const long long p0 = nb + static_cast<long long>(h) * W + w0;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            T o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o[i] = from_f<T>(acc[j][i]);
                const float f = to_f<T>(o[i]);
                a1[i] += f;
                a2[i] = fmaf(f, f, a2[i]);
            }
            if (sizeof(T) == 2) {
                *reinterpret_cast<uint4*>(out + (p0 + j) * 64 + oc * 8) = *reinterpret_cast<const uint4*>(o);
            } else {
                *reinterpret_cast<float4*>(out + (p0 + j) * 64 + oc * 8) = *reinterpret_cast<const float4*>(o);
                *reinterpret_cast<float4*>(out + (p0 + j) * 64 + oc * 8 + 4) = *reinterpret_cast<const float4*>(o + 4);
            }
        }
    }
    if (stat_sum != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s_red[i][threadIdx.x] = a1[i];
            s_red[8 + i][threadIdx.x] = a2[i];
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int which = threadIdx.x >> 3, o8 = threadIdx.x & 7;   // which in 0..15, octet o8
            float sacc = 0.f;
            for (int l = 0; l < 32; ++l) sacc += s_red[which][l * 8 + o8];
            const int c = o8 * 8 + (which & 7);
            double* dst = (which < 8 ? stat_sum : stat_sq) + static_cast<long long>(g) * 64 + c;
            atomicAdd(dst, static_cast<double>(sacc));
        }
    }
}

// dW[co][ci][tap] += sum_px G[px][co] * In[px + tap][ci] for the first convolution (CIN <= 4, Cout = 64, W % 4 == 0).
// One thread = CPT output channels x all 9*CIN taps, 4 adjacent pixels per step; warp-shuffle + atomics at the end.
template <typename T, int CIN, int CPT>
__global__ void __launch_bounds__(256, 2)
conv_first_wgrad_kernel(const T* __restrict__ g, const T* __restrict__ in, int N, int H, int W, float* __restrict__ dw) {
    constexpr int K = 9 * CIN;
    constexpr int NG = 64 / CPT;            // channel groups
    constexpr int LANES = 256 / NG;         // pixel-quad lanes per block
    const int cg = threadIdx.x % NG, ln = threadIdx.x / NG;
    const int W4 = W >> 2;
    const long long QW = static_cast<long long>(H) * W4;
    const long long Q = static_cast<long long>(N) * QW;
    float acc[CPT][K];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[i][k] = 0.f;
    for (long long q = blockIdx.x * static_cast<long long>(LANES) + ln; q < Q; q += static_cast<long long>(gridDim.x) * LANES) {
        const int w0 = static_cast<int>(q % W4) * 4, h = static_cast<int>((q / W4) % H);
        const long long nb = (q / QW) * H * W;
        const long long p0 = nb + static_cast<long long>(h) * W + w0;
        float gv[4][CPT];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (CPT == 8 && sizeof(T) == 2) {
                const uint4 u = *reinterpret_cast<const uint4*>(g + (p0 + j) * 64 + cg * 8);
                const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    gv[j][(2 * i) % CPT] = __uint_as_float(wv[i] << 16);
                    gv[j][(2 * i + 1) % CPT] = __uint_as_float(wv[i] & 0xffff0000u);
                }
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i) gv[j][i] = to_f<T>(g[(p0 + j) * 64 + cg * CPT + i]);
            }
        }
        float x[3][6][CIN];
        load_patch<T, CIN, 4>(in, nb, h, w0, H, W, x);
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < CIN; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        acc[i][t * CIN + c] = fmaf(gv[j][i], x[t / 3][j + t % 3][c], acc[i][t * CIN + c]);
    }
    // lanes of the same channel group inside a warp sit NG threads apart
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float v = acc[i][k];
            for (int off = NG; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if ((threadIdx.x & 31) < NG) {
                const int co = cg * CPT + i, tap = k / CIN, ci = k % CIN;
                atomicAdd(dw + (static_cast<long long>(co) * CIN + ci) * 9 + tap, v);
            }
        }
}

// ------------------------------------------------------------------------------------------------
// Row-sliding variants of the two first-layer kernels.  A thread owns 4 adjacent columns x CPT output channels and
// walks down ROWS image rows keeping the 3 x 6 input patch in registers: every new row costs one row of input
// (one 8-byte + two 2-byte loads for bf16, in_chns = 1) instead of the whole patch, and no index arithmetic.
// Block = 256 threads = (64 / CPT) channel groups x (256 * CPT / 64) column groups of ONE image row range, so a
// block never mixes BatchNorm statistics groups.  grid.x = N * ceil(H / ROWS) * ceil((W/4) / column groups).
// ------------------------------------------------------------------------------------------------
template <typename T, int CIN>
__device__ __forceinline__ void load_row6(const T* __restrict__ in, long long nb, int hh, int w0, int H, int W,
                                          float (&x)[6][CIN]) {
    const bool row_ok = hh >= 0 && hh < H;
    if (sizeof(T) == 2 && CIN == 1) {
        const unsigned short* base = reinterpret_cast<const unsigned short*>(in) + nb + static_cast<long long>(row_ok ? hh : 0) * W + w0;
        uint2 mid = make_uint2(0u, 0u);
        unsigned short lft = 0, rgt = 0;
        if (row_ok) {
            mid = __ldg(reinterpret_cast<const uint2*>(base));
            if (w0 > 0) lft = __ldg(base - 1);
            if (w0 + 4 < W) rgt = __ldg(base + 4);
        }
        x[0][0] = __uint_as_float(static_cast<uint32_t>(lft) << 16);
        x[1][0] = __uint_as_float(mid.x << 16);
        x[2][0] = __uint_as_float(mid.x & 0xffff0000u);
        x[3][0] = __uint_as_float(mid.y << 16);
        x[4][0] = __uint_as_float(mid.y & 0xffff0000u);
        x[5][0] = __uint_as_float(static_cast<uint32_t>(rgt) << 16);
    } else {
#pragma unroll
        for (int cidx = 0; cidx < 6; ++cidx) {
            const int ww = w0 + cidx - 1;
            const bool ok = row_ok && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < CIN; ++c)
                x[cidx][c] = ok ? to_f<T>(in[(nb + static_cast<long long>(hh) * W + ww) * CIN + c]) : 0.f;
        }
    }
}

template <typename T, int CIN, int ROWS>
__global__ void __launch_bounds__(256, CIN == 1 ? 3 : 2)
conv_first_fwd_rows_kernel(const T* __restrict__ in, int N, int H, int W, const T* __restrict__ wp, T* __restrict__ out,
                           double* __restrict__ stat_sum, double* __restrict__ stat_sq, int group_images) {
    constexpr int K = 9 * CIN;
    __shared__ float ws[K][64];
    __shared__ float s_red[16][256];
    for (int i = threadIdx.x; i < K * 64; i += 256) ws[i % K][i / K] = to_f<T>(wp[i]);   // wp is [64][K]
    __syncthreads();
    const int oc = threadIdx.x & 7, ln = threadIdx.x >> 3;
    const int WG = W >> 2, wgb = (WG + 31) >> 5, chunks = (H + ROWS - 1) / ROWS;
    int b = blockIdx.x;
    const int wg = (b % wgb) * 32 + ln; b /= wgb;
    const int h0 = (b % chunks) * ROWS;
    const int n = b / chunks;
    const int h1 = min(H, h0 + ROWS);
    const int w0 = wg * 4;
    const long long nb = static_cast<long long>(n) * H * W;
    float a1[8] = {}, a2[8] = {};
    if (wg < WG) {
        float x[3][6][CIN];
        load_row6<T, CIN>(in, nb, h0 - 1, w0, H, W, x[0]);
        load_row6<T, CIN>(in, nb, h0, w0, H, W, x[1]);
        for (int h = h0; h < h1; ++h) {
            load_row6<T, CIN>(in, nb, h + 1, w0, H, W, x[2]);
            asm volatile("" ::: "memory");     // keep the weight reads inside the loop (hoisting all of them spills)
            float acc[4][8] = {};
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    float wv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) wv[i] = ws[t * CIN + c][oc * 8 + i];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(x[t / 3][j + t % 3][c], wv[i], acc[j][i]);
                }
            const long long p0 = nb + static_cast<long long>(h) * W + w0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                T o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o[i] = from_f<T>(acc[j][i]);
                    const float f = to_f<T>(o[i]);
                    a1[i] += f;
                    a2[i] = fmaf(f, f, a2[i]);
                }
                if (sizeof(T) == 2) {
                    *reinterpret_cast<uint4*>(out + (p0 + j) * 64 + oc * 8) = *reinterpret_cast<const uint4*>(o);
                } else {
                    *reinterpret_cast<float4*>(out + (p0 + j) * 64 + oc * 8) = *reinterpret_cast<const float4*>(o);
                    *reinterpret_cast<float4*>(out + (p0 + j) * 64 + oc * 8 + 4) = *reinterpret_cast<const float4*>(o + 4);
                }
            }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx)
#pragma unroll
                for (int c = 0; c < CIN; ++c) { x[0][cidx][c] = x[1][cidx][c]; x[1][cidx][c] = x[2][cidx][c]; }
        }
    }
    if (stat_sum != nullptr) {
        const int g = min(n / group_images, 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s_red[i][threadIdx.x] = a1[i];
            s_red[8 + i][threadIdx.x] = a2[i];
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int which = threadIdx.x >> 3, o8 = threadIdx.x & 7;   // which in 0..15, octet o8
            float sacc = 0.f;
            for (int l = 0; l < 32; ++l) sacc += s_red[which][l * 8 + o8];
            const int c = o8 * 8 + (which & 7);
            double* dst = (which < 8 ? stat_sum : stat_sq) + static_cast<long long>(g) * 64 + c;
            atomicAdd(dst, static_cast<double>(sacc));
        }
    }
}

template <typename T, int CIN, int CPT, int ROWS>
__global__ void __launch_bounds__(256, 2)
conv_first_wgrad_rows_kernel(const T* __restrict__ g, const T* __restrict__ in, int N, int H, int W, float* __restrict__ dw,
                             float* __restrict__ partial) {
    constexpr int K = 9 * CIN;
    constexpr int NG = 64 / CPT;            // channel groups
    constexpr int LANES = 256 / NG;         // column groups per block
    __shared__ float s_acc[64 * K];
    for (int i = threadIdx.x; i < 64 * K; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    const int cg = threadIdx.x % NG, ln = threadIdx.x / NG;
    const int WG = W >> 2, wgb = (WG + LANES - 1) / LANES, chunks = (H + ROWS - 1) / ROWS;
    int b = blockIdx.x;
    const int wg = (b % wgb) * LANES + ln; b /= wgb;
    const int h0 = (b % chunks) * ROWS;
    const int n = b / chunks;
    const int h1 = min(H, h0 + ROWS);
    const int w0 = wg * 4;
    const long long nb = static_cast<long long>(n) * H * W;
    float acc[CPT][K];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[i][k] = 0.f;
    if (wg < WG) {
        float x[3][6][CIN];
        load_row6<T, CIN>(in, nb, h0 - 1, w0, H, W, x[0]);
        load_row6<T, CIN>(in, nb, h0, w0, H, W, x[1]);
        for (int h = h0; h < h1; ++h) {
            load_row6<T, CIN>(in, nb, h + 1, w0, H, W, x[2]);
            const long long p0 = nb + static_cast<long long>(h) * W + w0;
            float gv[4][CPT];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CPT == 8 && sizeof(T) == 2) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(g + (p0 + j) * 64 + cg * 8));
                    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        gv[j][(2 * i) % CPT] = __uint_as_float(wv[i] << 16);
                        gv[j][(2 * i + 1) % CPT] = __uint_as_float(wv[i] & 0xffff0000u);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < CPT; ++i) gv[j][i] = to_f<T>(g[(p0 + j) * 64 + cg * CPT + i]);
                }
            }
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int c = 0; c < CIN; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < CPT; ++i)
                            acc[i][t * CIN + c] = fmaf(gv[j][i], x[t / 3][j + t % 3][c], acc[i][t * CIN + c]);
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx)
#pragma unroll
                for (int c = 0; c < CIN; ++c) { x[0][cidx][c] = x[1][cidx][c]; x[1][cidx][c] = x[2][cidx][c]; }
        }
    }
    // lanes of the same channel group inside a warp sit NG threads apart; the 8 warps then add into shared memory one
    // after the other (a FIXED order: the result does not depend on scheduling, unlike shared-memory atomics)
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k)
            for (int off = NG; off < 32; off <<= 1) acc[i][k] += __shfl_xor_sync(0xffffffffu, acc[i][k], off);
    for (int wp = 0; wp < 8; ++wp) {
        if ((threadIdx.x >> 5) == wp && (threadIdx.x & 31) < NG) {
#pragma unroll
            for (int i = 0; i < CPT; ++i)
#pragma unroll
                for (int k = 0; k < K; ++k) s_acc[(cg * CPT + i) * K + k] += acc[i][k];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 64 * K; i += 256) {
        const int co = i / K, k = i % K, tap = k / CIN, ci = k % CIN;
        const long long idx = (static_cast<long long>(co) * CIN + ci) * 9 + tap;
        // deterministic mode: per-block partials, summed in block order by splitk_reduce_kernel
        if (partial != nullptr) partial[static_cast<long long>(blockIdx.x) * (64 * K) + idx] = s_acc[i];
        else atomicAdd(dw + idx, s_acc[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolution weight gradient:  dW[co][ci][kh][kw] += sum_pixels G[p][co] * In[p + (kh-1,kw-1)][ci]
//   g : [N,H,W,ldg] (+co_off) ; in : [N,H,W,ldi] (+ci_off) ; dw : fp32 PyTorch OIHW, atomically accumulated.
// Tile: 64 couts x 64 k (k = tap*Cin + ci) per block, pixel range split over blockIdx.z.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_simt_kernel(const T* __restrict__ g, long long ldg, int co_off, const T* __restrict__ in, long long ldi,
                          int ci_off, int N, int H, int W, int Cin, int Cout, float* __restrict__ dw,
                          long long px_per_split, float* __restrict__ partial) {
    __shared__ float Gs[16][64 + 4];
    __shared__ float Is[16][64 + 4];
    const int tid = threadIdx.x;
    const long long M = static_cast<long long>(N) * H * W;
    const int co0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int K = 9 * Cin;
    const long long p_begin = static_cast<long long>(blockIdx.z) * px_per_split;
    const long long p_end = min(p_begin + px_per_split, M);
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};

    // I-loader role: thread loads pixel pp = tid/16 of the chunk, k = k0 + tid%16 + 16*e
    int l_tap[4], l_ci[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = k0 + (tid & 15) + 16 * e;
        l_tap[e] = (k < K) ? k / Cin : -1;
        l_ci[e] = (k < K) ? k % Cin : 0;
    }
    for (long long pc = p_begin; pc < p_end; pc += 16) {
        const long long pm = pc + (tid >> 4);
        const bool ok = pm < p_end;
        const long long mm = ok ? pm : 0;
        const int w = static_cast<int>(mm % W), h = static_cast<int>((mm / W) % H),
                  n = static_cast<int>(mm / (static_cast<long long>(W) * H));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int co = co0 + (tid & 15) + 16 * e;
            Gs[tid >> 4][(tid & 15) + 16 * e] = (ok && co < Cout) ? to_f<T>(g[mm * ldg + co_off + co]) : 0.f;
            float v = 0.f;
            if (ok && l_tap[e] >= 0) {
                const int hh = h + l_tap[e] / 3 - 1, ww = w + l_tap[e] % 3 - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    v = to_f<T>(in[((static_cast<long long>(n) * H + hh) * W + ww) * ldi + ci_off + l_ci[e]]);
            }
            Is[tid >> 4][(tid & 15) + 16 * e] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Gs[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Is[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            const int tap = k / Cin, ci = k % Cin;
            const long long idx = (static_cast<long long>(co) * Cin + ci) * 9 + tap;
            // deterministic mode: one slab of partials per pixel split, summed in split order by splitk_reduce_kernel
            if (partial != nullptr) partial[static_cast<long long>(blockIdx.z) * Cout * K + idx] = acc[i][j];
            else atomicAdd(dw + idx, acc[i][j]);
        }
    }
}

// dst[i] += sum_{s = 0 .. splits-1} partial[s * numel + i], in split order (deterministic split-K of the FP32 verification mode)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long numel, float* __restrict__ dst) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float t = 0.f;
        for (int sp = 0; sp < splits; ++sp) t += partial[sp * numel + i];
        dst[i] += t;
    }
}

// ------------------------------------------------------------------------------------------------
// 2x2 / stride-2 transposed convolution (ConvTranspose2d(Cin, Co, 2, 2) with bias), direct kernels.
//   x : [N,H,W,ldx] (+xoff), w : fp32 PyTorch layout [Cin][Co][2][2], out : [N,2H,2W,ldo] (+ooff)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void convT2x2_fwd_simt_kernel(const T* __restrict__ x, long long ldx, int xoff, int N, int H, int W, int Cin,
                                         const float* __restrict__ w, const float* __restrict__ bias, int Co,
                                         T* __restrict__ out, long long ldo, int ooff, int round_w_bf16, int Ho, int Wo) {
    const long long total = static_cast<long long>(N) * 2 * H * 2 * W * Co;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Co);
        long long r = i / Co;
        const int ow = static_cast<int>(r % (2 * W)); r /= 2 * W;
        const int oh = static_cast<int>(r % (2 * H));
        const int n = static_cast<int>(r / (2 * H));
        const int tap = (oh & 1) * 2 + (ow & 1);
        const T* xp = x + ((static_cast<long long>(n) * H + (oh >> 1)) * W + (ow >> 1)) * ldx + xoff;
        float acc = 0.f;
        for (int ci = 0; ci < Cin; ++ci) {
            float wv = w[(static_cast<long long>(ci) * Co + co) * 4 + tap];
            if (round_w_bf16) wv = __bfloat162float(__float2bfloat16_rn(wv));
            acc = fmaf(to_f<T>(xp[ci]), wv, acc);
        }
        out[((static_cast<long long>(n) * Ho + oh) * Wo + ow) * ldo + ooff + co] = from_f<T>(acc + bias[co]);
    }
}

// dX[n,h,w,ci] = sum_{tap,co} dO[n,2h+dy,2w+dx,co] * W[ci][co][tap]
template <typename T>
__global__ void convT2x2_dgrad_simt_kernel(const T* __restrict__ go, long long ldg, int goff, int N, int H, int W, int Cin,
                                           const float* __restrict__ w, int Co, T* __restrict__ dx, long long ldd, int doff,
                                           int Ho, int Wo) {
    const long long total = static_cast<long long>(N) * H * W * Cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int ww = static_cast<int>(r % W); r /= W;
        const int hh = static_cast<int>(r % H);
        const int n = static_cast<int>(r / H);
        float acc = 0.f;
        for (int tap = 0; tap < 4; ++tap) {
            const T* gp = go + ((static_cast<long long>(n) * Ho + 2 * hh + (tap >> 1)) * Wo + 2 * ww + (tap & 1)) * ldg + goff;
            const float* wr = w + static_cast<long long>(ci) * Co * 4 + tap;
            for (int co = 0; co < Co; ++co) acc = fmaf(to_f<T>(gp[co]), wr[co * 4], acc);
        }
        dx[((static_cast<long long>(n) * H + hh) * W + ww) * ldd + doff + ci] = from_f<T>(acc);
    }
}

// dW[ci][co][tap] += sum_px X[px][ci] * dO[2px+tap][co]; one thread per (ci,co,tap), pixel range split over blockIdx.y
template <typename T>
__global__ void convT2x2_wgrad_simt_kernel(const T* __restrict__ x, long long ldx, int xoff, const T* __restrict__ go,
                                           long long ldg, int goff, int N, int H, int W, int Cin, int Co,
                                           float* __restrict__ dw, long long px_per_split, int Ho, int Wo,
                                           float* __restrict__ partial) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(Cin) * Co * 4) return;
    const int tap = static_cast<int>(idx & 3);
    const int co = static_cast<int>((idx >> 2) % Co);
    const int ci = static_cast<int>((idx >> 2) / Co);
    const long long M = static_cast<long long>(N) * H * W;
    const long long p0 = blockIdx.y * px_per_split, p1 = min(p0 + px_per_split, M);
    float acc = 0.f;
    for (long long pm = p0; pm < p1; ++pm) {
        const int ww = static_cast<int>(pm % W), hh = static_cast<int>((pm / W) % H),
                  n = static_cast<int>(pm / (static_cast<long long>(W) * H));
        const float xv = to_f<T>(x[pm * ldx + xoff + ci]);
        const float gv = to_f<T>(
            go[((static_cast<long long>(n) * Ho + 2 * hh + (tap >> 1)) * Wo + 2 * ww + (tap & 1)) * ldg + goff + co]);
        acc = fmaf(xv, gv, acc);
    }
    if (partial != nullptr) partial[static_cast<long long>(blockIdx.y) * Cin * Co * 4 + idx] = acc;
    else atomicAdd(dw + idx, acc);
}

// column sums over the valid [N, Hv, Wv] window of a [N, Ho, Wo, ld] buffer: out[c] += sum v[n,h,w, off + c]
// (bias gradient of the transposed conv; the F.pad border of the concat buffer is excluded)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ v, long long ld, int off, long long rows, int C, float* __restrict__ out,
                              int Hv, int Wv, int Ho, int Wo, float* __restrict__ partial) {
    // blockDim.x = 256: thread -> channel c = tid % cpb, row lane = tid / cpb
    const int cpb = min(C, 64);
    const int lanes = 256 / cpb;
    const int c = blockIdx.y * cpb + threadIdx.x % cpb;
    const int rl = threadIdx.x / cpb;
    float acc = 0.f;
    if (c < C && rl < lanes)
        for (long long r = blockIdx.x * static_cast<long long>(lanes) + rl; r < rows; r += static_cast<long long>(gridDim.x) * lanes) {
            const long long w_ = r % Wv, t_ = r / Wv;
            const long long px = ((t_ / Hv) * Ho + (t_ % Hv)) * Wo + w_;
            acc += to_f<T>(v[px * ld + off + c]);
        }
    __shared__ float s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0 && c < C) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += s[l * cpb + threadIdx.x % cpb];
        if (partial != nullptr) partial[static_cast<long long>(blockIdx.x) * C + c] = t;
        else atomicAdd(out + c, t);
    }
}

}  // namespace onet
