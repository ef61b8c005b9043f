// The first convolution of the U-Net (1 -> 64 or 3 -> 64 channels at FULL resolution, Onet_vanilla_20240606.py:111) is the one
// layer whose raw output is 64 x larger than its input: storing Y and reading it back (BatchNorm apply, BatchNorm backward
// twice, weight gradient) is 5 passes over a 1 GB tensor per step, while RECOMPUTING a pixel's 64 outputs from its 3 x 3 x Cin
// patch costs 9 * Cin * 64 FMAs.  These kernels never materialise Y (nor dY):
//
//   first_conv_fwd_kernel<MODE_STATS>   : BatchNorm partial sums of the conv output                       (reads x)
//   first_conv_fwd_kernel<MODE_APPLY>   : act = relu(bn(conv(x)))                                         (reads x, writes act)
//   first_conv_bwd_reduce_kernel        : s1 = sum dz, s2 = sum dz * (y - mean) * invstd, dz = relu'(.) g  (reads x, g)
//   first_conv_bwd_wgrad_kernel         : dY = BatchNorm backward in registers, dW += dY (x) patch         (reads x, g)
//
// The recomputed y is rounded to the storage type exactly where the stored tensor would have been (round_to<T>), and the FMA
// order is the one of conv_first_fwd_rows_kernel (simt_conv.cuh): the forward results (statistics, activations) are bit-identical
// to conv_first_fwd -> bn_relu_apply, and the backward evaluates the same expressions per element as bn_relu_bwd ->
// conv_first_wgrad (its sums are added in a different order).  Thread mapping as in simt_conv.cuh: a thread owns 4 adjacent
// columns x CPT channels and walks down ROWS image rows with the 3 x 6 input patch in registers; CPT is chosen so that the
// 9 * Cin * CPT weight-gradient accumulators plus the patch stay inside 128 registers (Cin = 1: 4 channels, Cin = 3: 1).
#pragma once
#include "elementwise.cuh"

namespace onet {

enum FirstMode : int { FIRST_STATS = 0, FIRST_APPLY = 1 };

template <typename T, int CIN, int ROWS, int MODE>
__global__ void __launch_bounds__(256, CIN == 1 ? 3 : 2)
first_conv_fwd_kernel(const T* __restrict__ in, int N, int H, int W, const T* __restrict__ wp, double* __restrict__ stat_sum,
                      double* __restrict__ stat_sq, const float* __restrict__ scale, const float* __restrict__ shift,
                      int group_images, T* __restrict__ out) {
    constexpr int K = 9 * CIN;
    __shared__ float ws[K][64];
    __shared__ float s_red[MODE == FIRST_STATS ? 16 : 1][256];
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
    const int g = min(n / group_images, 1);
    float a1[8] = {}, a2[8] = {};
    float sc[8], sh[8];
    if (MODE == FIRST_APPLY) {
        load8<float>(scale + g * 64 + oc * 8, sc);
        load8<float>(shift + g * 64 + oc * 8, sh);
    }
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
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float f = round_to<T>(acc[j][i]);          // the value a stored Y would hold
                    if (MODE == FIRST_STATS) {
                        a1[i] += f;
                        a2[i] = fmaf(f, f, a2[i]);
                    } else {
                        o[i] = bn_relu_value<T>(f, sc[i], sh[i]);
                    }
                }
                if (MODE == FIRST_APPLY) store8<T>(out + (p0 + j) * 64 + oc * 8, o);
            }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx)
#pragma unroll
                for (int c = 0; c < CIN; ++c) { x[0][cidx][c] = x[1][cidx][c]; x[1][cidx][c] = x[2][cidx][c]; }
        }
    }
    if (MODE == FIRST_STATS) {       // same block reduction and double atomics as conv_first_fwd_rows_kernel
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

// Recomputed y for this thread's 4 columns x CPT channels (channel group cg) of image row h, as the forward kernels produce it.
template <typename T, int CIN, int CPT>
__device__ __forceinline__ void first_recompute_y(const float (&x)[3][6][CIN], const float (*ws)[64], int cg, float (&y)[4][CPT]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < CPT; ++i) y[j][i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const float wv = ws[t * CIN + c][cg * CPT + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j][i] = fmaf(x[t / 3][j + t % 3][c], wv, y[j][i]);
            }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < CPT; ++i) y[j][i] = round_to<T>(y[j][i]);
}

template <typename T, int CPT>
__device__ __forceinline__ void first_load_g(const T* __restrict__ g, long long px, int cg, float (&gv)[CPT]) {
    const T* src = g + px * 64 + cg * CPT;
    if constexpr (CPT == 4 && sizeof(T) == 2) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
        gv[0] = __uint_as_float(u.x << 16); gv[1] = __uint_as_float(u.x & 0xffff0000u);
        gv[2] = __uint_as_float(u.y << 16); gv[3] = __uint_as_float(u.y & 0xffff0000u);
    } else if constexpr (CPT == 4 && sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        gv[0] = v.x; gv[1] = v.y; gv[2] = v.z; gv[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < CPT; ++i) gv[i] = to_f<T>(src[i]);
    }
}

struct FirstBwdArgs {
    int N, H, W, group_images;
    const float* scale; const float* shift; const float* mean; const float* invstd;   // [G][64]
    double* sums;                                                                     // [G][2][64]
    double count;
    float* dw;                                                                        // [64][CIN][3][3]
    float* partial;                                                                   // per-block partials (deterministic mode) or nullptr
};

// Pass 1: BatchNorm-backward sums.  Thread = 4 columns x CPT channels; block = NG channel groups x LANES column groups of one
// image row range (one statistics group per block).
template <typename T, int CIN, int CPT, int ROWS>
__global__ void __launch_bounds__(256, 2)
first_conv_bwd_reduce_kernel(const T* __restrict__ in, const T* __restrict__ wp, const T* __restrict__ g, const FirstBwdArgs a) {
    constexpr int K = 9 * CIN;
    constexpr int NG = 64 / CPT, LANES = 256 / NG;
    __shared__ float ws[K][64];
    __shared__ float s_red[2][64];
    for (int i = threadIdx.x; i < K * 64; i += 256) ws[i % K][i / K] = to_f<T>(wp[i]);
    if (threadIdx.x < 128) s_red[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
    __syncthreads();
    const int H = a.H, W = a.W;
    const int cg = threadIdx.x % NG, ln = threadIdx.x / NG;
    const int WG = W >> 2, wgb = (WG + LANES - 1) / LANES, chunks = (H + ROWS - 1) / ROWS;
    int b = blockIdx.x;
    const int wg = (b % wgb) * LANES + ln; b /= wgb;
    const int h0 = (b % chunks) * ROWS;
    const int n = b / chunks;
    const int h1 = min(H, h0 + ROWS);
    const int w0 = wg * 4;
    const long long nb = static_cast<long long>(n) * H * W;
    const int grp = min(n / a.group_images, 1);
    float sc[CPT], sh[CPT], mu[CPT], acc1[CPT], acc2[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = grp * 64 + cg * CPT + i;
        sc[i] = a.scale[c]; sh[i] = a.shift[c]; mu[i] = a.mean[c];
        acc1[i] = 0.f; acc2[i] = 0.f;
    }
    if (wg < WG) {
        float x[3][6][CIN];
        load_row6<T, CIN>(in, nb, h0 - 1, w0, H, W, x[0]);
        load_row6<T, CIN>(in, nb, h0, w0, H, W, x[1]);
        for (int h = h0; h < h1; ++h) {
            load_row6<T, CIN>(in, nb, h + 1, w0, H, W, x[2]);
            const long long p0 = nb + static_cast<long long>(h) * W + w0;
            float gv[4][CPT];
#pragma unroll
            for (int j = 0; j < 4; ++j) first_load_g<T, CPT>(g, p0 + j, cg, gv[j]);
            float y[4][CPT];
            first_recompute_y<T, CIN, CPT>(x, ws, cg, y);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    const float dz = relu_open(y[j][i], sc[i], sh[i]) ? gv[j][i] : 0.f;
                    acc1[i] += dz;
                    acc2[i] = fmaf(dz, y[j][i] - mu[i], acc2[i]);
                }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx)
#pragma unroll
                for (int c = 0; c < CIN; ++c) { x[0][cidx][c] = x[1][cidx][c]; x[1][cidx][c] = x[2][cidx][c]; }
        }
    }
    // lanes of one channel group sit NG threads apart inside a warp; then shared-memory float atomics per block and one
    // double atomic per channel and block (the order-dependence is at the 1e-16 level of the double sums)
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        float v1 = acc1[i], v2 = acc2[i] * a.invstd[grp * 64 + cg * CPT + i];
        for (int off = NG; off < 32; off <<= 1) {
            v1 += __shfl_xor_sync(0xffffffffu, v1, off);
            v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        acc1[i] = v1; acc2[i] = v2;
    }
    for (int wp_ = 0; wp_ < 8; ++wp_) {          // warps add one after the other: fixed order
        if ((threadIdx.x >> 5) == wp_ && (threadIdx.x & 31) < NG) {
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                s_red[0][cg * CPT + i] += acc1[i];
                s_red[1][cg * CPT + i] += acc2[i];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 128) {
        const int stat = threadIdx.x >> 6, c = threadIdx.x & 63;
        atomicAdd(a.sums + (static_cast<long long>(grp) * 2 + stat) * 64 + c, static_cast<double>(s_red[stat][c]));
    }
}

// Pass 2: dY = sc * dz - k1 - (y - mean) * k2 in registers (rounded to the storage type, as the stored dY would be), and the weight
// gradient dW[co][ci][tap] += sum dY[p][co] * x[p + tap][ci] - the body of conv_first_wgrad_rows_kernel with dY computed in place.
template <typename T, int CIN, int CPT, int ROWS>
__global__ void __launch_bounds__(256, 2)
first_conv_bwd_wgrad_kernel(const T* __restrict__ in, const T* __restrict__ wp, const T* __restrict__ g, const FirstBwdArgs a) {
    constexpr int K = 9 * CIN;
    constexpr int NG = 64 / CPT, LANES = 256 / NG;
    __shared__ float ws[K][64];
    __shared__ float s_acc[64 * K];
    for (int i = threadIdx.x; i < K * 64; i += 256) ws[i % K][i / K] = to_f<T>(wp[i]);
    for (int i = threadIdx.x; i < 64 * K; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    const int H = a.H, W = a.W;
    const int cg = threadIdx.x % NG, ln = threadIdx.x / NG;
    const int WG = W >> 2, wgb = (WG + LANES - 1) / LANES, chunks = (H + ROWS - 1) / ROWS;
    int b = blockIdx.x;
    const int wg = (b % wgb) * LANES + ln; b /= wgb;
    const int h0 = (b % chunks) * ROWS;
    const int n = b / chunks;
    const int h1 = min(H, h0 + ROWS);
    const int w0 = wg * 4;
    const long long nb = static_cast<long long>(n) * H * W;
    const int grp = min(n / a.group_images, 1);
    const float inv_n = static_cast<float>(1.0 / a.count);
    float sc[CPT], sh[CPT], mu[CPT], k1[CPT], k2[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = grp * 64 + cg * CPT + i;
        sc[i] = a.scale[c]; sh[i] = a.shift[c]; mu[i] = a.mean[c];
        const float m1 = static_cast<float>(a.sums[(static_cast<long long>(grp) * 2 + 0) * 64 + cg * CPT + i]) * inv_n;
        const float m2 = static_cast<float>(a.sums[(static_cast<long long>(grp) * 2 + 1) * 64 + cg * CPT + i]) * inv_n;
        k1[i] = sc[i] * m1;
        k2[i] = sc[i] * a.invstd[c] * m2;
    }
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
            for (int j = 0; j < 4; ++j) first_load_g<T, CPT>(g, p0 + j, cg, gv[j]);
            {
                float y[4][CPT];
                first_recompute_y<T, CIN, CPT>(x, ws, cg, y);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        const float dz = relu_open(y[j][i], sc[i], sh[i]) ? gv[j][i] : 0.f;
                        gv[j][i] = round_to<T>(fmaf(sc[i], dz, -fmaf(y[j][i] - mu[i], k2[i], k1[i])));  // dY as bn_bwd_px_kernel stores it
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
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k)
            for (int off = NG; off < 32; off <<= 1) acc[i][k] += __shfl_xor_sync(0xffffffffu, acc[i][k], off);
    for (int wp_ = 0; wp_ < 8; ++wp_) {
        if ((threadIdx.x >> 5) == wp_ && (threadIdx.x & 31) < NG) {
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
        if (a.partial != nullptr) a.partial[static_cast<long long>(blockIdx.x) * (64 * K) + idx] = s_acc[i];
        else atomicAdd(a.dw + idx, s_acc[i]);
    }
}

}  // namespace onet
