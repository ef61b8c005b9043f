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

// ROUND_Y: the recomputed conv output is rounded to the storage type before use (what a stored Y would hold: in_chns = 3, whose
// statistics come from the same rounded values); in_chns = 1 uses the fp32 value everywhere - its statistics are analytic
// (first_gram_kernel) and nothing of this layer's output or output gradient is ever stored, so there is nothing to round.
template <typename T, int CIN, int ROWS, int MODE, bool ROUND_Y = true>
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
                    const float f = ROUND_Y ? round_to<T>(acc[j][i]) : acc[j][i];          // the value a stored Y would hold
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


// =====================================================================================================
// in_chns = 1: statistics and the "y-dependent" part of the weight gradient in closed form.
//
// The conv output is LINEAR in the 9-pixel patch v_p:  y[p][c] = w_c . v_p.  With the patch moments of a statistics group,
//     S[k] = sum_p v_p[k],      G[k][k'] = sum_p v_p[k] v_p[k']            (9 + 81 numbers, first_gram_kernel)
// the BatchNorm statistics need no pass over the 64-channel tensor:  sum_p y = w_c . S,  sum_p y^2 = w_c^T G w_c.
// Backward: dY = sc dz - k1 - k2 (y - mu)  (k1 = sc s1/n, k2 = sc invstd s2/n) gives
//     dW[c][k] = sum_p dY[p][c] v_p[k] = sc A[c][k] - k1 S[k] - k2 ((G w_c)[k] - mu S[k]),   A[c][k] = sum_p dz[p][c] v_p[k],
// so ONE pass over g (first_conv_bwd_fused_kernel: recompute y for the ReLU mask, accumulate s1, s2 and A) replaces the reduce pass
// + the weight-gradient pass, and neither y nor dY is rounded or stored anywhere.
// =====================================================================================================
constexpr int kGramK = 9;
constexpr int kGramSize = kGramK + kGramK * kGramK;      // S[9] then G[9][9], doubles, per statistics group

template <typename T, int ROWS>
__global__ void __launch_bounds__(256, 3)
first_gram_kernel(const T* __restrict__ in, int N, int H, int W, int group_images, double* __restrict__ gram) {
    constexpr int NACC = kGramK + kGramK * (kGramK + 1) / 2;      // S + upper triangle of G = 54
    __shared__ float s_red[8][NACC];
    const int WG = W >> 2, wgb = (WG + 255) >> 8, chunks = (H + ROWS - 1) / ROWS;
    int b = blockIdx.x;
    const int wg = (b % wgb) * 256 + threadIdx.x; b /= wgb;
    const int h0 = (b % chunks) * ROWS;
    const int n = b / chunks;
    const int h1 = min(H, h0 + ROWS);
    const int w0 = wg * 4;
    const long long nb = static_cast<long long>(n) * H * W;
    const int grp = min(n / group_images, 1);
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    if (wg < WG) {
        float x[3][6][1];
        load_row6<T, 1>(in, nb, h0 - 1, w0, H, W, x[0]);
        load_row6<T, 1>(in, nb, h0, w0, H, W, x[1]);
        for (int h = h0; h < h1; ++h) {
            load_row6<T, 1>(in, nb, h + 1, w0, H, W, x[2]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[kGramK];
#pragma unroll
                for (int t = 0; t < 9; ++t) v[t] = x[t / 3][j + t % 3][0];
                int o = kGramK;
#pragma unroll
                for (int k = 0; k < kGramK; ++k) {
                    acc[k] += v[k];
#pragma unroll
                    for (int k2 = k; k2 < kGramK; ++k2) { acc[o] = fmaf(v[k], v[k2], acc[o]); ++o; }
                }
            }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) { x[0][cidx][0] = x[1][cidx][0]; x[1][cidx][0] = x[2][cidx][0]; }
        }
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float v = acc[i];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        float t = 0.f;
        for (int wv = 0; wv < static_cast<int>(blockDim.x >> 5); ++wv) t += s_red[wv][threadIdx.x];       // fixed order inside the block
        // expand the triangle index into (k, k2) and add to both symmetric entries
        int i = threadIdx.x;
        double* dst = gram + static_cast<long long>(grp) * kGramSize;
        if (i < kGramK) {
            atomicAdd(dst + i, static_cast<double>(t));
        } else {
            i -= kGramK;
            int k = 0;
            while (i >= kGramK - k) { i -= kGramK - k; ++k; }
            const int k2 = k + i;
            atomicAdd(dst + kGramK + k * kGramK + k2, static_cast<double>(t));
            if (k2 != k) atomicAdd(dst + kGramK + k2 * kGramK + k, static_cast<double>(t));
        }
    }
}

// stat_sum[g][c] = w_c . S_g,  stat_sq[g][c] = w_c^T G_g w_c   (the sums onet_bn_finalize expects); one thread per (g, c).
// K = 9 * in_chns patch elements (tap-major, channel fastest: the order of the packed weights); moments of group g: S[K] then G[K][K].
template <typename T>
__global__ void first_stats_from_gram_kernel(const T* __restrict__ wp, const double* __restrict__ gram, int G, int K,
                                             double* __restrict__ stat_sum, double* __restrict__ stat_sq) {
    const int c = threadIdx.x, g = blockIdx.x;
    if (c >= 64 || g >= G) return;
    const double* S = gram + static_cast<long long>(g) * (K + K * K);
    const double* Gm = S + K;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < K; ++k) {
        const double wk = static_cast<double>(to_f<T>(wp[c * K + k]));
        s += wk * S[k];
        double r = 0.0;
        for (int k2 = 0; k2 < K; ++k2) r += Gm[k * K + k2] * static_cast<double>(to_f<T>(wp[c * K + k2]));
        q += wk * r;
    }
    stat_sum[g * 64 + c] = s;
    stat_sq[g * 64 + c] = q;
}

struct FirstFusedArgs {
    int N, H, W, group_images;
    const float* scale; const float* shift; const float* mean; const float* invstd;   // [G][64]
    double* sums;                                                                     // [G][2][64]
    float* acc_a;                                                                     // A [G][64][9] (atomics), or
    float* partial;                                                                   // per-block partials [blocks][64*9] (deterministic mode)
};

// One backward pass over g (in_chns = 1): s1 = sum dz, s2 = sum dz (y - mu) invstd, A[c][k] = sum dz v[k].
template <typename T, int CPT, int ROWS>
__global__ void __launch_bounds__(256, 2)
first_conv_bwd_fused_kernel(const T* __restrict__ in, const T* __restrict__ wp, const T* __restrict__ g, const FirstFusedArgs a) {
    constexpr int CIN = 1, K = 9;
    constexpr int NG = 64 / CPT, LANES = 256 / NG;
    __shared__ float ws[K][64];
    __shared__ float s_acc[64 * K];
    __shared__ float s_red[2][64];
    for (int i = threadIdx.x; i < K * 64; i += 256) ws[i % K][i / K] = to_f<T>(wp[i]);
    for (int i = threadIdx.x; i < 64 * K; i += 256) s_acc[i] = 0.f;
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
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    float y = 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t) y = fmaf(x[t / 3][j + t % 3][0], ws[t][cg * CPT + i], y);
                    const float dz = relu_open(y, sc[i], sh[i]) ? gv[j][i] : 0.f;
                    acc1[i] += dz;
                    acc2[i] = fmaf(dz, y - mu[i], acc2[i]);
                    gv[j][i] = dz;
                }
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < CPT; ++i) acc[i][t] = fmaf(gv[j][i], x[t / 3][j + t % 3][0], acc[i][t]);
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) { x[0][cidx][0] = x[1][cidx][0]; x[1][cidx][0] = x[2][cidx][0]; }
        }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        float v1 = acc1[i], v2 = acc2[i] * a.invstd[grp * 64 + cg * CPT + i];
        for (int off = NG; off < 32; off <<= 1) {
            v1 += __shfl_xor_sync(0xffffffffu, v1, off);
            v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        acc1[i] = v1; acc2[i] = v2;
#pragma unroll
        for (int k = 0; k < K; ++k)
            for (int off = NG; off < 32; off <<= 1) acc[i][k] += __shfl_xor_sync(0xffffffffu, acc[i][k], off);
    }
    for (int wp_ = 0; wp_ < 8; ++wp_) {          // warps add one after the other: fixed order
        if ((threadIdx.x >> 5) == wp_ && (threadIdx.x & 31) < NG) {
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                s_red[0][cg * CPT + i] += acc1[i];
                s_red[1][cg * CPT + i] += acc2[i];
#pragma unroll
                for (int k = 0; k < K; ++k) s_acc[(cg * CPT + i) * K + k] += acc[i][k];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 128) {
        const int stat = threadIdx.x >> 6, c = threadIdx.x & 63;
        atomicAdd(a.sums + (static_cast<long long>(grp) * 2 + stat) * 64 + c, static_cast<double>(s_red[stat][c]));
    }
    for (int i = threadIdx.x; i < 64 * K; i += 256) {
        if (a.partial != nullptr) a.partial[static_cast<long long>(blockIdx.x) * (64 * K) + i] = s_acc[i];
        else atomicAdd(a.acc_a + static_cast<long long>(grp) * 64 * K + i, s_acc[i]);
    }
}

// dW[c][k] += sum_g [ sc A_g[c][k] - k1 S_g[k] - k2 ((G_g w_c)[k] - mu S_g[k]) ];  one thread per (c, k), K = 9 * CIN.
// derive_s2: sums[g][1][c] = sum dz (y - mu) invstd is not given but follows from A (y = w_c . v):  invstd (w_c . A_g[c][:] - mu s1);
// every thread of channel c evaluates it for itself and the k == 0 thread stores it for bn_param_grad_kernel.
template <typename T>
__global__ void first_bwd_assemble_kernel(const T* __restrict__ wp, const double* __restrict__ gram, const float* __restrict__ acc_a,
                                          double* __restrict__ sums, const float* __restrict__ scale, const float* __restrict__ mean,
                                          const float* __restrict__ invstd, int G, double count, float* __restrict__ dw, int derive_s2,
                                          int CIN) {
    const int K = 9 * CIN;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * K) return;
    const int c = i / K, k = i % K;
    const double inv_n = 1.0 / count;            // 0 for eval-mode statistics (count = inf): the mean / projection terms vanish
    double t = 0.0;
    for (int g = 0; g < G; ++g) {
        const double* S = gram + static_cast<long long>(g) * (K + K * K);
        const double* Gm = S + K;
        const double sc = scale[g * 64 + c], mu = mean[g * 64 + c], is = invstd[g * 64 + c];
        const float* Ac = acc_a + (static_cast<long long>(g) * 64 + c) * K;
        double s2;
        if (derive_s2) {
            double wa = 0.0;
            for (int k2i = 0; k2i < K; ++k2i) wa += static_cast<double>(to_f<T>(wp[c * K + k2i])) * static_cast<double>(Ac[k2i]);
            s2 = is * (wa - mu * sums[(g * 2 + 0) * 64 + c]);
            if (k == 0) sums[(g * 2 + 1) * 64 + c] = s2;
        } else {
            s2 = sums[(g * 2 + 1) * 64 + c];
        }
        const double k1 = sc * sums[(g * 2 + 0) * 64 + c] * inv_n;
        const double k2 = sc * is * s2 * inv_n;
        double gw = 0.0;
        for (int k2i = 0; k2i < K; ++k2i) gw += Gm[k * K + k2i] * static_cast<double>(to_f<T>(wp[c * K + k2i]));
        t += sc * static_cast<double>(Ac[k]) - k1 * S[k] - k2 * (gw - mu * S[k]);
    }
    // k = tap * CIN + ci (packed order);  dw is [64][CIN][3][3]
    dw[(static_cast<long long>(c) * CIN + k % CIN) * 9 + k / CIN] += static_cast<float>(t);
}

}  // namespace onet
