// HBM-bound kernels of the Onet path: BatchNorm statistics finalisation, BN+ReLU(+2x2 max-pool) apply that
// writes straight into the skip-concat buffer, the matching backward passes (ReLU mask, max-pool routing and
// BatchNorm backward fused), the dot-product head + sigmoid + JSD loss forward/backward, input preparation,
// weight packing and Adam.  All activations NHWC; 8 channels (16 B of bf16 / 32 B of fp32) per thread access.
#pragma once
#include "simt_conv.cuh"
#include "fastdiv.cuh"

namespace onet {

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p) = u;
}
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f<T>(from_f<T>(v)); }

constexpr float kBnEps = 1e-5f;

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics -> per-group (mean, invstd, scale, shift) and the running-buffer update.
// stats layout: [G][C] doubles for sum and sumsq; outputs [G][C] floats.  One thread per channel; the
// groups (twin branches) are folded into the running buffers SEQUENTIALLY, group 0 first, which is what
// two successive calls of the same nn.BatchNorm2d do in the reference (Onet_vanilla_20240606.py:175,181).
// For the non-shared twin the per-group pointers differ.
// ------------------------------------------------------------------------------------------------
struct BnGroupPtrs {
    const float* gamma[2];
    const float* beta[2];
    float* running_mean[2];
    float* running_var[2];
};

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sq, int G, int C,
                                   double count, BnGroupPtrs ptrs, float momentum, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int g = 0; g < G; ++g) {
        const double m = sum[g * C + c] / count;
        double var = sq[g * C + c] / count - m * m;
        var = var > 0.0 ? var : 0.0;
        const float mf = static_cast<float>(m);
        const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
        const float sc = ptrs.gamma[g][c] * is;
        mean[g * C + c] = mf;
        invstd[g * C + c] = is;
        scale[g * C + c] = sc;
        shift[g * C + c] = ptrs.beta[g][c] - mf * sc;
        if (ptrs.running_mean[g] != nullptr) {
            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
            ptrs.running_mean[g][c] = (1.f - momentum) * ptrs.running_mean[g][c] + momentum * mf;
            ptrs.running_var[g][c] = (1.f - momentum) * ptrs.running_var[g][c] + momentum * static_cast<float>(unb);
        }
    }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_prepare_kernel(int G, int C, BnGroupPtrs ptrs, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int g = 0; g < G; ++g) {
        const float sc = ptrs.gamma[g][c] / sqrtf(ptrs.running_var[g][c] + kBnEps);
        scale[g * C + c] = sc;
        shift[g * C + c] = ptrs.beta[g][c] - ptrs.running_mean[g][c] * sc;
    }
}

// ------------------------------------------------------------------------------------------------
// BN + ReLU apply, optional fused 2x2 max-pool.  One thread = one 2x2 pixel window x 8 channels, UNR windows
// in flight per thread (all loads issued before the first use).
//   y   : raw conv output [N,H,W,C]
//   out : [N,H,W,ldo] (+ooff)  (e.g. the skip half of a concat buffer)
//   pool: [N,H/2,W/2,C] or nullptr (floor semantics of nn.MaxPool2d(2))
// The backward pass recomputes the window arg-max from y with the same arithmetic, so nothing else is saved.
// ------------------------------------------------------------------------------------------------
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 u; };
template <typename T> __device__ __forceinline__ Raw8<T> ldraw(const T* p);
template <> __device__ __forceinline__ Raw8<float> ldraw<float>(const float* p) {
    Raw8<float> r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p + 4));
    return r;
}
template <> __device__ __forceinline__ Raw8<__nv_bfloat16> ldraw<__nv_bfloat16>(const __nv_bfloat16* p) {
    Raw8<__nv_bfloat16> r;
    r.u = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
template <typename T> __device__ __forceinline__ Raw8<T> zero_raw();
template <> __device__ __forceinline__ Raw8<float> zero_raw<float>() {
    Raw8<float> r; r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; return r;
}
template <> __device__ __forceinline__ Raw8<__nv_bfloat16> zero_raw<__nv_bfloat16>() {
    Raw8<__nv_bfloat16> r; r.u = make_uint4(0u, 0u, 0u, 0u); return r;
}
__device__ __forceinline__ void unpack(const Raw8<float>& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
    const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// post-activation value exactly as the forward pass stores it (the backward pass relies on this being one expression)
template <typename T>
__device__ __forceinline__ float bn_relu_value(float y, float sc, float sh) {
    return round_to<T>(fmaxf(fmaf(y, sc, sh), 0.f));
}

// ReLU mask of the backward pass: the stored activation round(max(y * sc + sh, 0)) is > 0 exactly when the fp32 pre-activation
// is (rounding to bf16 keeps the sign and, with fp32's exponent range, never flushes a positive value to zero)
__device__ __forceinline__ bool relu_open(float y, float sc, float sh) { return fmaf(y, sc, sh) > 0.f; }

template <typename T, int UNR>
__global__ void __launch_bounds__(256, 3)
bn_relu_apply_kernel(const T* __restrict__ y, int N, int H, int W, int C, const float* __restrict__ scale,
                     const float* __restrict__ shift, int group_images, T* __restrict__ out, long long ldo, int ooff,
                     T* __restrict__ pool, unsigned short* __restrict__ pool_arg, FastDiv fd_oc, FastDiv fd_w2, FastDiv fd_h2) {
    // (window, channel octet) index < 2^31 (checked by the launcher): decomposed with multiply-high divisions - the three 64-bit
    // div/mod pairs this loop used to do were ~300 instructions per window, twice the work on its 32 elements
    const int OC = C >> 3, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1, HP = H >> 1, WP = W >> 1;
    const int total = N * H2 * W2 * OC;
    const int nthreads = static_cast<int>(gridDim.x * blockDim.x);
    for (int idx0 = static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x); idx0 < total; idx0 += nthreads * UNR) {
        Raw8<T> r[UNR][4];
        int oc_[UNR], n_[UNR], h2_[UNR], w2_[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int idx = idx0 + u * nthreads;
            const bool live = idx < total && idx >= 0;
            const int q0 = live ? idx : 0;
            const int q1 = fd_div(q0, fd_oc);
            oc_[u] = q0 - q1 * OC;
            const int q2 = fd_div(q1, fd_w2);
            w2_[u] = q1 - q2 * W2;
            const int q3 = fd_div(q2, fd_h2);
            h2_[u] = q2 - q3 * H2;
            n_[u] = live ? q3 : -1;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int h = 2 * h2_[u] + (d >> 1), w = 2 * w2_[u] + (d & 1);
                const bool ok = live && h < H && w < W;
                r[u][d] = ok ? ldraw<T>(y + ((static_cast<long long>(n_[u]) * H + h) * W + w) * C + oc_[u] * 8) : zero_raw<T>();
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (n_[u] < 0) break;
            const int n = n_[u], oc = oc_[u];
            const int g = n >= group_images ? 1 : 0;
            float sc[8], sh[8], mx[8];
            load8<float>(scale + g * C + oc * 8, sc);
            load8<float>(shift + g * C + oc * 8, sh);
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = -1.f;           // post-ReLU values are >= 0: position 0 always enters
            uint32_t best = 0;                                  // first maximum of the window per channel, 2 bits each
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int h = 2 * h2_[u] + (d >> 1), w = 2 * w2_[u] + (d & 1);
                if (h < H && w < W) {
                    float v[8];
                    unpack(r[u][d], v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] = bn_relu_value<T>(v[i], sc[i], sh[i]);
                        if (v[i] > mx[i]) { mx[i] = v[i]; best = (best & ~(3u << (2 * i))) | (static_cast<uint32_t>(d) << (2 * i)); }
                    }
                    store8<T>(out + ((static_cast<long long>(n) * H + h) * W + w) * ldo + ooff + oc * 8, v);
                }
            }
            if (pool != nullptr && h2_[u] < HP && w2_[u] < WP) {
                const long long pw = (static_cast<long long>(n) * HP + h2_[u]) * WP + w2_[u];
                store8<T>(pool + pw * C + oc * 8, mx);
                // the routing of the pooled gradient, kept for the backward pass: 2 bytes per 32 activations instead of
                // recomputing 32 BatchNorm + ReLU + rounding + compare chains there
                if (pool_arg != nullptr) pool_arg[pw * OC + oc] = static_cast<unsigned short>(best);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of (BN -> ReLU [-> skip / 2x2 max-pool]).  The gradient w.r.t. the post-ReLU activation is the
// sum of up to two dense sources (g1, g2: e.g. the skip half of d(concat) and the head's dL) and a pooled
// source gp (gradient of the max-pool output, routed to the first maximum of each 2x2 window exactly as
// ATen's max_pool2d backward does; the arg-max is recomputed from y).
// Pass 1 (reduce): s1 = sum dZ, s2 = sum dZ * xhat per group/channel  (accumulated as sum dZ*(y-mean), scaled by
//                  invstd once per thread at the end).
// Pass 2 (apply):  dY = gamma*invstd * (dZ - s1/n - xhat * s2/n)  =  sc*dZ - k1 - (y-mean)*k2.
// Two thread mappings: one pixel x 8 channels (no pooled source) and one 2x2 window x 8 channels (pooled source).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct BnBwdArgs {
    const T* y; int N, H, W, C;
    const float* scale; const float* shift; const float* mean; const float* invstd;   // [G][C]
    int group_images;
    const T* g1; long long ld1; int off1;
    const T* g2; long long ld2; int off2;
    const T* gp;                                 // [N,H/2,W/2,C] or nullptr
    const unsigned short* gp_arg;                // [N,H/2,W/2,C/8] arg-max bits written by bn_relu_apply_kernel, or nullptr (recompute)
    // Last layer of the U-Net fused with the head backward (bn_bwd_px_kernel only): the incoming gradient is not a stored
    // tensor but g[p][c] = round(g1[p][c] * g1_scale[p]) with g1 = the local feature L and g1_scale = the per-pixel dV of
    // the head (head_bwd_scalars_kernel); the reduce pass also writes dL[p][c] = g1_scale[p] * relu(bn(y))[p][c] + dl_add[p].
    const float* g1_scale;                       // [N*H*W] or nullptr
    const float* dl_add;                         // [N*H*W]
    T* dl_out;                                   // [N,H,W,C] or nullptr
    double* sums;                                // [G][2][C]
    double count;                                // elements per channel per group
    T* dy;                                       // [N,H,W,C]
    FastDiv fd_w2, fd_h2;                        // window mapping: / ceil(W/2), / ceil(H/2)
};

// block-wide reduction of the per-thread (acc1, acc2) over the pixel lanes and one double atomic per channel
template <typename T, int NT = 256>
__device__ __forceinline__ void bn_bwd_block_reduce(const BnBwdArgs<T>& a, int g, const float (&acc1)[8],
                                                    const float (&acc2)[8], float (*s_red)[NT]) {
    const int OC = a.C >> 3, LANES = NT / OC;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s_red[i][threadIdx.x] = acc1[i];
        s_red[8 + i][threadIdx.x] = acc2[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * OC; t += NT) {
        const int which = t / OC, o = t % OC;
        float s = 0.f;
        for (int l = 0; l < LANES; ++l) s += s_red[which][l * OC + o];
        const int stat = which >> 3, i = which & 7;
        atomicAdd(a.sums + (static_cast<long long>(g) * 2 + stat) * a.C + o * 8 + i, static_cast<double>(s));
    }
}

// ---- pixel mapping: one thread = one pixel x 8 channels, UNROLL pixels in flight
template <typename T, int UNROLL, bool HAS_G2, bool APPLY>
__global__ void __launch_bounds__(256, 2)
bn_bwd_px_kernel(const BnBwdArgs<T> a) {
    __shared__ float s_red[APPLY ? 1 : 16][256];
    const int OC = a.C >> 3, LANES = 256 / OC;
    const int oc = threadIdx.x % OC, ln = threadIdx.x / OC;
    const int g = blockIdx.y;
    const long long HW = static_cast<long long>(a.H) * a.W;
    const long long p_begin = static_cast<long long>(g) * a.group_images * HW;
    const long long p_end = static_cast<long long>(g == static_cast<int>(gridDim.y) - 1 ? a.N : (g + 1) * a.group_images) * HW;
    constexpr bool has_g2 = HAS_G2;
    float acc1[8] = {}, acc2[8] = {};
    if (ln < LANES) {
        float sc[8], sh[8], mu[8], k1[8], k2[8];
        load8<float>(a.scale + g * a.C + oc * 8, sc);
        load8<float>(a.shift + g * a.C + oc * 8, sh);
        load8<float>(a.mean + g * a.C + oc * 8, mu);
        if (APPLY) {
            const float inv_n = static_cast<float>(1.0 / a.count);
            float is[8];
            load8<float>(a.invstd + g * a.C + oc * 8, is);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float m1 = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 0) * a.C + oc * 8 + i]) * inv_n;
                const float m2 = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 1) * a.C + oc * 8 + i]) * inv_n;
                k1[i] = sc[i] * m1;
                k2[i] = sc[i] * is[i] * m2;
            }
        }
        const long long stride = static_cast<long long>(gridDim.x) * LANES;
        for (long long p = p_begin + blockIdx.x * static_cast<long long>(LANES) + ln; p < p_end; p += stride * UNROLL) {
            Raw8<T> ry[UNROLL], rg1[UNROLL], rg2[has_g2 ? UNROLL : 1];
            float gsr[UNROLL], gar[UNROLL];          // head-fused last layer: per-pixel scalars, loaded with the tensors
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long q = p + u * stride;
                const bool ok = q < p_end;
                ry[u] = ok ? ldraw<T>(a.y + q * a.C + oc * 8) : zero_raw<T>();
                rg1[u] = ok ? ldraw<T>(a.g1 + q * a.ld1 + a.off1 + oc * 8) : zero_raw<T>();
                if (has_g2) rg2[has_g2 ? u : 0] = ok ? ldraw<T>(a.g2 + q * a.ld2 + a.off2 + oc * 8) : zero_raw<T>();
                if (a.g1_scale != nullptr) {
                    gsr[u] = ok ? __ldg(a.g1_scale + q) : 0.f;
                    gar[u] = (ok && !APPLY && a.dl_out != nullptr) ? __ldg(a.dl_add + q) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long q = p + u * stride;
                if (q >= p_end) break;
                float y[8], gg[8], o[8];
                unpack(ry[u], y);
                unpack(rg1[u], gg);
                if (has_g2) {
                    float g2v[8];
                    unpack(rg2[has_g2 ? u : 0], g2v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) gg[i] += g2v[i];
                }
                if (a.g1_scale != nullptr) {       // head-fused last layer: g = dV * L, dL = dV * H + d(a) with H = relu(bn(y))
                    const float gs = gsr[u];
                    if (!APPLY && a.dl_out != nullptr) {
                        const float ga = gar[u];
                        float dl[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) dl[i] = fmaf(gs, bn_relu_value<T>(y[i], sc[i], sh[i]), ga);
                        store8<T>(a.dl_out + q * a.C + oc * 8, dl);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) gg[i] = round_to<T>(gs * gg[i]);      // what a stored dH would hold
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dz = relu_open(y[i], sc[i], sh[i]) ? gg[i] : 0.f;
                    const float yc = y[i] - mu[i];
                    if (APPLY) {
                        o[i] = fmaf(sc[i], dz, -fmaf(yc, k2[i], k1[i]));
                    } else {
                        acc1[i] += dz;
                        acc2[i] = fmaf(dz, yc, acc2[i]);
                    }
                }
                if (APPLY) store8<T>(a.dy + q * a.C + oc * 8, o);
            }
        }
        if (!APPLY) {
            float is[8];
            load8<float>(a.invstd + g * a.C + oc * 8, is);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc2[i] *= is[i];
        }
    }
    if (!APPLY) bn_bwd_block_reduce<T>(a, g, acc1, acc2, s_red);
}

// ---- window mapping: one thread = one 2x2 window x 8 channels (pooled gradient source present)
// 128 threads x 3 CTAs per SM: a window holds 13 16-byte loads plus 40 per-channel constants, which does not fit the
// 128 registers of a 256 x 2 configuration (it spilled 136-160 B per thread); 170 registers are available here.
constexpr int kBnWinThreads = 128;
template <typename T, bool HAS_G2, bool APPLY>
__global__ void __launch_bounds__(kBnWinThreads, 3)
bn_bwd_win_kernel(const BnBwdArgs<T> a) {
    __shared__ float s_red[APPLY ? 1 : 16][kBnWinThreads];
    const int OC = a.C >> 3, LANES = kBnWinThreads / OC;
    const int oc = threadIdx.x % OC, ln = threadIdx.x / OC;
    const int g = blockIdx.y;
    const int H = a.H, W = a.W, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1, HP = H >> 1, WP = W >> 1;
    const int n_begin = g * a.group_images;
    const int n_end = (g == static_cast<int>(gridDim.y) - 1) ? a.N : (g + 1) * a.group_images;
    const int q_end = (n_end - n_begin) * H2 * W2;            // < 2^31 (checked by the launcher)
    float acc1[8] = {}, acc2[8] = {};
    if (ln < LANES) {
        float sc[8], sh[8], mu[8], k1[8], k2[8];
        load8<float>(a.scale + g * a.C + oc * 8, sc);
        load8<float>(a.shift + g * a.C + oc * 8, sh);
        load8<float>(a.mean + g * a.C + oc * 8, mu);
        if (APPLY) {
            const float inv_n = static_cast<float>(1.0 / a.count);
            float is[8];
            load8<float>(a.invstd + g * a.C + oc * 8, is);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float m1 = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 0) * a.C + oc * 8 + i]) * inv_n;
                const float m2 = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 1) * a.C + oc * 8 + i]) * inv_n;
                k1[i] = sc[i] * m1;
                k2[i] = sc[i] * is[i] * m2;
            }
        }
        const int stride = static_cast<int>(gridDim.x) * LANES;
        for (int q = static_cast<int>(blockIdx.x) * LANES + ln; q < q_end; q += stride) {
            const int t = fd_div(q, a.fd_w2);
            const int w2 = q - t * W2;
            const int t2 = fd_div(t, a.fd_h2);
            const int h2 = t - t2 * H2;
            const int n = n_begin + t2;
            Raw8<T> ry[4], rg1[4], rg2[HAS_G2 ? 4 : 1], rp;
            bool ok[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int h = 2 * h2 + (d >> 1), w = 2 * w2 + (d & 1);
                ok[d] = h < H && w < W;
                const long long px = (static_cast<long long>(n) * H + h) * W + w;
                ry[d] = ok[d] ? ldraw<T>(a.y + px * a.C + oc * 8) : zero_raw<T>();
                rg1[d] = ok[d] ? ldraw<T>(a.g1 + px * a.ld1 + a.off1 + oc * 8) : zero_raw<T>();
                if (HAS_G2) rg2[HAS_G2 ? d : 0] = ok[d] ? ldraw<T>(a.g2 + px * a.ld2 + a.off2 + oc * 8) : zero_raw<T>();
            }
            const bool pooled = h2 < HP && w2 < WP;
            rp = pooled ? ldraw<T>(a.gp + ((static_cast<long long>(n) * HP + h2) * WP + w2) * a.C + oc * 8) : zero_raw<T>();
            // first maximum of the window per channel (2 bits each), exactly as the forward pass pooled it
            uint32_t best = 0;
            if (a.gp_arg != nullptr) {
                if (pooled) best = __ldg(a.gp_arg + ((static_cast<long long>(n) * HP + h2) * WP + w2) * OC + oc);
            } else {
                float mx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) mx[i] = -1.f;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    float y[8];
                    unpack(ry[d], y);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float act = bn_relu_value<T>(y[i], sc[i], sh[i]);
                        if (act > mx[i]) { mx[i] = act; best = (best & ~(3u << (2 * i))) | (static_cast<uint32_t>(d) << (2 * i)); }
                    }
                }
            }
            float gpv[8];
            unpack(rp, gpv);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (!ok[d]) continue;
                float y[8], gg[8], o[8];
                unpack(ry[d], y);
                unpack(rg1[d], gg);
                if (HAS_G2) {
                    float g2v[8];
                    unpack(rg2[HAS_G2 ? d : 0], g2v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) gg[i] += g2v[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool is_best = pooled && ((best >> (2 * i)) & 3u) == static_cast<uint32_t>(d);
                    const float gsum = gg[i] + (is_best ? gpv[i] : 0.f);
                    const float dz = relu_open(y[i], sc[i], sh[i]) ? gsum : 0.f;
                    const float yc = y[i] - mu[i];
                    if (APPLY) {
                        o[i] = fmaf(sc[i], dz, -fmaf(yc, k2[i], k1[i]));
                    } else {
                        acc1[i] += dz;
                        acc2[i] = fmaf(dz, yc, acc2[i]);
                    }
                }
                if (APPLY) {
                    const int h = 2 * h2 + (d >> 1), w = 2 * w2 + (d & 1);
                    store8<T>(a.dy + ((static_cast<long long>(n) * H + h) * W + w) * a.C + oc * 8, o);
                }
            }
        }
        if (!APPLY) {
            float is[8];
            load8<float>(a.invstd + g * a.C + oc * 8, is);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc2[i] *= is[i];
        }
    }
    if (!APPLY) bn_bwd_block_reduce<T, kBnWinThreads>(a, g, acc1, acc2, s_red);
}

// dgamma[c] (+)= sum_g s2[g][c], dbeta[c] (+)= sum_g s1[g][c]; per-group targets may alias (shared twin)
__global__ void bn_param_grad_kernel(const double* __restrict__ sums, int G, int C, float* dgamma0, float* dbeta0,
                                     float* dgamma1, float* dbeta1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int g = 0; g < G; ++g) {
        float* dg = g == 0 ? dgamma0 : dgamma1;
        float* db = g == 0 ? dbeta0 : dbeta1;
        db[c] += static_cast<float>(sums[(static_cast<long long>(g) * 2 + 0) * C + c]);
        dg[c] += static_cast<float>(sums[(static_cast<long long>(g) * 2 + 1) * C + c]);
    }
}

// 2x2 max-pool (floor semantics) of a channel window of an NHWC buffer: the inference path, where BatchNorm + ReLU are folded
// into the convolution epilogue and only the pooled map is still missing (nn.MaxPool2d(2), Onet_vanilla_20240606.py:67)
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const T* __restrict__ in, long long ldi, int ioff, int N, int H, int W, int C, T* __restrict__ out) {
    const int OC = C >> 3, HP = H >> 1, WP = W >> 1;
    const long long total = static_cast<long long>(N) * HP * WP * OC;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int oc = static_cast<int>(idx % OC);
        long long q = idx / OC;
        const int w2 = static_cast<int>(q % WP); q /= WP;
        const int h2 = static_cast<int>(q % HP);
        const long long n = q / HP;
        Raw8<T> r[4];
#pragma unroll
        for (int d = 0; d < 4; ++d)
            r[d] = ldraw<T>(in + ((n * H + 2 * h2 + (d >> 1)) * W + 2 * w2 + (d & 1)) * ldi + ioff + oc * 8);
        float mx[8], v[8];
        unpack(r[0], mx);
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            unpack(r[d], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], v[i]);
        }
        store8<T>(out + ((n * HP + h2) * WP + w2) * C + oc * 8, mx);
    }
}

// zero channels [coff, coff+C) of the pixels with h >= Hv or w >= Wv of a [N,Ho,Wo,ld] buffer: the border F.pad adds when the
// skip tensor is one pixel larger than the up-sampled map (Onet_vanilla_20240606.py:92-96)
template <typename T>
__global__ void zero_border_kernel(T* __restrict__ buf, int N, int Ho, int Wo, long long ld, int coff, int C, int Hv, int Wv) {
    const int nb = (Ho - Hv) * Wo + Hv * (Wo - Wv);          // border pixels per image: bottom rows, then right columns
    const long long total = static_cast<long long>(N) * nb * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long r = i / C;
        const int b = static_cast<int>(r % nb);
        const long long n = r / nb;
        int h, w;
        if (b < (Ho - Hv) * Wo) { h = Hv + b / Wo; w = b % Wo; }
        else { const int bb = b - (Ho - Hv) * Wo; h = bb / (Wo - Wv); w = Wv + bb % (Wo - Wv); }
        buf[((n * Ho + h) * Wo + w) * ld + coff + c] = from_f<T>(0.f);
    }
}

// dst[c] += sums[c]: folds double column sums (e.g. the transposed convolution's bias gradient, accumulated by the
// epilogue of the convolution that produced d(concat)) into an fp32 gradient
__global__ void add_colsums_kernel(const double* __restrict__ sums, int C, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[c] += static_cast<float>(sums[c]);
}

// ------------------------------------------------------------------------------------------------
// Head: V = sum_p L_p*H_p per branch, S = softmax([Vt,Vd]), a = sum_p Lt_p, b = sum_p Ld_p, and the JSD
// loss  (1/2N) sum [ sp(-a St) + sp(a Sd) + sp(-b Sd) + sp(b St) ]  with the reference's piecewise
// softplus `sp` (Onet_vanilla_20240606.py:237-251, including its ln2 plateau below -37).
// 8 threads per pixel, 8 channels each (C = 64).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sp_ref(float x, float& val, float& der) {
    if (x <= -37.f) {            // exp(x) ~ 0+ is re-matched by the (-37,18] branch -> log(1+exp(exp(x)))
        const float e = expf(x), ee = expf(e);
        val = logf(1.f + ee);
        der = ee / (1.f + ee) * e;
    } else if (x <= 18.f) {
        const float e = expf(x);
        val = logf(1.f + e);
        der = e / (1.f + e);
    } else if (x < 33.3f) {
        const float e = expf(-x);
        val = x + e;
        der = 1.f - e;
    } else {
        val = x;
        der = 1.f;
    }
}

template <typename T>
struct HeadArgs {
    const T* L; long long ldl; int offl;     // twin local features  [2B,H,W,ldl]
    const T* Hf; long long ldh; int offh;    // twin global features [2B,H,W,ldh]
    int B; long long HW;                     // images per branch, pixels per image
    float* Vt; float* Vd; float* S;          // (B,1,H,W), (B,1,H,W), (B,2,H,W) fp32
    float* a; float* b;                      // (B,H,W) fp32 channel sums of Lt / Ld
    double* loss_acc;                        // sum of the four softplus terms over all pixels
    // backward only
    const float* gscale;                     // upstream d(loss) scalar (device) or nullptr (=> no fused loss grad)
    const float* gVt; const float* gVd; const float* gS;   // optional external gradients, fp32
    T* dL; T* dH;                            // [2B,H,W,64] dense
    // Fused with the last layer's BatchNorm + ReLU: Hf then points to that layer's RAW conv output and the head applies
    // h = relu(hf * scale + shift) itself (top / down branch constants, [64] each); nullptr = Hf holds the activation.
    const float* hsc_t; const float* hsh_t; const float* hsc_d; const float* hsh_d;
};

// head backward, per-pixel part only: dV of both branches (gv[0 .. npx) top, gv[npx .. 2 npx) down) and the gradient of the
// loss's channel sums (gab, same layout) - everything head_bwd_kernel computes before it touches L and H.
__device__ __forceinline__ void head_pixel_grads(const float* __restrict__ Vt, const float* __restrict__ Vd, const float* __restrict__ ain,
                                                 const float* __restrict__ bin, const float* gscale, const float* gVt, const float* gVd,
                                                 const float* gS, long long p, long long npx, long long HW, float& g_vt, float& g_vd,
                                                 float& g_a, float& g_b) {
    const float gsv = gscale != nullptr ? *gscale : 0.f;
    const float cc = gsv / (2.f * static_cast<float>(npx));
    const float vt = Vt[p], vd = Vd[p], sa = ain[p], sb = bin[p];
    const float mx = fmaxf(vt, vd);
    const float et = expf(vt - mx), ed = expf(vd - mx);
    const float inv = 1.f / (et + ed);
    const float st = et * inv, sd = ed * inv;
    float g_st = 0.f, g_sd = 0.f;
    g_a = 0.f; g_b = 0.f;
    if (gscale != nullptr) {
        float v, d1, d2, d3, d4;
        sp_ref(-sa * st, v, d1);
        sp_ref(sa * sd, v, d2);
        sp_ref(-sb * sd, v, d3);
        sp_ref(sb * st, v, d4);
        g_a = (-st * d1 + sd * d2) * cc;
        g_b = (-sd * d3 + st * d4) * cc;
        g_st = (-sa * d1 + sb * d4) * cc;
        g_sd = (sa * d2 - sb * d3) * cc;
    }
    if (gS != nullptr) {
        const long long n = p / HW, hw = p % HW;
        g_st += gS[(n * 2 + 0) * HW + hw];
        g_sd += gS[(n * 2 + 1) * HW + hw];
    }
    const float gsm = st * sd * (g_st - g_sd);          // softmax backward
    g_vt = gsm; g_vd = -gsm;
    if (gVt != nullptr) g_vt += gVt[p];
    if (gVd != nullptr) g_vd += gVd[p];
}

__global__ void __launch_bounds__(256)
head_bwd_scalars_kernel(const float* __restrict__ Vt, const float* __restrict__ Vd, const float* __restrict__ ain,
                        const float* __restrict__ bin, const float* gscale, const float* gVt, const float* gVd, const float* gS,
                        long long npx, long long HW, float* __restrict__ gv, float* __restrict__ gab) {
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < npx;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        float g_vt, g_vd, g_a, g_b;
        head_pixel_grads(Vt, Vd, ain, bin, gscale, gVt, gVd, gS, p, npx, HW, g_vt, g_vd, g_a, g_b);
        gv[p] = g_vt; gv[p + npx] = g_vd;
        gab[p] = g_a; gab[p + npx] = g_b;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const HeadArgs<T> a) {
    const long long npx = static_cast<long long>(a.B) * a.HW;
    const int sub = threadIdx.x & 7;
    // the 8 threads of a pixel leave the loop together, the four pixel groups of a warp need not (B*H*W % 4 != 0):
    // shuffle inside the group's own lanes only
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24);
    float hsc_t[8], hsh_t[8], hsc_d[8], hsh_d[8];
    if (a.hsc_t != nullptr) {
        load8<float>(a.hsc_t + sub * 8, hsc_t);
        load8<float>(a.hsh_t + sub * 8, hsh_t);
        load8<float>(a.hsc_d + sub * 8, hsc_d);
        load8<float>(a.hsh_d + sub * 8, hsh_d);
    }
    float lsum = 0.f;
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npx;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const long long pd = p + npx;   // same pixel of the down-branch image
        float lt[8], ht[8], ld[8], hd[8];
        load8<T>(a.L + p * a.ldl + a.offl + sub * 8, lt);
        load8<T>(a.Hf + p * a.ldh + a.offh + sub * 8, ht);
        load8<T>(a.L + pd * a.ldl + a.offl + sub * 8, ld);
        load8<T>(a.Hf + pd * a.ldh + a.offh + sub * 8, hd);
        if (a.hsc_t != nullptr) {        // Hf is the last layer's raw conv output: BatchNorm + ReLU here, no separate pass
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ht[i] = bn_relu_value<T>(ht[i], hsc_t[i], hsh_t[i]);
                hd[i] = bn_relu_value<T>(hd[i], hsc_d[i], hsh_d[i]);
            }
        }
        float vt = 0.f, vd = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            vt = fmaf(lt[i], ht[i], vt);
            vd = fmaf(ld[i], hd[i], vd);
            sa += lt[i];
            sb += ld[i];
        }
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) {
            vt += __shfl_xor_sync(gmask, vt, o);
            vd += __shfl_xor_sync(gmask, vd, o);
            sa += __shfl_xor_sync(gmask, sa, o);
            sb += __shfl_xor_sync(gmask, sb, o);
        }
        // every lane of the pixel holds the four sums: the softmax is evaluated redundantly, the four softplus terms and the six
        // stores are spread over the pixel's lanes (one lane doing all of it left 7 of 8 lanes idle through ~100 instructions)
        const float mx = fmaxf(vt, vd);
        const float et = expf(vt - mx), ed = expf(vd - mx);
        const float inv = 1.f / (et + ed);
        const float st = et * inv, sd = ed * inv;
        if (sub < 4) {
            const float xarg = sub == 0 ? -sa * st : (sub == 1 ? sa * sd : (sub == 2 ? -sb * sd : sb * st));
            float v, d;
            sp_ref(xarg, v, d);
            lsum += v;
        }
        const long long n = p / a.HW, hw = p % a.HW;
        if (sub == 0) a.Vt[p] = vt;
        else if (sub == 1) a.Vd[p] = vd;
        else if (sub == 2) a.S[(n * 2 + 0) * a.HW + hw] = st;
        else if (sub == 3) a.S[(n * 2 + 1) * a.HW + hw] = sd;
        else if (sub == 4) a.a[p] = sa;
        else if (sub == 5) a.b[p] = sb;
    }
    __shared__ float s_l[256];
    s_l[threadIdx.x] = lsum;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if (threadIdx.x < s) s_l[threadIdx.x] += s_l[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && a.loss_acc != nullptr) atomicAdd(a.loss_acc, static_cast<double>(s_l[0]));
}

template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const HeadArgs<T> a) {
    const long long npx = static_cast<long long>(a.B) * a.HW;
    const int sub = threadIdx.x & 7;
    const float gs = a.gscale != nullptr ? *a.gscale : 0.f;
    const float cc = gs / (2.f * static_cast<float>(npx));
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npx;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const long long pd = p + npx;
        const float vt = a.Vt[p], vd = a.Vd[p], sa = a.a[p], sb = a.b[p];
        const float mx = fmaxf(vt, vd);
        const float et = expf(vt - mx), ed = expf(vd - mx);
        const float inv = 1.f / (et + ed);
        const float st = et * inv, sd = ed * inv;
        float g_a = 0.f, g_b = 0.f, g_st = 0.f, g_sd = 0.f;
        if (a.gscale != nullptr) {
            float v, d1, d2, d3, d4;
            sp_ref(-sa * st, v, d1);
            sp_ref(sa * sd, v, d2);
            sp_ref(-sb * sd, v, d3);
            sp_ref(sb * st, v, d4);
            g_a = (-st * d1 + sd * d2) * cc;
            g_b = (-sd * d3 + st * d4) * cc;
            g_st = (-sa * d1 + sb * d4) * cc;
            g_sd = (sa * d2 - sb * d3) * cc;
        }
        if (a.gS != nullptr) {
            const long long n = p / a.HW, hw = p % a.HW;
            g_st += a.gS[(n * 2 + 0) * a.HW + hw];
            g_sd += a.gS[(n * 2 + 1) * a.HW + hw];
        }
        const float gsm = st * sd * (g_st - g_sd);          // softmax backward
        float g_vt = gsm, g_vd = -gsm;
        if (a.gVt != nullptr) g_vt += a.gVt[p];
        if (a.gVd != nullptr) g_vd += a.gVd[p];
        float lt[8], ht[8], ld[8], hd[8], o[8];
        load8<T>(a.L + p * a.ldl + a.offl + sub * 8, lt);
        load8<T>(a.Hf + p * a.ldh + a.offh + sub * 8, ht);
        load8<T>(a.L + pd * a.ldl + a.offl + sub * 8, ld);
        load8<T>(a.Hf + pd * a.ldh + a.offh + sub * 8, hd);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(g_vt, ht[i], g_a);
        store8<T>(a.dL + p * 64 + sub * 8, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g_vt * lt[i];
        store8<T>(a.dH + p * 64 + sub * 8, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(g_vd, hd[i], g_b);
        store8<T>(a.dL + pd * 64 + sub * 8, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g_vd * ld[i];
        store8<T>(a.dH + pd * 64 + sub * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------
// input preparation: X (B,Cin,H,W) fp32 NCHW  ->  twin batch [2B,H,W,Cin] NHWC with the complementary pair
// (X, clip(1 - X + bias, 0, 1))  (Onet_vanilla_20240606.py:175,180-181)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void prep_input_kernel(const float* __restrict__ x, int B, int Cin, long long HW, float bias, T* __restrict__ out) {
    const long long total = static_cast<long long>(B) * HW * Cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cin);
        const long long px = i / Cin;          // n*HW + hw
        const long long n = px / HW, hw = px % HW;
        const float v = x[(n * Cin + c) * HW + hw];
        out[i] = from_f<T>(v);
        out[i + total] = from_f<T>(fminf(fmaxf(1.f - v + bias, 0.f), 1.f));
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing (once per optimizer step)
//   conv  w [Co][Ci][3][3] fp32 -> wf [Co][tap][Ci] (fwd B operand) and wd [Ci][8-tap][Co] (dgrad B operand)
//   convT w [Ci][Co][2][2] fp32 -> wf [(tap,co)][ci] (fwd B operand) and wd [ci][(tap,co)] (dgrad B operand)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int Co, int Ci, T* __restrict__ wf, T* __restrict__ wd) {
    const long long total = static_cast<long long>(Co) * Ci * 9;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % 9);
        const int ci = static_cast<int>((i / 9) % Ci);
        const int co = static_cast<int>(i / (9LL * Ci));
        const T v = from_f<T>(w[i]);
        wf[(static_cast<long long>(co) * 9 + tap) * Ci + ci] = v;
        if (wd != nullptr) wd[(static_cast<long long>(ci) * 9 + (8 - tap)) * Co + co] = v;
    }
}
template <typename T>
__global__ void pack_convT_w_kernel(const float* __restrict__ w, int Ci, int Co, T* __restrict__ wf, T* __restrict__ wd) {
    const long long total = static_cast<long long>(Ci) * Co * 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i & 3);
        const int co = static_cast<int>((i >> 2) % Co);
        const int ci = static_cast<int>((i >> 2) / Co);
        const T v = from_f<T>(w[i]);
        wf[(static_cast<long long>(tap) * Co + co) * Ci + ci] = v;
        wd[static_cast<long long>(ci) * 4 * Co + tap * Co + co] = v;
    }
}

// All layers in ONE launch with coalesced reads and 64-byte write segments: a block transposes a 32 x 32 (outer x
// inner channel) tile through shared memory.
constexpr int kPackMaxLayers = 24;
struct PackJob {
    const float* w;      // conv: [Co][Ci][3][3]; convT: [Ci][Co][2][2]
    void* wf;
    void* wd;            // may be nullptr (conv only)
    int d0, d1;          // outer / inner channel counts of w (conv: Co, Ci; convT: Ci, Co)
    int taps;            // 9 (conv) or 4 (convT)
    int tile_begin;      // first block index of this layer
};
struct PackJobs {
    int n;
    PackJob job[kPackMaxLayers];
};

// One 32 x 32 (outer x inner) tile of one layer, TAPS = 9 (conv) or 4 (up-conv) as a compile-time constant: the index
// decompositions below are per ELEMENT, and with a runtime divisor they were most of the kernel's instructions.
template <typename T, int TAPS>
__device__ __forceinline__ void pack_tile(const PackJob& j, int t, T (*s)[32][34]) {
    const int tiles1 = (j.d1 + 31) / 32;
    const int o0 = (t / tiles1) * 32, i0 = (t % tiles1) * 32;
    const int n1 = min(32, j.d1 - i0);
    const int row = n1 * TAPS;                       // contiguous floats per outer index inside this tile
    if (n1 == 32) {
        for (int idx = threadIdx.x; idx < 32 * 32 * TAPS; idx += 256) {
            const int ol = idx / (32 * TAPS), r = idx - ol * (32 * TAPS);
            const int il = r / TAPS, tap = r - il * TAPS;
            if (o0 + ol < j.d0)
                s[tap][ol][il] = from_f<T>(j.w[(static_cast<long long>(o0 + ol) * j.d1 + i0) * TAPS + r]);
        }
    } else {
        for (int idx = threadIdx.x; idx < 32 * row; idx += 256) {
            const int ol = idx / row, r = idx - ol * row;
            const int il = r / TAPS, tap = r - il * TAPS;
            if (o0 + ol < j.d0)
                s[tap][ol][il] = from_f<T>(j.w[(static_cast<long long>(o0 + ol) * j.d1 + i0) * TAPS + r]);
        }
    }
    __syncthreads();
    T* wf = static_cast<T*>(j.wf);
    T* wd = static_cast<T*>(j.wd);
    for (int idx = threadIdx.x; idx < TAPS * 32 * 32; idx += 256) {
        const int fast = idx & 31, slow = (idx >> 5) & 31, tap = idx >> 10;
        if (TAPS == 9) {
            // conv: outer = co, inner = ci.  wf[co][tap][ci] (ci fastest), wd[ci][8-tap][co] (co fastest)
            if (o0 + slow < j.d0 && fast < n1)
                wf[(static_cast<long long>(o0 + slow) * 9 + tap) * j.d1 + i0 + fast] = s[tap][slow][fast];
            if (wd != nullptr && o0 + fast < j.d0 && slow < n1)
                wd[(static_cast<long long>(i0 + slow) * 9 + (8 - tap)) * j.d0 + o0 + fast] = s[tap][fast][slow];
        } else {
            // convT: outer = ci, inner = co.  wf[(tap,co)][ci] (ci fastest), wd[ci][(tap,co)] (co fastest)
            if (o0 + fast < j.d0 && slow < n1)
                wf[(static_cast<long long>(tap) * j.d1 + i0 + slow) * j.d0 + o0 + fast] = s[tap][fast][slow];
            if (o0 + slow < j.d0 && fast < n1)
                wd[static_cast<long long>(o0 + slow) * 4 * j.d1 + tap * j.d1 + i0 + fast] = s[tap][slow][fast];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_all_weights_kernel(const PackJobs jobs) {
    __shared__ T s[9][32][34];
    int li = 0;
    while (li + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.job[li + 1].tile_begin) ++li;
    const PackJob& j = jobs.job[li];
    const int t = blockIdx.x - j.tile_begin;
    if (j.taps == 9) pack_tile<T, 9>(j, t, s);
    else pack_tile<T, 4>(j, t, s);
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, no weight decay / amsgrad) over one flat fp32 parameter arena
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                            float grad_scale) {
    const float step = lr / bc1;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// Graph-capturable variant: the step count and the hyper-parameters live in device memory so that a captured
// launch stays valid while they change.  hyper = {lr, beta1, beta2, eps}; *step is the 1-based count of THIS step
// (adam_tick_kernel increments it right before).
__global__ void adam_tick_kernel(int* step) { *step += 1; }

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const float* __restrict__ hyper, const int* __restrict__ step, float grad_scale) {
    const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3];
    const float t = static_cast<float>(*step);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float stepsz = lr / bc1;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= stepsz * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// label = 1 iff Vd > Vt (argmax of the 2-way softmax, ties -> 0), Onet_vanilla_20240606.py:193-202
__global__ void predict_label_kernel(const float* __restrict__ Vt, const float* __restrict__ Vd, long long n,
                                     long long* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = Vd[i] > Vt[i] ? 1 : 0;
}

// The same decision as one byte per pixel, 16 pixels per thread: the mask the tiled inference path copies back to the host
// (8x fewer bytes than the reference's int64 argmax; widened on the host only when the caller asks for int64).
__global__ void __launch_bounds__(256)
predict_label_u8_kernel(const float* __restrict__ Vt, const float* __restrict__ Vd, long long n, unsigned char* __restrict__ out) {
    const long long n16 = n >> 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(Vt) | reinterpret_cast<uintptr_t>(Vd) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        for (long long i = t0; i < n16; i += stride) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(Vt) + 4 * i + k);
                const float4 b = __ldg(reinterpret_cast<const float4*>(Vd) + 4 * i + k);
                w[k] = (b.x > a.x ? 1u : 0u) | (b.y > a.y ? 1u << 8 : 0u) | (b.z > a.z ? 1u << 16 : 0u) | (b.w > a.w ? 1u << 24 : 0u);
            }
            reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (long long i = (n16 << 4) + t0; i < n; i += stride) out[i] = Vd[i] > Vt[i] ? 1 : 0;
    } else {
        for (long long i = t0; i < n; i += stride) out[i] = Vd[i] > Vt[i] ? 1 : 0;
    }
}

// Evaluation step next to the path (test_simclutter, Train_Onet_on_simclutter_20250407.py:109-147): predicted label
// (1 iff Vd > Vt, predict_label :193-202) against the ground-truth label, reduced on the device to the 2 x 2 confusion
// counts counts[pred * 2 + gt] from which utils_20231218.py's _acc / _miou / _detection_rate / _false_alarm_rate /
// _target_iou and re_assign_label all follow - one 32-byte read-back per batch instead of five reductions with .item().
__global__ void __launch_bounds__(256)
eval_confusion_kernel(const float* __restrict__ Vt, const float* __restrict__ Vd, const long long* __restrict__ gt, long long n,
                      unsigned long long* __restrict__ counts) {
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int pred = Vd[i] > Vt[i] ? 1 : 0;
        const int g = gt[i] != 0 ? 1 : 0;
        c[pred * 2 + g] += 1u;
    }
    __shared__ unsigned int s[4];
    if (threadIdx.x < 4) s[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned int v = c[k];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s[threadIdx.x]) atomicAdd(counts + threadIdx.x, static_cast<unsigned long long>(s[threadIdx.x]));
}

// ------------------------------------------------------------------------------------------------
// tensor_normal_per_frame (utils_20231218.py:673-689): every (image, channel) frame scaled to [0,1] by its own
// minimum and maximum, out = (v - min) / (max - min + eps) with eps = np.spacing(1) rounded to fp32 as torch does when
// it adds the Python float to an fp32 tensor.  Used on the stage-1 response maps that feed the second Onet of the
// two-stage cascade (Train_Onet_on_simclutter_20250407.py:296-390) and on datasets at load time.
// Pass 1: per-frame min / max as order-preserving integer keys (atomicMin / atomicMax); pass 2: scale.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void frame_minmax_init_kernel(int* __restrict__ keys, int frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < frames) {
        keys[2 * f] = 0x7fffffff;
        keys[2 * f + 1] = static_cast<int>(0x80000000u);
    }
}

__global__ void __launch_bounds__(256)
frame_minmax_kernel(const float* __restrict__ x, long long hw, int* __restrict__ keys) {
    const int f = blockIdx.y;
    const float* xf = x + static_cast<long long>(f) * hw;
    float mn = INFINITY, mx = -INFINITY;
    const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(xf) & 15) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(xf);
        for (long long i = start; i < (hw >> 2); i += stride) {
            const float4 v = x4[i];
            mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
            mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
    } else {
        for (long long i = start; i < hw; i += stride) {
            mn = fminf(mn, xf[i]);
            mx = fmaxf(mx, xf[i]);
        }
    }
    for (int off = 16; off >= 1; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    __shared__ float s_mn[8], s_mx[8];
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
        if (mn <= mx) {          // a block that saw no element keeps (+inf, -inf)
            atomicMin(keys + 2 * f, float_key(mn));
            atomicMax(keys + 2 * f + 1, float_key(mx));
        }
    }
}

__global__ void __launch_bounds__(256)
frame_normalize_kernel(const float* __restrict__ x, long long hw, const int* __restrict__ keys, float eps,
                       float* __restrict__ out) {
    const int f = blockIdx.y;
    const float mn = key_float(keys[2 * f]);
    const float den = (key_float(keys[2 * f + 1]) - mn) + eps;
    const float* xf = x + static_cast<long long>(f) * hw;
    float* of = out + static_cast<long long>(f) * hw;
    const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if ((hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(xf) | reinterpret_cast<uintptr_t>(of)) & 15) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(xf);
        float4* o4 = reinterpret_cast<float4*>(of);
        for (long long i = start; i < (hw >> 2); i += stride) {
            float4 v = x4[i];
            v.x = (v.x - mn) / den; v.y = (v.y - mn) / den; v.z = (v.z - mn) / den; v.w = (v.w - mn) / den;
            o4[i] = v;
        }
    } else {
        for (long long i = start; i < hw; i += stride) of[i] = (xf[i] - mn) / den;
    }
}

}  // namespace onet
