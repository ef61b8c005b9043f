// HBM-bound kernels of the Onet path: BatchNorm statistics finalisation, BN+ReLU(+2x2 max-pool) apply that
// writes straight into the skip-concat buffer, the matching backward passes (ReLU mask, max-pool routing and
// BatchNorm backward fused), the dot-product head + sigmoid + JSD loss forward/backward, input preparation,
// weight packing and Adam.  All activations NHWC; 8 channels (16 B of bf16 / 32 B of fp32) per thread access.
#pragma once
#include "simt_conv.cuh"

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
// BN + ReLU apply, optional fused 2x2 max-pool.  One thread = one 2x2 pixel quad x 8 channels.
//   y   : raw conv output [N,H,W,C]
//   out : [N,H,W,ldo] (+ooff)  (e.g. the skip half of a concat buffer)
//   pool: [N,H/2,W/2,C] or nullptr (floor semantics of nn.MaxPool2d(2))
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const T* __restrict__ y, int N, int H, int W, int C, const float* __restrict__ scale,
                     const float* __restrict__ shift, int group_images, T* __restrict__ out, long long ldo, int ooff,
                     T* __restrict__ pool, uint8_t* __restrict__ amax) {
    const int OC = C >> 3, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1, HP = H >> 1, WP = W >> 1;
    const long long total = static_cast<long long>(N) * H2 * W2 * OC;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int oc = static_cast<int>(idx % OC);
        long long q = idx / OC;
        const int w2 = static_cast<int>(q % W2); q /= W2;
        const int h2 = static_cast<int>(q % H2);
        const int n = static_cast<int>(q / H2);
        const int g = min(n / group_images, 1);
        float sc[8], sh[8], mx[8];
        uint32_t best[8];
        load8<float>(scale + g * C + oc * 8, sc);
        load8<float>(shift + g * C + oc * 8, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i) { mx[i] = -1.f; best[i] = 0; }   // post-ReLU values are >= 0 -> first pixel always wins first
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int h = 2 * h2 + dy, w = 2 * w2 + dx;
                if (h < H && w < W) {
                    const long long px = (static_cast<long long>(n) * H + h) * W + w;
                    float v[8];
                    load8<T>(y + px * C + oc * 8, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] = round_to<T>(fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f));
                        if (v[i] > mx[i]) { mx[i] = v[i]; best[i] = dy * 2 + dx; }   // first maximum, like ATen
                    }
                    store8<T>(out + px * ldo + ooff + oc * 8, v);
                }
            }
        if (pool != nullptr && h2 < HP && w2 < WP) {
            const long long po = ((static_cast<long long>(n) * HP + h2) * WP + w2) * C + oc * 8;
            store8<T>(pool + po, mx);
            if (amax != nullptr)
                *reinterpret_cast<uint2*>(amax + po) = make_uint2(best[0] | (best[1] << 8) | (best[2] << 16) | (best[3] << 24),
                                                                  best[4] | (best[5] << 8) | (best[6] << 16) | (best[7] << 24));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of (BN -> ReLU [-> skip / 2x2 max-pool]).  The gradient w.r.t. the post-ReLU activation is the
// sum of up to two dense sources (g1, g2: e.g. the skip half of d(concat) and the head's dL) and a pooled
// source gp (gradient of the max-pool output, routed to the first maximum of each 2x2 window exactly as
// ATen's max_pool2d backward does).  Pass 1 (reduce): s1 = sum dZ, s2 = sum dZ * xhat per group/channel.
// Pass 2 (apply): dY = gamma * invstd * (dZ - s1/n - xhat * s2/n).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct BnBwdArgs {
    const T* y; int N, H, W, C;
    const float* scale; const float* shift; const float* mean; const float* invstd;   // [G][C]
    int group_images;
    const T* g1; long long ld1; int off1;
    const T* g2; long long ld2; int off2;
    const T* gp;                                 // [N,H/2,W/2,C] or nullptr
    const uint8_t* amax;                         // [N,H/2,W/2,C] position (0..3) of the window maximum, with gp
    double* sums;                                // [G][2][C]
    double count;                                // elements per channel per group
    T* dy;                                       // [N,H,W,C]
};

// raw 8-channel vector (16 B of bf16 / 32 B of fp32) kept packed so that all loads of a work item can be issued
// before any of them is consumed
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

// Gradient w.r.t. the BN output (dz) and xhat of one pixel x 8 channels.
template <typename T>
__device__ __forceinline__ void bn_bwd_pixel(const Raw8<T>& ry, const Raw8<T>& rg1, const Raw8<T>& rg2, bool has_g2,
                                             const float (&sc)[8], const float (&sh)[8], const float (&mu)[8],
                                             const float (&is)[8], const float (&extra)[8], float (&dz)[8], float (&xh)[8]) {
    float y[8], g[8];
    unpack(ry, y);
    unpack(rg1, g);
    if (has_g2) {
        float g2[8];
        unpack(rg2, g2);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += g2[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float z = fmaf(y[i], sc[i], sh[i]);
        xh[i] = (y[i] - mu[i]) * is[i];
        dz[i] = (round_to<T>(fmaxf(z, 0.f)) > 0.f) ? g[i] + extra[i] : 0.f;
    }
}

template <typename T>
struct BnBwdCtx {      // per-thread channel constants
    float sc[8], sh[8], mu[8], is[8];
    __device__ __forceinline__ void load(const BnBwdArgs<T>& a, int g, int oc) {
        load8<float>(a.scale + g * a.C + oc * 8, sc);
        load8<float>(a.shift + g * a.C + oc * 8, sh);
        load8<float>(a.mean + g * a.C + oc * 8, mu);
        load8<float>(a.invstd + g * a.C + oc * 8, is);
    }
};

// block-wide reduction of the per-thread (acc1, acc2) over the pixel lanes and one double atomic per channel
template <typename T>
__device__ __forceinline__ void bn_bwd_block_reduce(const BnBwdArgs<T>& a, int g, const float (&acc1)[8],
                                                    const float (&acc2)[8], float (*s_red)[256]) {
    const int OC = a.C >> 3, LANES = 256 / OC;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s_red[i][threadIdx.x] = acc1[i];
        s_red[8 + i][threadIdx.x] = acc2[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * OC; t += 256) {
        const int which = t / OC, o = t % OC;
        float s = 0.f;
        for (int l = 0; l < LANES; ++l) s += s_red[which][l * OC + o];
        const int stat = which >> 3, i = which & 7;
        atomicAdd(a.sums + (static_cast<long long>(g) * 2 + stat) * a.C + o * 8 + i, static_cast<double>(s));
    }
}

// Pooled-gradient contribution for pixel q: gp[n,h/2,w/2,c] where this pixel is the window maximum.
template <typename T>
__device__ __forceinline__ void pooled_extra(const BnBwdArgs<T>& a, long long q, int oc, Raw8<T>& rp, uint2& am, int& pos) {
    const unsigned W = a.W, H = a.H;
    const unsigned qq = static_cast<unsigned>(q);
    const unsigned w = qq % W, t = qq / W;
    const unsigned h = t % H, n = t / H;
    const unsigned HP = H >> 1, WP = W >> 1;
    const unsigned h2 = h >> 1, w2 = w >> 1;
    pos = (h & 1) * 2 + (w & 1);
    if (h2 < HP && w2 < WP) {
        const long long po = ((static_cast<long long>(n) * HP + h2) * WP + w2) * a.C + oc * 8;
        rp = ldraw<T>(a.gp + po);
        am = __ldg(reinterpret_cast<const uint2*>(a.amax + po));
    } else {
        rp = zero_raw<T>();
        am = make_uint2(0xffffffffu, 0xffffffffu);
    }
}
__device__ __forceinline__ void select_extra(const float (&gpv)[8], uint2 am, int pos, float (&extra)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned b = ((i < 4 ? am.x : am.y) >> (8 * (i & 3))) & 0xffu;
        extra[i] = (b == static_cast<unsigned>(pos)) ? gpv[i] : 0.f;
    }
}

// ---- one thread = one pixel x 8 channels, UNROLL pixels in flight
template <typename T, int UNROLL, bool POOL>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_px_kernel(const BnBwdArgs<T> a) {
    __shared__ float s_red[16][256];
    const int OC = a.C >> 3, LANES = 256 / OC;
    const int oc = threadIdx.x % OC, ln = threadIdx.x / OC;
    const int g = blockIdx.y;
    const long long HW = static_cast<long long>(a.H) * a.W;
    const long long p_begin = static_cast<long long>(g) * a.group_images * HW;
    const long long p_end = static_cast<long long>(g == static_cast<int>(gridDim.y) - 1 ? a.N : (g + 1) * a.group_images) * HW;
    const bool has_g2 = a.g2 != nullptr;
    constexpr bool has_gp = POOL;
    float acc1[8] = {}, acc2[8] = {};
    if (ln < LANES) {
        BnBwdCtx<T> c;
        c.load(a, g, oc);
        const long long stride = static_cast<long long>(gridDim.x) * LANES;
        for (long long p = p_begin + blockIdx.x * static_cast<long long>(LANES) + ln; p < p_end; p += stride * UNROLL) {
            Raw8<T> ry[UNROLL], rg1[UNROLL], rg2[UNROLL], rp[UNROLL];
            uint2 am[UNROLL];
            int pos[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long q = p + u * stride;
                const bool ok = q < p_end;
                ry[u] = ok ? ldraw<T>(a.y + q * a.C + oc * 8) : zero_raw<T>();
                rg1[u] = ok ? ldraw<T>(a.g1 + q * a.ld1 + a.off1 + oc * 8) : zero_raw<T>();
                rg2[u] = (ok && has_g2) ? ldraw<T>(a.g2 + q * a.ld2 + a.off2 + oc * 8) : zero_raw<T>();
                if (ok && has_gp) pooled_extra<T>(a, q, oc, rp[u], am[u], pos[u]);
                else { rp[u] = zero_raw<T>(); am[u] = make_uint2(0xffffffffu, 0xffffffffu); pos[u] = 0; }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float dz[8], xh[8], gpv[8], extra[8] = {};
                if (has_gp) {
                    unpack(rp[u], gpv);
                    select_extra(gpv, am[u], pos[u], extra);
                }
                bn_bwd_pixel<T>(ry[u], rg1[u], rg2[u], has_g2, c.sc, c.sh, c.mu, c.is, extra, dz, xh);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc1[i] += dz[i];
                    acc2[i] = fmaf(dz[i], xh[i], acc2[i]);
                }
            }
        }
    }
    bn_bwd_block_reduce<T>(a, g, acc1, acc2, s_red);
}

template <typename T, int UNROLL, bool POOL>
__global__ void __launch_bounds__(256)
bn_bwd_apply_px_kernel(const BnBwdArgs<T> a) {
    const int OC = a.C >> 3, LANES = 256 / OC;
    const int oc = threadIdx.x % OC, ln = threadIdx.x / OC;
    const int g = blockIdx.y;
    if (ln >= LANES) return;
    const long long HW = static_cast<long long>(a.H) * a.W;
    const long long p_begin = static_cast<long long>(g) * a.group_images * HW;
    const long long p_end = static_cast<long long>(g == static_cast<int>(gridDim.y) - 1 ? a.N : (g + 1) * a.group_images) * HW;
    const bool has_g2 = a.g2 != nullptr;
    constexpr bool has_gp = POOL;
    const float inv_n = static_cast<float>(1.0 / a.count);
    BnBwdCtx<T> c;
    c.load(a, g, oc);
    float m1[8], m2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m1[i] = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 0) * a.C + oc * 8 + i]) * inv_n;
        m2[i] = static_cast<float>(a.sums[(static_cast<long long>(g) * 2 + 1) * a.C + oc * 8 + i]) * inv_n;
    }
    const long long stride = static_cast<long long>(gridDim.x) * LANES;
    for (long long p = p_begin + blockIdx.x * static_cast<long long>(LANES) + ln; p < p_end; p += stride * UNROLL) {
        Raw8<T> ry[UNROLL], rg1[UNROLL], rg2[UNROLL], rp[UNROLL];
        uint2 am[UNROLL];
        int pos[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long q = p + u * stride;
            const bool ok = q < p_end;
            ry[u] = ok ? ldraw<T>(a.y + q * a.C + oc * 8) : zero_raw<T>();
            rg1[u] = ok ? ldraw<T>(a.g1 + q * a.ld1 + a.off1 + oc * 8) : zero_raw<T>();
            rg2[u] = (ok && has_g2) ? ldraw<T>(a.g2 + q * a.ld2 + a.off2 + oc * 8) : zero_raw<T>();
            if (ok && has_gp) pooled_extra<T>(a, q, oc, rp[u], am[u], pos[u]);
            else { rp[u] = zero_raw<T>(); am[u] = make_uint2(0xffffffffu, 0xffffffffu); pos[u] = 0; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long q = p + u * stride;
            if (q >= p_end) break;
            float dz[8], xh[8], o[8], gpv[8], extra[8] = {};
            if (has_gp) {
                unpack(rp[u], gpv);
                select_extra(gpv, am[u], pos[u], extra);
            }
            bn_bwd_pixel<T>(ry[u], rg1[u], rg2[u], has_g2, c.sc, c.sh, c.mu, c.is, extra, dz, xh);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = c.sc[i] * (dz[i] - m1[i] - xh[i] * m2[i]);
            store8<T>(a.dy + q * a.C + oc * 8, o);
        }
    }
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
};

template <typename T>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const HeadArgs<T> a) {
    const long long npx = static_cast<long long>(a.B) * a.HW;
    const int sub = threadIdx.x & 7;
    float lsum = 0.f;
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npx;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const long long pd = p + npx;   // same pixel of the down-branch image
        float lt[8], ht[8], ld[8], hd[8];
        load8<T>(a.L + p * a.ldl + a.offl + sub * 8, lt);
        load8<T>(a.Hf + p * a.ldh + a.offh + sub * 8, ht);
        load8<T>(a.L + pd * a.ldl + a.offl + sub * 8, ld);
        load8<T>(a.Hf + pd * a.ldh + a.offh + sub * 8, hd);
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
            vt += __shfl_xor_sync(0xffffffffu, vt, o);
            vd += __shfl_xor_sync(0xffffffffu, vd, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (sub == 0) {
            const float mx = fmaxf(vt, vd);
            const float et = expf(vt - mx), ed = expf(vd - mx);
            const float inv = 1.f / (et + ed);
            const float st = et * inv, sd = ed * inv;
            const long long n = p / a.HW, hw = p % a.HW;
            a.Vt[p] = vt;
            a.Vd[p] = vd;
            a.S[(n * 2 + 0) * a.HW + hw] = st;
            a.S[(n * 2 + 1) * a.HW + hw] = sd;
            a.a[p] = sa;
            a.b[p] = sb;
            float v1, v2, v3, v4, d;
            sp_ref(-sa * st, v1, d);
            sp_ref(sa * sd, v2, d);
            sp_ref(-sb * sd, v3, d);
            sp_ref(sb * st, v4, d);
            lsum += (v1 + v2) + (v3 + v4);
        }
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

// label = 1 iff Vd > Vt (argmax of the 2-way softmax, ties -> 0), Onet_vanilla_20240606.py:193-202
__global__ void predict_label_kernel(const float* __restrict__ Vt, const float* __restrict__ Vd, long long n,
                                     long long* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = Vd[i] > Vt[i] ? 1 : 0;
}

}  // namespace onet
