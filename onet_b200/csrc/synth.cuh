// On-device synthesis of the training frames: the clutter background and the Gaussian extended
// targets of Rayleigh_bg_Gaussian_EOT_generator_20230208.py, so that data never limits multi-GPU throughput
// (the reference's CPU generators take 0.1 - 9 s per frame).
//
//   rayleigh_fill_kernel   : Rayleigh(sigma) amplitudes, get_rayleigh_frame :219-221 (scipy.stats.rayleigh.rvs)
//   kclutter_fill_kernel   : compound-Gaussian K-distributed amplitudes, Rayleigh speckle x sqrt(Gamma(nu, 1/nu)) texture:
//                            the marginal distribution the reference's K generator targets
//                            (K_distributed_SeaClutter_Simulation_20210919.py:469-526) WITHOUT its spatial correlation
//   add_targets_kernel     : add_gaussian_template_on_clutter_v3 :62-176 (swerling type 0), the targets of one frame
//                            composited one after the other exactly in the reference's order
//
// Random numbers: Philox4x32-10 keyed by (seed, stream), counter = element index / 4; every element is a pure function
// of (seed, stream, index), independent of the launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace onet {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// uniform in (0, 1]: never 0, so log() is finite
__device__ __forceinline__ float u01(uint32_t r) { return (static_cast<float>(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256)
rayleigh_fill_kernel(float* __restrict__ out, long long n, float sigma, uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id) {
    const long long quads = (n + 3) >> 2;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < quads;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), stream_id, 0u, seed_lo, seed_hi);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = sigma * sqrtf(-2.0f * logf(u01(rr[j])));
        if (4 * q + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            reinterpret_cast<float4*>(out)[q] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * q + j < n) out[4 * q + j] = v[j];
        }
    }
}

// amplitude = Rayleigh(1) speckle * sqrt(texture), texture ~ Gamma(shape nu, scale 1/nu) as the mean of nu unit exponentials
__global__ void __launch_bounds__(256)
kclutter_fill_kernel(float* __restrict__ out, long long n, int nu, uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float speckle = 0.f, tex = 0.f;
        int drawn = 0;
        for (uint32_t blk = 0; drawn < nu + 1; ++blk) {
            const Philox4 r = philox4x32_10(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), stream_id, blk, seed_lo, seed_hi);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (drawn == 0) speckle = sqrtf(-2.0f * logf(u01(rr[j])));
                else if (drawn <= nu) tex -= logf(u01(rr[j]));
                ++drawn;
            }
        }
        out[i] = speckle * sqrtf(tex / static_cast<float>(nu));
    }
}

// One target of one frame, precomputed on the host from (cx, cy, w, h, theta) exactly as the reference derives them.
struct SynthTarget {
    int lx, ly;            // top-left corner of the template window in the frame      (:76-79)
    int wr, hr;            // half width / half height of the window: (2hr+1) x (2wr+1) (gaussian_kernel2d :36-37)
    float a, b, c;         // quadratic form of the rotated Gaussian                    (:47-49)
    float thr;             // mask threshold kgauss.max() - 2 * kgauss.std() (:142); negative = reduce it on the device
};

// One CTA per frame.  erc = mean(bg^2) of the untouched background (:221 / :189), kcoef = sqrt(10^(snr/10) * erc) (:90),
// then for every target in order: template = kgauss * kcoef, bg += (template > bg) * template, mask |= kgauss > thr.
__global__ void __launch_bounds__(256)
add_targets_kernel(float* __restrict__ frames, unsigned char* __restrict__ masks, int H, int W,
                   const SynthTarget* __restrict__ targets, int targets_per_frame, float snr_gain /* 10^(snr/20) */,
                   float* __restrict__ erc_out) {
    const int f = blockIdx.x;
    float* bg = frames + static_cast<long long>(f) * H * W;
    unsigned char* mk = masks + static_cast<long long>(f) * H * W;
    __shared__ double s_part[8], s_part2[8];
    __shared__ float s_kcoef, s_thr;
    double acc = 0.0;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const float v = bg[i];
        acc += static_cast<double>(v) * v;
    }
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[w];
        const float erc = static_cast<float>(t / (static_cast<double>(H) * W));
        if (erc_out != nullptr) erc_out[f] = erc;
        s_kcoef = snr_gain * sqrtf(erc);
    }
    __syncthreads();
    const float kcoef = s_kcoef;
    for (int t = 0; t < targets_per_frame; ++t) {
        const SynthTarget tg = targets[static_cast<long long>(f) * targets_per_frame + t];
        const int wt = 2 * tg.wr + 1, ht = 2 * tg.hr + 1;
        float thr = tg.thr;
        if (thr < 0.f) {          // population standard deviation of the template over its window; its maximum is kgauss(0,0) = 1
            double s1 = 0.0, s2 = 0.0;
            for (int i = threadIdx.x; i < wt * ht; i += blockDim.x) {
                const float fx = static_cast<float>(i % wt - tg.wr), fy = static_cast<float>(i / wt - tg.hr);
                const double kg = static_cast<double>(expf(-(tg.a * fx * fx + 2.0f * tg.b * fx * fy + tg.c * fy * fy)));
                s1 += kg;
                s2 += kg * kg;
            }
            for (int off = 16; off >= 1; off >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            }
            if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = s1; s_part2[threadIdx.x >> 5] = s2; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double t1 = 0.0, t2 = 0.0;
                for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { t1 += s_part[w]; t2 += s_part2[w]; }
                const double m1 = t1 / (wt * ht), m2 = t2 / (wt * ht);
                s_thr = static_cast<float>(1.0 - 2.0 * sqrt(fmax(m2 - m1 * m1, 0.0)));
            }
            __syncthreads();
            thr = s_thr;
        }
        for (int i = threadIdx.x; i < wt * ht; i += blockDim.x) {
            const int ky = i / wt - tg.hr, kx = i % wt - tg.wr;
            const int y = tg.ly + ky + tg.hr, x = tg.lx + kx + tg.wr;
            if (y < 0 || y >= H || x < 0 || x >= W) continue;        // the host rejects such targets like the reference does
            const float fx = static_cast<float>(kx), fy = static_cast<float>(ky);
            const float kg = expf(-(tg.a * fx * fx + 2.0f * tg.b * fx * fy + tg.c * fy * fy));
            const float tmpl = kg * kcoef;
            const long long p = static_cast<long long>(y) * W + x;
            const float bk = bg[p];
            bg[p] = bk + (tmpl > bk ? tmpl : 0.0f);
            if (kg > thr) mk[p] = 1;
        }
        __syncthreads();      // the next target may overlap this one
    }
}

}  // namespace onet
