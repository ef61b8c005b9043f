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

// ------------------------------------------------------------------------------------------------
// Correlated K-distributed clutter field (K_distributed_SeaClutter_Simulation_20210919.py:469-526).  The two FFT pairs of
// the pipeline run through torch.fft (cuFFT) on the host side (onet_b200/synth.py); the element-wise stages are here, in
// double precision like the reference (the generator is offline work: fidelity first, and still ~10^3 x the CPU's 9 s/frame).
// ------------------------------------------------------------------------------------------------

// standard normal white noise (Box-Muller on Philox), one pair per counter value
__global__ void __launch_bounds__(256)
normal_fill_kernel(double* __restrict__ out, long long n, uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id) {
    const long long pairs = (n + 1) >> 1;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < pairs;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), stream_id, 0u, seed_lo, seed_hi);
        // 53-bit uniforms in (0, 1] and [0, 1)
        const double u1 = (static_cast<double>((static_cast<unsigned long long>(r.x) << 21) | (r.y >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
        const double u2 = static_cast<double>((static_cast<unsigned long long>(r.z) << 21) | (r.w >> 11)) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        out[2 * q] = rad * cs;
        if (2 * q + 1 < n) out[2 * q + 1] = rad * sn;
    }
}

// Memoryless non-linear transform mnlt(x, v) (:83-91): the Gamma(shape v, scale 1) quantile of Phi(x), integer v.
//   x <= 0: solve P(v, y) = p,      p = erfc(-x / sqrt 2) / 2,  P by its series  y^v e^-y / v! * (1 + y/(v+1) + ...)
//   x >  0: solve Q(v, y) = q,      q = erfc( x / sqrt 2) / 2,  Q = e^-y sum_{k<v} y^k / k!          (no cancellation in either tail)
// Halley iterations from the Wilson-Hilferty cube (or the leading term of the series in the far lower tail).
__device__ __forceinline__ double mnlt_int(double x, int v) {
    double fact = 1.0;                                   // (v-1)!
    for (int k = 2; k < v; ++k) fact *= k;
    const double c9 = 1.0 / (9.0 * v);
    const double t = 1.0 - c9 + x * sqrt(c9);
    const bool lower = x <= 0.0;
    const double target = 0.5 * erfc((lower ? -x : x) * 0.70710678118654752440);
    double y = v * t * t * t;
    if (lower) {
        const double y_tail = pow(target * fact * v, 1.0 / v);      // P ~ y^v / v!
        if (!(t > 0.0) || y < y_tail) y = y_tail;
    }
    if (!(y > 1e-300)) y = 1e-300;
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
        const double pdf = exp((v - 1) * log(y) - y) / fact;
        double F;                                         // monotone increasing in y in both branches
        if (lower) {
            double term = 1.0, sum = 1.0;
            for (int k = 1; k < 200; ++k) {
                term *= y / (v + k);
                sum += term;
                if (term < 1e-17 * sum) break;
            }
            F = pdf * y / v * sum - target;
        } else {
            double term = 1.0, sum = 1.0;
            for (int k = 1; k < v; ++k) { term *= y / k; sum += term; }
            F = target - exp(-y) * sum;
        }
        const double d = F / pdf;
        double step = d / (1.0 - 0.5 * d * ((v - 1) / y - 1.0));
        if (!(step == step) || fabs(step) > 0.75 * y) step = copysign(0.75 * y, d);      // stay positive, bounded relative change
        y -= step;
        if (fabs(step) <= 2e-15 * y) break;
    }
    return y;
}

__global__ void __launch_bounds__(256)
kfield_mnlt_kernel(const double* __restrict__ x, long long n, int v, double* __restrict__ y) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] = mnlt_int(x[i], v);
}

// coeff_acf_polyn (:121-139): per frame the three sums  S_n = sum exp(-x^2) H_n(x) g(x),  H_0 = 1, H_1 = 2x, H_2 = 4x^2 - 2.
// sums: [frames][3], zero-initialised; grid (blocks, frames)
__global__ void __launch_bounds__(256)
kfield_coeff_sums_kernel(const double* __restrict__ x, const double* __restrict__ g, long long per_frame, double* __restrict__ sums) {
    const int f = blockIdx.y;
    const double* xf = x + static_cast<long long>(f) * per_frame;
    const double* gf = g + static_cast<long long>(f) * per_frame;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_frame;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double xv = xf[i], e = exp(-xv * xv) * gf[i];
        s0 += e;
        s1 += e * (2.0 * xv);
        s2 += e * (4.0 * xv * xv - 2.0);
    }
    for (int off = 16; off >= 1; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    __shared__ double sh[3][8];
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; sh[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        atomicAdd(sums + 3 * f + threadIdx.x, t);
    }
}

// solve_acf_polyn (:141-164): out[f][i] = np.roots([a, b, 1 - acf[i]])[0] as (re, im), with [a, b, 1] = coeffs[f] (normalised):
// the root of larger magnitude when the roots are real, the one with positive imaginary part otherwise.
__global__ void __launch_bounds__(256)
kfield_acf_root_kernel(const double* __restrict__ coeffs /* [frames][2] = a, b */, const double* __restrict__ acf, long long per_frame,
                       double2* __restrict__ out) {
    const int f = blockIdx.y;
    const double a = coeffs[2 * f], b = coeffs[2 * f + 1];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_frame;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double c = 1.0 - acf[i];
        const double disc = b * b - 4.0 * a * c;
        const double s = sqrt(fabs(disc));
        double2 r;
        if (disc >= 0.0) { r.x = (-b - copysign(s, b)) / (2.0 * a); r.y = 0.0; }
        else { r.x = -b / (2.0 * a); r.y = s / (2.0 * a); }
        out[static_cast<long long>(f) * per_frame + i] = r;
    }
}

// amplitude = | speckle * sqrt(texture) |  (:517-518), complex speckle, fp32 out
__global__ void __launch_bounds__(256)
kfield_amplitude_kernel(const double2* __restrict__ speckle, const double* __restrict__ texture, long long n, float* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double2 z = speckle[i];
        out[i] = static_cast<float>(hypot(z.x, z.y) * sqrt(texture[i]));
    }
}

}  // namespace onet
