// Head forward (Onet.forward, Onet_vanilla_20240606.py:176-189, + compute_loss :221-267) for bf16 storage on warp-level MMAs.
//
// head_fwd_kernel (elementwise.cuh) spends ~450 warp instructions per 4 pixels - unpacking 4 x 64 bf16 values, the BatchNorm +
// ReLU + rounding of the two global-feature rows, 128 FMAs for the channel dot products, shuffles - on 2 KB of input: it runs at
// 2.3x its HBM time.  Here a warp owns 16 pixels:  V[p] = sum_c L[p][c] H[p][c]  is the DIAGONAL of the 16 x 16 product L H^T, and
// a = sum_c L[p][c] is L times a ones matrix; both are mma.sync.m16n8k16 on the packed bf16 words exactly as they lie in memory
// (products of bf16 values are exact in fp32, the accumulation is fp32: same arithmetic as the FMA loop up to the order of the
// additions).  15/16 of the MMA work is thrown away, which is still 8 x fewer issue slots than the FMAs + unpacking it replaces;
// the fused BatchNorm + ReLU of H packs its result straight into the B-operand words.
//
// Fragment <-> channel mapping: the contraction slot (k-step ks, column 2q + j [+ 8]) is channel 16 q + 4 ks + j [+ 2], so a thread
// (r = lane / 4, q = lane % 4) loads for the pixels r and r + 8 the 32 contiguous bytes of channels [16 q, 16 q + 16) of each
// tensor - two 16-byte loads - and uses the words unchanged as A fragments (L) or B fragments (H).
#pragma once
#include "tc_common.cuh"
#include "elementwise.cuh"
#include "first_layer_mma.cuh"

namespace onet {

__device__ __forceinline__ float sp_val(float x) {          // value branch of sp_ref
    if (x <= -37.f) return logf(1.f + expf(expf(x)));
    if (x <= 18.f) return logf(1.f + expf(x));
    if (x < 33.3f) return x + expf(-x);
    return x;
}

__global__ void __launch_bounds__(256, 2)
head_fwd_mma_kernel(const HeadArgs<__nv_bfloat16> a, const FastDiv fd_hw) {
    __shared__ __align__(16) float s_aff[4][64];       // hsc_t, hsh_t, hsc_d, hsh_d
    __shared__ float s_l[8];
    const bool fused = a.hsc_t != nullptr;
    if (fused) {
        const int c = threadIdx.x & 63, which = threadIdx.x >> 6;
        const float* src = which == 0 ? a.hsc_t : (which == 1 ? a.hsh_t : (which == 2 ? a.hsc_d : a.hsh_d));
        s_aff[which][c] = src[c];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int npx = a.B * static_cast<int>(a.HW);
    const int HW = static_cast<int>(a.HW);
    const int pj = lane & 15, hf = lane >> 4;                       // tail: lane = (pixel of the tile, branch half)
    const int src = (pj & 7) * 4 + ((pj & 7) >> 1);                 // lane that holds the diagonal element of pixels pj & 7 and (pj & 7) + 8
    const bool odd = (r & 1) != 0;
    float lsum = 0.f;
    const int ntiles = (npx + 15) >> 4;
    for (int tile = blockIdx.x * 8 + warp; tile < ntiles; tile += gridDim.x * 8) {
        const int p0 = tile << 4;
        uint4 Lr[2][2][2], Hr[2][2][2];            // [branch][row: pixel r / r + 8][16-byte half]
#pragma unroll
        for (int br = 0; br < 2; ++br)
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                const int p = p0 + r + 8 * row;
                const bool ok = p < npx;
                const long long px = static_cast<long long>(p) + (br ? npx : 0);
                const uint4* lp = reinterpret_cast<const uint4*>(a.L + px * a.ldl + a.offl + q * 16);
                const uint4* hp = reinterpret_cast<const uint4*>(a.Hf + px * a.ldh + a.offh + q * 16);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    Lr[br][row][h2] = ok ? __ldg(lp + h2) : make_uint4(0u, 0u, 0u, 0u);
                    Hr[br][row][h2] = ok ? __ldg(hp + h2) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
        float dv[2][2][4], ds[2][4];               // dv[branch][n-tile]: L H^T;  ds[branch]: L . ones
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            uint32_t Lw[2][8], Hw[2][8];
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                const uint4 l0 = Lr[br][row][0], l1 = Lr[br][row][1], h0 = Hr[br][row][0], h1 = Hr[br][row][1];
                Lw[row][0] = l0.x; Lw[row][1] = l0.y; Lw[row][2] = l0.z; Lw[row][3] = l0.w;
                Lw[row][4] = l1.x; Lw[row][5] = l1.y; Lw[row][6] = l1.z; Lw[row][7] = l1.w;
                Hw[row][0] = h0.x; Hw[row][1] = h0.y; Hw[row][2] = h0.z; Hw[row][3] = h0.w;
                Hw[row][4] = h1.x; Hw[row][5] = h1.y; Hw[row][6] = h1.z; Hw[row][7] = h1.w;
            }
            if (fused) {        // Hf is the last layer's raw conv output: h = round_bf16(relu(y * scale + shift)), packed in place
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {        // words 2 w4, 2 w4 + 1 = channels 16 q + 4 w4 .. + 3
                    const float4 sc = *reinterpret_cast<const float4*>(&s_aff[2 * br][q * 16 + 4 * w4]);
                    const float4 sh = *reinterpret_cast<const float4*>(&s_aff[2 * br + 1][q * 16 + 4 * w4]);
#pragma unroll
                    for (int row = 0; row < 2; ++row) {
                        const uint32_t w0 = Hw[row][2 * w4], w1 = Hw[row][2 * w4 + 1];
                        const float y0 = __uint_as_float(w0 << 16), y1 = __uint_as_float(w0 & 0xffff0000u);
                        const float y2 = __uint_as_float(w1 << 16), y3 = __uint_as_float(w1 & 0xffff0000u);
                        Hw[row][2 * w4] = pack_bf16x2(fmaxf(fmaf(y0, sc.x, sh.x), 0.f), fmaxf(fmaf(y1, sc.y, sh.y), 0.f));
                        Hw[row][2 * w4 + 1] = pack_bf16x2(fmaxf(fmaf(y2, sc.z, sh.z), 0.f), fmaxf(fmaf(y3, sc.w, sh.w), 0.f));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { dv[br][0][i] = 0.f; dv[br][1][i] = 0.f; ds[br][i] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t af[4] = {Lw[0][2 * ks], Lw[1][2 * ks], Lw[0][2 * ks + 1], Lw[1][2 * ks + 1]};
                mma_bf16_16816(dv[br][0], af, Hw[0][2 * ks], Hw[0][2 * ks + 1]);      // columns = pixels 0..7
                mma_bf16_16816(dv[br][1], af, Hw[1][2 * ks], Hw[1][2 * ks + 1]);      // columns = pixels 8..15
                mma_bf16_16816(ds[br], af, 0x3f803f80u, 0x3f803f80u);                // every column = channel sum
            }
        }
        // diagonal: D[r][r] is column r = 2 (r/2) + (r & 1) of row r -> thread (r, q = r / 2), register r & 1 (n-tile 0) and
        // 2 + (r & 1) (row r + 8 of n-tile 1).  Hand pixel pj of the tile to lanes pj (top-branch terms) and 16 + pj (down).
        const float vt_lo = odd ? dv[0][0][1] : dv[0][0][0], vt_hi = odd ? dv[0][1][3] : dv[0][1][2];
        const float vd_lo = odd ? dv[1][0][1] : dv[1][0][0], vd_hi = odd ? dv[1][1][3] : dv[1][1][2];
        const float t0 = __shfl_sync(0xffffffffu, vt_lo, src), t1 = __shfl_sync(0xffffffffu, vt_hi, src);
        const float d0 = __shfl_sync(0xffffffffu, vd_lo, src), d1 = __shfl_sync(0xffffffffu, vd_hi, src);
        const float a0 = __shfl_sync(0xffffffffu, ds[0][0], src), a1 = __shfl_sync(0xffffffffu, ds[0][2], src);
        const float b0 = __shfl_sync(0xffffffffu, ds[1][0], src), b1 = __shfl_sync(0xffffffffu, ds[1][2], src);
        const bool hi = pj >= 8;
        const float vt = hi ? t1 : t0, vd = hi ? d1 : d0, sa = hi ? a1 : a0, sb = hi ? b1 : b0;
        const float mx = fmaxf(vt, vd);
        const float et = expf(vt - mx), ed = expf(vd - mx);
        const float inv = 1.f / (et + ed);
        const float st = et * inv, sd = ed * inv;
        const int p = p0 + pj;
        if (p < npx) {
            // half 0: the two Lt terms, half 1: the two Ld terms of the Jensen-Shannon estimate
            lsum += hf == 0 ? sp_val(-sa * st) + sp_val(sa * sd) : sp_val(-sb * sd) + sp_val(sb * st);
            const int n = fd_div(p, fd_hw), hw = p - n * HW;
            if (hf == 0) {
                a.Vt[p] = vt;
                a.S[(static_cast<long long>(n) * 2 + 0) * HW + hw] = st;
                a.a[p] = sa;
            } else {
                a.Vd[p] = vd;
                a.S[(static_cast<long long>(n) * 2 + 1) * HW + hw] = sd;
                a.b[p] = sb;
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) s_l[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0 && a.loss_acc != nullptr) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_l[w];
        atomicAdd(a.loss_acc, static_cast<double>(t));
    }
}

}  // namespace onet
