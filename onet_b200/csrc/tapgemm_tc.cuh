// tcgen05 / TMEM / TMA "tap-GEMM" kernels for sm_100a: the tensor-core engine behind the 3x3 convolutions
// (fwd, dgrad, wgrad) and the 2x2/stride-2 transposed convolutions (fwd, dgrad, wgrad) of the Onet U-Nets.
//
// Activations are NHWC, bf16 (OpBf16) or fp32 read as TF32 (OpTf32, tc_common.cuh).  Every operand tile is fetched by TMA from a 5-D view (c, w, q, h, n) of an
// activation tensor with the 128-byte swizzle, so a "tap" (a filter offset) is nothing but a coordinate
// offset of the box; out-of-bounds rows are zero-filled by the TMA unit, which implements the conv padding.
//
//   pixel-major kernel (tapgemm_px_kernel):   D[pixel][cout] = sum_taps sum_cin In[pixel (+) tap][cin] * Wt[cout][tap][cin]
//       A = 128 pixels x 64 channels (K-major), B = BN couts x 64 channels (K-major), fp32 accumulators in TMEM.
//       Epilogues: raw conv output (bf16) + per-channel BatchNorm partial sums, or the 2x2 pixel-shuffle
//       scatter (+bias) of the transposed convolution straight into the skip-concat buffer.
//   weight-gradient kernel (tapgemm_wg_kernel): dW[m][n][tap] = sum_pixels G[pixel (-) tap][m] * In[pixel][n]
//       both operands MN-major (pixels are the K dimension), split-K over pixel slabs, fp32 atomics.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer, then the epilogue warps (pixel-major
// kernels: 8, two per TMEM lane quarter splitting the column chunks; weight-gradient kernels: 4).  Persistent CTAs, static round-robin tile schedule,
// smem ring of STAGES operand stages, double-buffered TMEM accumulators (pixel-major kernel).
#pragma once
#include "tc_common.cuh"
#include "fastdiv.cuh"

namespace onet {

constexpr int kMaxTaps = 9;

enum EpiMode : int { EPI_STORE = 0, EPI_CONVT = 1 };
// (EPI_STORE with PxParams::scale != nullptr = inference: BatchNorm(eval) + ReLU folded into the store)

struct PxParams {
    // output pixel grid and its tiling (all tile dims are powers of two, TW*TH*TN <= 128)
    int N, H, W;
    int TW, TH, TN, log_tw, log_th;
    int tiles_w, tiles_h, tiles_n;
    int num_m_tiles, num_n_tiles;
    FastDiv fd_w, fd_wh, fd_m, fd_pm;   // / tiles_w, / (tiles_w * tiles_h), / num_m_tiles, / ceil(num_m_tiles / 2); px_set_fastdiv()
    int valid_rows;            // TW*TH*TN
    int ntaps, k_chunks, cin;  // K = ntaps * cin, cin = Op::kKC * k_chunks (one K chunk = 128 bytes of channels)
    int4 taps[kMaxTaps];       // coordinate offsets (dc, dw, dq, dh) of each tap in the 5-D input view
    int tap_w[kMaxTaps];       // index of each tap in the packed weight tensor (K coordinate = tap_w * cin + channel)
    // epilogue
    int epi_mode;
    void* out;                 // EPI_STORE: [N,H,W,ldo]; EPI_CONVT: [N,2H,2W,ldo]; elements of Op::T
    long long ldo;             // channels per pixel of the output buffer
    int out_coff;              // first output channel inside the buffer
    double* stat_sum;          // [groups][cout_total] or nullptr
    double* stat_sq;
    int cout_total;
    int group_images;          // images per BatchNorm statistics group (twin branch)
    const float* scale;        // EPI_STORE, inference: per-group per-channel BatchNorm(eval) scale / shift [G][cout_total];
    const float* shift;        //   the epilogue then stores relu(acc * scale + shift) instead of the raw conv output
    // EPI_STORE, backward: the conv output is the gradient g w.r.t. the post-ReLU activation of the PREVIOUS layer; with
    // red_y set the epilogue reduces that layer's BatchNorm-backward sums instead of the output statistics:
    //   stat_sum[grp][c] += sum dz,  stat_sq[grp][c] += sum dz * (y - mean) * invstd,  dz = (bn_relu(y) > 0) ? g : 0
    // (the reduce pass of bn_bwd_px_kernel, elementwise.cuh, without re-reading g from HBM; kernels instantiated with RED)
    const __nv_bfloat16* red_y;   // [N,H,W,cout_total] raw conv output of the previous layer, or nullptr
    const float* red_scale;       // [G][cout_total]
    const float* red_shift;
    const float* red_mean;
    const float* red_invstd;
    int stat_gstride;             // doubles between the statistics groups of stat_sum / stat_sq (cout_total, or 2 * cout_total)
    const float* bias;         // EPI_CONVT: [co_per_tap]
    int co_per_tap;            // EPI_CONVT: output channels per 2x2 position
    int Ho, Wo;                // EPI_CONVT: height / width of the fine grid the buffer holds (>= 2H, 2W; F.pad border beyond)
};

inline void px_set_fastdiv(PxParams& p) {
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_wh = make_fastdiv(p.tiles_w * p.tiles_h);
    p.fd_m = make_fastdiv(p.num_m_tiles);
    p.fd_pm = make_fastdiv((p.num_m_tiles + 1) >> 1);
}

// Output pixel of an epilogue thread (TMEM lane quarter q, lane) in pixel tile m_tile.  Computed BEFORE the wait for the tile's
// accumulator so that it is off the accumulator-drain path.
struct PxRowCoord {
    int n, h, w, nt;
    bool valid;
};
__device__ __forceinline__ void px_tile_coord(const PxParams& p, int m_tile, int& wt, int& ht, int& nt) {
    nt = fd_div(m_tile, p.fd_wh);
    const int rem = m_tile - nt * (p.tiles_w * p.tiles_h);
    ht = fd_div(rem, p.fd_w);
    wt = rem - ht * p.tiles_w;
}
__device__ __forceinline__ PxRowCoord px_row_coord(const PxParams& p, int m_tile, int q, int lane) {
    const int row = q * 32 + lane;
    const int w_l = row & (p.TW - 1), h_l = (row >> p.log_tw) & (p.TH - 1), n_l = row >> (p.log_tw + p.log_th);
    int wt, ht, nt;
    px_tile_coord(p, m_tile, wt, ht, nt);
    PxRowCoord c;
    c.w = wt * p.TW + w_l; c.h = ht * p.TH + h_l; c.n = nt * p.TN + n_l; c.nt = nt;
    c.valid = (row < p.valid_rows) && (c.w < p.W) && (c.h < p.H) && (c.n < p.N);
    return c;
}
// statistics group of pixel tile nt (at most two groups: the twin branches; a tile never straddles them)
__device__ __forceinline__ int px_group(const PxParams& p, int nt) { return (nt * p.TN >= p.group_images) ? 1 : 0; }

template <int BN>
struct PxCfg {
    static constexpr int kABytes = 128 * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : 2 * BN;   // two accumulators
    static constexpr int kAuxBytes = 1024 + 4 * 2 * BN * 4;           // barriers + per-warp stat partials
    static constexpr int kSmemBytes = kStages * kStageBytes + kAuxBytes + 1024;  // + alignment slack
};

// Per-thread running BatchNorm statistics of the persistent CTA: epilogue thread (ew, lane) owns output column
// c = ew * 32 + lane of the current n-tile (BN <= 32 * kPxEpiWarps) for both statistics groups.  The sums are flushed
// with fp64 atomics only when the n-tile changes and at the end of the kernel - not once per tile: with 65536 tiles
// hitting the same 2 x 64 addresses the per-tile atomics were serialised in L2.
struct PxStatAcc {
    double s0, q0, s1, q1;
    int n_tile;
    // Per-ROW (per-thread) running sums of this warp's columns, folded across lanes only when the statistics group changes and
    // at the end: a per-tile shuffle butterfly costs 124 warp shuffles per 32-column chunk, and shuffles go through the
    // shared-memory crossbar that the MMAs and TMA saturate (they made the forward 64 -> 128 layer 53 % slower than its dgrad
    // form).  BN = 64: one chunk per warp, 32 columns per thread.  BN = 128 / 256 (2 / 4 chunks per warp; bf16): before
    // accumulating, lane pairs (quads) swap halves (quarters) of the PACKED bf16 words of a chunk - 8 (16) shuffles - so a thread
    // holds 2 (4) rows x 16 (8) columns and 32 accumulators per statistic cover all of the warp's chunks:
    //   rs / rq [cc * (32 >> STEPS) + i]  =  chunk cc, column 16 (lane & 1) + 8 ((lane >> 1) & 1) [STEPS = 2] + i
    float rs[32], rq[32];
    int grp, pending;
    __device__ __forceinline__ void reset(int nt) { s0 = q0 = s1 = q1 = 0.0; n_tile = nt; }
    __device__ __forceinline__ void reset_rows() {
#pragma unroll
        for (int j = 0; j < 32; ++j) { rs[j] = 0.f; rq[j] = 0.f; }
        pending = 0;
    }
};

template <int BN> struct PxRowAcc {
    // lane-exchange steps; -1: no row form.  (256-column tiles could use 2 steps - px_rowacc_packed<2, .> - but the four-way
    // chunk dispatch spills in the 168-register kernels, and their statistics cost 3 % at most: they keep the per-tile butterfly.)
    static constexpr int kSteps = BN == 64 ? 0 : (BN == 128 ? 1 : -1);
    static constexpr int kChunks = BN >= 64 ? BN / 64 : 1;                                   // 32-column chunks per epilogue warp
    static constexpr int kVals = kSteps >= 0 ? (32 >> kSteps) : 32;                          // running sums per chunk and statistic
};

__device__ __forceinline__ void px_acc_word(uint32_t w, float& s_lo, float& s_hi, float& q_lo, float& q_hi) {
    const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
    s_lo += lo; s_hi += hi;
    q_lo = fmaf(lo, lo, q_lo); q_hi = fmaf(hi, hi, q_hi);
}

// Chunk CC of a warp: pk = the 16 packed bf16 words (32 columns) of this thread's row, as stored.
template <int STEPS, int CC>
__device__ __forceinline__ void px_rowacc_packed(const uint32_t (&pk)[16], bool valid, int lane, PxStatAcc& a) {
    constexpr int V = 32 >> STEPS;
    uint32_t k[8], r[8];
    const bool b1 = (lane & 1) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t lo = valid ? pk[j] : 0u, hi = valid ? pk[8 + j] : 0u;
        k[j] = b1 ? hi : lo;
        r[j] = __shfl_xor_sync(0xffffffffu, b1 ? lo : hi, 1);
    }
    if constexpr (STEPS == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            px_acc_word(k[j], a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
            px_acc_word(r[j], a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
        }
    } else {
        const bool b2 = (lane & 2) != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k2 = b2 ? k[4 + j] : k[j], r2 = b2 ? r[4 + j] : r[j];
            const uint32_t k3 = __shfl_xor_sync(0xffffffffu, b2 ? k[j] : k[4 + j], 2);
            const uint32_t r3 = __shfl_xor_sync(0xffffffffu, b2 ? r[j] : r[4 + j], 2);
            px_acc_word(k2, a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
            px_acc_word(r2, a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
            px_acc_word(k3, a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
            px_acc_word(r3, a.rs[CC * V + 2 * j], a.rs[CC * V + 2 * j + 1], a.rq[CC * V + 2 * j], a.rq[CC * V + 2 * j + 1]);
        }
    }
}

// transposing butterfly: on return lane j holds in v[0] / s2[0] the sum over the warp's 32 lanes of element j
__device__ __forceinline__ void px_butterfly(float (&v)[32], float (&s2)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            const float send2 = up ? s2[i] : s2[i + off];
            const float keep2 = up ? s2[i + off] : s2[i];
            s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
        }
    }
}

__device__ __forceinline__ void px_butterfly1(float (&v)[32], int lane) {     // sums only (column sums without squares)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// Fold the per-row running sums into the per-column fp64 accumulators of the owner threads.  Collective over the
// kPxEpiWarps epilogue warps (named barrier 1).  Rare: when the statistics group or the n-tile changes and at the end.
template <int BN>
__device__ __forceinline__ void px_stat_fold_rows(PxStatAcc& a, int q, int ew, int lane, float* s_part) {
    using RA = PxRowAcc<BN>;
    const int half = ew >> 2;
#pragma unroll
    for (int cc = 0; cc < RA::kChunks; ++cc) {
        float v[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {       // this lane's columns of the chunk, zero elsewhere (PxStatAcc layout)
            const bool own = (RA::kSteps < 1 || ((i >> 4) & 1) == (lane & 1)) && (RA::kSteps < 2 || ((i >> 3) & 1) == ((lane >> 1) & 1));
            v[i] = own ? a.rs[cc * RA::kVals + (i & (RA::kVals - 1))] : 0.f;
            s2[i] = own ? a.rq[cc * RA::kVals + (i & (RA::kVals - 1))] : 0.f;
        }
        px_butterfly(v, s2, lane);
        const int ch = half + cc * (kPxEpiWarps / 4);
        s_part[(q * 2 + 0) * BN + ch * 32 + lane] = v[0];
        s_part[(q * 2 + 1) * BN + ch * 32 + lane] = s2[0];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kPxEpiWarps) : "memory");
    const int c = ew * 32 + lane;
    if (c < BN) {
        float s = 0.f, sq = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s += s_part[(e * 2 + 0) * BN + c];
            sq += s_part[(e * 2 + 1) * BN + c];
        }
        if (a.grp == 0) { a.s0 += static_cast<double>(s); a.q0 += static_cast<double>(sq); }
        else { a.s1 += static_cast<double>(s); a.q1 += static_cast<double>(sq); }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kPxEpiWarps) : "memory");
    a.reset_rows();
}
template <int BN>
__device__ __forceinline__ void px_stat_flush(const PxParams& p, PxStatAcc& a, int ew, int lane, float* s_part) {
    if constexpr (PxRowAcc<BN>::kSteps >= 0) {
        if (p.stat_sum != nullptr && a.pending) px_stat_fold_rows<BN>(a, (ew & 3), ew, lane, s_part);
    }
    const int c = ew * 32 + lane;
    if (a.n_tile >= 0 && c < BN && p.stat_sum != nullptr) {
        const long long col = static_cast<long long>(a.n_tile) * BN + c;
        if (a.s0 != 0.0 || a.q0 != 0.0) {
            atomicAdd(p.stat_sum + col, a.s0);
            if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + col, a.q0);
        }
        if (a.s1 != 0.0 || a.q1 != 0.0) {
            atomicAdd(p.stat_sum + p.stat_gstride + col, a.s1);
            if (p.stat_sq != nullptr) atomicAdd(p.stat_sq + p.stat_gstride + col, a.q1);
        }
    }
    a.reset(-1);
}

// RED kernels: the epilogue of a tile reads the previous layer's raw output y for the tile's pixels.  Hint the rows of the
// NEXT tile into the cache one tile ahead (L1 for 64-wide tiles: 16 KB per tile; L2 otherwise), so that the loads in the
// epilogue do not stall it for a DRAM round trip per 32-channel chunk.
template <int BN>
__device__ __forceinline__ void px_red_prefetch(const PxParams& p, int m_tile, int n_tile, int q, int ew, int lane) {
    if (m_tile >= p.num_m_tiles) return;
    const int half = ew >> 2;
    const PxRowCoord rc = px_row_coord(p, m_tile, q, lane);
    if (!rc.valid) return;
    const __nv_bfloat16* yrow = p.red_y + ((static_cast<long long>(rc.n) * p.H + rc.h) * p.W + rc.w) * p.cout_total + n_tile * BN;
#pragma unroll
    for (int ch = half; ch < BN / 32; ch += kPxEpiWarps / 4) {
        if (BN == 64) asm volatile("prefetch.global.L1 [%0];" ::"l"(yrow + ch * 32));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(yrow + ch * 32));
    }
}

// Store epilogue of the pixel-major kernels: raw bf16 output + BatchNorm partial sums.
// `rc`: this thread's output pixel (px_row_coord, computed before the accumulator wait); `arrive_bar`: the accumulator-drained
// barrier; `remote`: it is a shared::cluster address in the peer (leader) CTA.
// bf16 tiles of up to 128 columns release the accumulator as soon as this warp's columns are in registers (packed to bf16) and do
// the stores and the statistics afterwards: with 64 input channels a tile's MMAs take about as long as its epilogue, and an
// accumulator held through the whole epilogue stalled the MMA warp (forward 64 -> 64 was 14 % slower than its dgrad form).
template <int BN, bool RED = false, class Op = OpBf16>
__device__ __forceinline__ void px_store_epilogue(const PxParams& p, const PxRowCoord& rc, int n_tile, int acc, uint32_t tmem_base, int q,
                                                  int ew, int lane, float* s_part, uint32_t arrive_bar, bool remote,
                                                  PxStatAcc& sacc) {
    static_assert(!(RED && Op::kTf32), "the fused BatchNorm-backward reduce exists for the bf16 kernels only");
    using OT = typename Op::T;
    // ew = 0 .. kPxEpiWarps-1; warps ew and ew+4 share TMEM lane quarter q and split the 32-column chunks
    const int half = ew >> 2;
    const int n = rc.n, h = rc.h, w = rc.w, nt = rc.nt, co0 = n_tile * BN;
    const bool valid = rc.valid;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
    OT* orow = static_cast<OT*>(p.out) + ((static_cast<long long>(n) * p.H + h) * p.W + w) * p.ldo + p.out_coff + co0;
    const bool do_stats = p.stat_sum != nullptr;
    const int grp = px_group(p, nt);
    // per-row running sums (PxStatAcc): 64-column tiles always; 128-column bf16 tiles after a packed lane exchange
    constexpr bool kPackedAcc = !Op::kTf32 && !RED && PxRowAcc<BN>::kSteps >= 1;
    constexpr int kChunksPerWarp = BN >= 64 ? BN / 64 : 1;
    constexpr bool kEarlyRelease = !Op::kTf32 && BN >= 64 && kChunksPerWarp <= 2;
    const bool rowacc = do_stats && (BN == 64 || kPackedAcc);
    if (rowacc) {      // (collective) fold the running row sums when the statistics group or n-tile changes
        if (sacc.n_tile != n_tile) {
            px_stat_flush<BN>(p, sacc, ew, lane, s_part);
            sacc.reset(n_tile);
            sacc.grp = grp;
        } else if (sacc.grp != grp) {
            if (sacc.pending) px_stat_fold_rows<BN>(sacc, q, ew, lane, s_part);
            sacc.grp = grp;
        }
        sacc.pending = 1;
    }
    auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (remote) mbar_arrive_cluster(arrive_bar);
            else mbar_arrive(arrive_bar);
        }
    };
    // accumulator chunk -> registers (inference: BatchNorm(eval) + ReLU applied), bf16: packed as stored
    auto load_chunk = [&](int ch, uint32_t (&r)[32], uint32_t (&pk)[16]) {
        tmem_ld_32x32(t_addr + ch * 32, r);
        tmem_ld_wait();
        if (p.scale != nullptr) {      // inference: y = relu(acc * scale[c] + shift[c])
            const float4* sc4 = reinterpret_cast<const float4*>(p.scale + static_cast<long long>(grp) * p.cout_total + co0 + ch * 32);
            const float4* sh4 = reinterpret_cast<const float4*>(p.shift + static_cast<long long>(grp) * p.cout_total + co0 + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 sc = __ldg(sc4 + j), sh = __ldg(sh4 + j);
                r[4 * j] = __float_as_uint(fmaxf(fmaf(__uint_as_float(r[4 * j]), sc.x, sh.x), 0.f));
                r[4 * j + 1] = __float_as_uint(fmaxf(fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y), 0.f));
                r[4 * j + 2] = __float_as_uint(fmaxf(fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z), 0.f));
                r[4 * j + 3] = __float_as_uint(fmaxf(fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w), 0.f));
            }
        }
        if constexpr (!Op::kTf32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        }
    };
    // store + statistics of one chunk (r: fp32 values, used by the tf32 kernels only; pk: packed bf16)
    auto finish_chunk = [&](int ch, const uint32_t* r, const uint32_t (&pk)[16]) {
        if constexpr (Op::kTf32) {     // fp32 storage: 128 bytes per pixel and chunk, stored as they are
            if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
        } else {
            if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
        }
        if constexpr (kPackedAcc) {
            if (do_stats) {
                const int cc = (ch - half) / (kPxEpiWarps / 4);      // warp-uniform
                if (cc == 0) px_rowacc_packed<1, 0>(pk, valid, lane, sacc);
                else px_rowacc_packed<1, 1>(pk, valid, lane, sacc);
            }
        } else if (do_stats) {
            float v[32], s2[32];
            if constexpr (Op::kTf32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = valid ? __uint_as_float(r[j]) : 0.f;
                    s2[j] = v[j] * v[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {     // statistics of the values as stored (bf16-rounded)
                    const uint32_t wv = valid ? pk[j] : 0u;
                    const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xffff0000u);
                    v[2 * j] = lo; v[2 * j + 1] = hi;
                    s2[2 * j] = lo * lo; s2[2 * j + 1] = hi * hi;
                }
            }
            if (RED) {      // BatchNorm-backward reduce of the previous layer: v = dz, s2 = dz * (y - mean) * invstd
                const long long cbase = static_cast<long long>(grp) * p.cout_total + co0 + ch * 32;
                const uint4* y4 = reinterpret_cast<const uint4*>(
                    p.red_y + ((static_cast<long long>(n) * p.H + h) * p.W + w) * p.cout_total + co0 + ch * 32);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const uint4 yr = valid ? __ldg(y4 + j4) : make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t yw[4] = {yr.x, yr.y, yr.z, yr.w};
                    const float4 sca = __ldg(reinterpret_cast<const float4*>(p.red_scale + cbase) + 2 * j4);
                    const float4 scb = __ldg(reinterpret_cast<const float4*>(p.red_scale + cbase) + 2 * j4 + 1);
                    const float4 sha = __ldg(reinterpret_cast<const float4*>(p.red_shift + cbase) + 2 * j4);
                    const float4 shb = __ldg(reinterpret_cast<const float4*>(p.red_shift + cbase) + 2 * j4 + 1);
                    const float4 mua = __ldg(reinterpret_cast<const float4*>(p.red_mean + cbase) + 2 * j4);
                    const float4 mub = __ldg(reinterpret_cast<const float4*>(p.red_mean + cbase) + 2 * j4 + 1);
                    const float4 isa = __ldg(reinterpret_cast<const float4*>(p.red_invstd + cbase) + 2 * j4);
                    const float4 isb = __ldg(reinterpret_cast<const float4*>(p.red_invstd + cbase) + 2 * j4 + 1);
                    const float is[8] = {isa.x, isa.y, isa.z, isa.w, isb.x, isb.y, isb.z, isb.w};
                    const float sc[8] = {sca.x, sca.y, sca.z, sca.w, scb.x, scb.y, scb.z, scb.w};
                    const float sh[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
                    const float mu[8] = {mua.x, mua.y, mua.z, mua.w, mub.x, mub.y, mub.z, mub.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float y = __uint_as_float((e & 1) ? (yw[e >> 1] & 0xffff0000u) : (yw[e >> 1] << 16));
                        // ReLU mask = sign of the pre-activation (relu_open, elementwise.cuh)
                        const float dz = fmaf(y, sc[e], sh[e]) > 0.f ? v[8 * j4 + e] : 0.f;
                        v[8 * j4 + e] = dz;
                        s2[8 * j4 + e] = dz * ((y - mu[e]) * is[e]);
                    }
                }
            }
            if (BN == 64) {          // one chunk per warp: keep per-row running sums, no cross-lane traffic per tile
#pragma unroll
                for (int j = 0; j < 32; ++j) { sacc.rs[j] += v[j]; sacc.rq[j] += s2[j]; }
            } else if (p.stat_sq != nullptr) {
                px_butterfly(v, s2, lane);
                s_part[(q * 2 + 0) * BN + ch * 32 + lane] = v[0];
                s_part[(q * 2 + 1) * BN + ch * 32 + lane] = s2[0];
            } else {                 // column sums only (bias gradient of the up-convolution)
                px_butterfly1(v, lane);
                s_part[(q * 2 + 0) * BN + ch * 32 + lane] = v[0];
                s_part[(q * 2 + 1) * BN + ch * 32 + lane] = 0.f;
            }
        }
    };
    if constexpr (kEarlyRelease) {
        uint32_t pk[kChunksPerWarp][16];
#pragma unroll
        for (int cc = 0; cc < kChunksPerWarp; ++cc) {
            uint32_t r[32];
            load_chunk(half + cc * (kPxEpiWarps / 4), r, pk[cc]);
        }
        release();
#pragma unroll
        for (int cc = 0; cc < kChunksPerWarp; ++cc) finish_chunk(half + cc * (kPxEpiWarps / 4), nullptr, pk[cc]);
    } else {
#pragma unroll 1
        for (int ch = half; ch < BN / 32; ch += kPxEpiWarps / 4) {
            uint32_t r[32], pk[16];
            load_chunk(ch, r, pk);
            finish_chunk(ch, r, pk);
        }
        release();
    }
    if (do_stats && !rowacc) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kPxEpiWarps) : "memory");
        if (sacc.n_tile != n_tile) {
            px_stat_flush<BN>(p, sacc, ew, lane, s_part);
            sacc.reset(n_tile);
        }
        const int c = ew * 32 + lane;
        if (c < BN) {
            float s = 0.f, sq = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s += s_part[(e * 2 + 0) * BN + c];
                sq += s_part[(e * 2 + 1) * BN + c];
            }
            if (grp == 0) { sacc.s0 += static_cast<double>(s); sacc.q0 += static_cast<double>(sq); }
            else { sacc.s1 += static_cast<double>(s); sacc.q1 += static_cast<double>(sq); }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kPxEpiWarps) : "memory");
    }
}

template <int BN, class Op = OpBf16>
__global__ void __launch_bounds__(kPxThreads, 1)
tapgemm_px_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PxParams p) {
    using OT = typename Op::T;
    constexpr int KC = Op::kKC;
    using Cfg = PxCfg<BN>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t aux = base + STAGES * Cfg::kStageBytes;
    uint8_t* gen_aux = gen_base + STAGES * Cfg::kStageBytes;
    // aux layout: full[STAGES] | empty[STAGES] | tfull[2] | tempty[2] | tmem_ptr | ... | stat partials @1024
    const uint32_t bar_full = aux, bar_empty = aux + 8 * STAGES, bar_tfull = aux + 16 * STAGES,
                   bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * STAGES + 32);
    float* s_part = reinterpret_cast<float*>(gen_aux + 1024);   // [4 warps][2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, kPxEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int k_iters = p.ntaps * p.k_chunks;
    const uint32_t a_tx = static_cast<uint32_t>(p.valid_rows) * 128u;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = fd_div(tile, p.fd_m), m_tile = tile - n_tile * p.num_m_tiles;
                int wt, ht, nt;
                px_tile_coord(p, m_tile, wt, ht, nt);
                const int w0 = wt * p.TW, h0 = ht * p.TH, n0 = nt * p.TN, co0 = n_tile * BN;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sA = base + stage * Cfg::kStageBytes, sB = sA + Cfg::kABytes;
                        const uint32_t fb = bar_full + 8 * stage;
                        mbar_expect_tx(fb, a_tx + Cfg::kBBytes);
                        const int4 tp = p.taps[t];
                        tma_load_5d(sA, &tmA, fb, tp.x + kc * KC, w0 + tp.y, tp.z, h0 + tp.w, n0);
                        tma_load_2d(sB, &tmB, fb, p.tap_w[t] * p.cin + kc * KC, co0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one elected thread)
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc<Op>(128, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sA = base + stage * Cfg::kStageBytes, sB = sA + Cfg::kABytes;
                    const uint64_t da = umma_smem_desc(sA, 16, 1024), db = umma_smem_desc(sB, 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)   // 4 x (32 B of K: 16 bf16 / 8 tf32) inside the 128-B swizzle row
                        umma<Op>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0);
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (4 warps)
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int ew = warp - 2;           // 0..kPxEpiWarps-1; warps ew and ew+4 share quarter q and split the column chunks
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const int w_l = row & (p.TW - 1), h_l = (row >> p.log_tw) & (p.TH - 1), n_l = row >> (p.log_tw + p.log_th);
        PxStatAcc sacc;
        sacc.reset(-1);
        sacc.reset_rows();
        sacc.grp = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n_tile = fd_div(tile, p.fd_m), m_tile = tile - n_tile * p.num_m_tiles;
            int wt, ht, nt;
            px_tile_coord(p, m_tile, wt, ht, nt);
            const int w = wt * p.TW + w_l, h = ht * p.TH + h_l, n = nt * p.TN + n_l, co0 = n_tile * BN;
            const bool valid = (row < p.valid_rows) && (w < p.W) && (h < p.H) && (n < p.N);
            PxRowCoord rc;
            rc.n = n; rc.h = h; rc.w = w; rc.nt = nt; rc.valid = valid;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

            if (p.epi_mode == EPI_STORE) {
                px_store_epilogue<BN, false, Op>(p, rc, n_tile, acc, tmem_base, q, ew, lane, s_part, bar_tempty + 8 * acc, false, sacc);
            } else {
                // EPI_CONVT: column = (tap, co); scatter to the 2x upsampled grid, add bias.  A tile may span several
                // taps (BN up to 4 * co_per_tap); every 32-column chunk lies inside one tap (co_per_tap % 32 == 0).
#pragma unroll 1
                for (int ch = 2 * half; ch < BN / 32; ch += kPxEpiWarps / 2) {   // two 32-column chunks in flight per iteration
                    uint32_t r[2][32];
                    tmem_ld_32x32(t_addr + ch * 32, r[0]);
                    tmem_ld_32x32(t_addr + ch * 32 + 32, r[1]);
                    tmem_ld_wait();
                    if constexpr (!Op::kTf32) {
                        if (p.TW >= 8 && (p.co_per_tap & 63) == 0) {
                            // Full-line stores.  A thread owns one pixel; written directly, every 16-byte store of a warp would go to a
                            // different 128-byte line (32 LSU wavefronts per instruction: the scatter was LSU-bound at ~3 TB/s).  The two
                            // chunks are the 64 channels = 128 contiguous bytes of ONE fine pixel: transpose the 8 x 16-byte pieces across
                            // the 8 lanes that hold 8 consecutive pixels of a tile row, so that 8 consecutive lanes write one whole line.
                            const int colg = co0 + ch * 32;
                            const int tap = colg / p.co_per_tap, cbase = colg % p.co_per_tap;
                            const int dy = tap >> 1, dx = tap & 1;
                            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + cbase);
                            uint4 qv[8];
#pragma unroll
                            for (int u = 0; u < 2; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 b0 = __ldg(bias4 + 8 * u + 2 * j), b1 = __ldg(bias4 + 8 * u + 2 * j + 1);
                                    qv[4 * u + j] =
                                        make_uint4(pack_bf16x2(__uint_as_float(r[u][j * 8 + 0]) + b0.x, __uint_as_float(r[u][j * 8 + 1]) + b0.y),
                                                   pack_bf16x2(__uint_as_float(r[u][j * 8 + 2]) + b0.z, __uint_as_float(r[u][j * 8 + 3]) + b0.w),
                                                   pack_bf16x2(__uint_as_float(r[u][j * 8 + 4]) + b1.x, __uint_as_float(r[u][j * 8 + 5]) + b1.y),
                                                   pack_bf16x2(__uint_as_float(r[u][j * 8 + 6]) + b1.z, __uint_as_float(r[u][j * 8 + 7]) + b1.w));
                                }
                            const int jl = lane & 7;
#pragma unroll
                            for (int b = 1; b < 8; b <<= 1) {          // 8 x 8 transpose of 16-byte pieces inside each group of 8 lanes
                                const bool up = (jl & b) != 0;
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    if (e & b) continue;
                                    const uint4 send = up ? qv[e] : qv[e | b];
                                    uint4 recv;
                                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, b);
                                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, b);
                                    recv.z = __shfl_xor_sync(0xffffffffu, send.z, b);
                                    recv.w = __shfl_xor_sync(0xffffffffu, send.w, b);
                                    if (up) qv[e] = recv; else qv[e | b] = recv;
                                }
                            }
                            // now qv[i] = piece jl of pixel (w - jl + i): lanes jl = 0..7 of a group cover that pixel's 128 bytes
                            const bool vrow = (row < p.valid_rows) && (h < p.H) && (n < p.N);       // uniform inside the group of 8 lanes
                            const int wb = w - jl;
                            __nv_bfloat16* lbase = static_cast<__nv_bfloat16*>(p.out) +
                                                   ((static_cast<long long>(n) * p.Ho + (2 * h + dy)) * p.Wo + (2 * wb + dx)) * p.ldo + p.out_coff + cbase;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (vrow && wb + i < p.W) reinterpret_cast<uint4*>(lbase + static_cast<long long>(i) * 2 * p.ldo)[jl] = qv[i];
                            continue;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int colg = co0 + (ch + u) * 32;
                        const int tap = colg / p.co_per_tap, cbase = colg % p.co_per_tap;
                        const int dy = tap >> 1, dx = tap & 1;
                        OT* dstT = static_cast<OT*>(p.out) +
                                   ((static_cast<long long>(n) * p.Ho + (2 * h + dy)) * p.Wo + (2 * w + dx)) * p.ldo + p.out_coff + cbase;
                        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + cbase);
                        if (valid) {
                            uint4* dst = reinterpret_cast<uint4*>(dstT);
                            if constexpr (Op::kTf32) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 b0 = __ldg(bias4 + j);
                                    dst[j] = make_uint4(__float_as_uint(__uint_as_float(r[u][j * 4 + 0]) + b0.x),
                                                        __float_as_uint(__uint_as_float(r[u][j * 4 + 1]) + b0.y),
                                                        __float_as_uint(__uint_as_float(r[u][j * 4 + 2]) + b0.z),
                                                        __float_as_uint(__uint_as_float(r[u][j * 4 + 3]) + b0.w));
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
                                    dst[j] = make_uint4(pack_bf16x2(__uint_as_float(r[u][j * 8 + 0]) + b0.x, __uint_as_float(r[u][j * 8 + 1]) + b0.y),
                                                        pack_bf16x2(__uint_as_float(r[u][j * 8 + 2]) + b0.z, __uint_as_float(r[u][j * 8 + 3]) + b0.w),
                                                        pack_bf16x2(__uint_as_float(r[u][j * 8 + 4]) + b1.x, __uint_as_float(r[u][j * 8 + 5]) + b1.y),
                                                        pack_bf16x2(__uint_as_float(r[u][j * 8 + 6]) + b1.z, __uint_as_float(r[u][j * 8 + 7]) + b1.w));
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            }
        }
        px_stat_flush<BN>(p, sacc, ew, lane, s_part);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// =====================================================================================================
// weight-gradient kernel
// =====================================================================================================
// Work-unit order of the split-K weight-gradient kernels.  The decode below expects index = tile * ksplit + ks; with
// ks_slowest the CTAs walk the units pixel range by pixel range instead, so that all (m, n) tiles of one pixel range run
// at the same time and share its G / In slabs through L2 (ncu: 2.6x fewer DRAM bytes for 256 -> 256 at 64 x 64).
__device__ __forceinline__ int wg_unit_index(int u, int num_units, int ksplit, int ks_slowest) {
    if (!ks_slowest) return u;
    const int tiles = num_units / ksplit;
    return (u % tiles) * ksplit + u / tiles;
}

constexpr int kWgMaxAcc = 3;
constexpr int kWgMaxGroups = 4;

struct WgGroup {          // one work-unit type: up to 3 accumulators, each M=128 made of two 64-channel blocks
    int nacc;
    int tapA[kWgMaxAcc], offA[kWgMaxAcc];   // tap index, channel offset (relative to the m-tile base) of rows 0..63
    int tapB[kWgMaxAcc], offB[kWgMaxAcc];   // ... of rows 64..127; tapB < 0 -> unused (rows ignored)
};

struct WgParams {
    int N, H, W;                         // pixel grid that is reduced over
    int TW, TH, TN;                      // pixel slab = TW*TH*TN = 64 pixels (bf16) / 32 pixels (tf32)
    int tiles_w, tiles_h, tiles_n, num_px_tiles;
    int ksplit, px_tiles_per_split;
    int ngroups, num_m_tiles, num_n_tiles;
    int m_tile_channels;                 // 128 (two adjacent blocks) or 64 (two taps)
    int4 taps[kMaxTaps];                 // coordinate offsets (dc, dw, dq, dh) applied to the M-side (G) operand
    int ntaps;
    WgGroup groups[kWgMaxGroups];
    float* out;                          // dW, PyTorch layout, accumulated with atomics
    int m_total, n_total;                // out index = (m*n_total + n)*ntaps + t, or (n*m_total + m)*ntaps + t
    int out_transposed;
    int ks_slowest;
};

// One operand "box" = one pixel slab x one 128-byte row of channels: 64 pixels x 64 bf16 channels (8 KB), or - fp32 operands
// read as TF32 - 32 pixels x 32 channels (4 KB; the slab is halved so that three stages of N-side + 3 accumulators x M = 128
// still fit in shared memory).  An M = 128 accumulator is made of 128 / kKC boxes, LBO = one box apart.
template <int BNW, class Op = OpBf16>
struct WgCfg {
    static constexpr int kSlabPx = Op::kTf32 ? 32 : 64;
    static constexpr int kBoxBytes = kSlabPx * 128;
    static constexpr int kBoxesPerHalf = 64 / Op::kKC;            // boxes per 64-channel half of an accumulator's M operand
    static constexpr int kNBoxes = BNW / Op::kKC;
    static constexpr int kNBytes = kNBoxes * kBoxBytes;          // N-side (unshifted) operand per stage
    static constexpr int kAccBytes = 2 * kBoxesPerHalf * kBoxBytes;
    static constexpr int kMBytes = kWgMaxAcc * kAccBytes;         // M-side: up to 3 accumulators x (2 x 64 channels)
    static constexpr int kStageBytes = kNBytes + kMBytes;
    static constexpr int kStages = 3;
    static constexpr int kKSteps = 4;                             // MMAs per slab: 64 px / K16 (bf16), 32 px / K8 (tf32)
    static constexpr int kKStepUnits = (Op::kTf32 ? 1024 : 2048) / 16;   // descriptor advance per MMA, in 16-byte units
    static constexpr int kTmemCols = 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
};

template <int BNW, class Op = OpBf16>
__global__ void __launch_bounds__(192, 1)
tapgemm_wg_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI, const WgParams p) {
    using Cfg = WgCfg<BNW, Op>;
    constexpr int STAGES = Cfg::kStages;
    constexpr int KC = Op::kKC, BOX = Cfg::kBoxBytes, BPH = Cfg::kBoxesPerHalf;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - raw);
    const uint32_t aux = base + STAGES * Cfg::kStageBytes;
    uint8_t* gen_aux = gen_base + STAGES * Cfg::kStageBytes;
    const uint32_t bar_full = aux, bar_empty = aux + 8 * STAGES, bar_tfull = aux + 16 * STAGES,
                   bar_tempty = bar_tfull + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_aux + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmI);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_smem));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // unit = ((g * num_m_tiles + mt) * num_n_tiles + nt) * ksplit + ks
    const int num_units = p.ngroups * p.num_m_tiles * p.num_n_tiles * p.ksplit;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const int nt = (unit / p.ksplit) % p.num_n_tiles;
                const int mt = (unit / (p.ksplit * p.num_n_tiles)) % p.num_m_tiles;
                const int g = unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles);
                const WgGroup& grp = p.groups[g];
                const int m0 = mt * p.m_tile_channels, n0 = nt * BNW;
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                int nboxes = 0;
                for (int a = 0; a < grp.nacc; ++a) nboxes += (grp.tapB[a] >= 0) ? 2 * BPH : BPH;
                const uint32_t tx = static_cast<uint32_t>(nboxes) * BOX + Cfg::kNBytes;
                for (int pt = px_begin; pt < px_end; ++pt) {
                    const int wt = pt % p.tiles_w, ht = (pt / p.tiles_w) % p.tiles_h, nn = pt / (p.tiles_w * p.tiles_h);
                    const int w0 = wt * p.TW, h0 = ht * p.TH, nimg0 = nn * p.TN;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    const uint32_t fb = bar_full + 8 * stage;
                    mbar_expect_tx(fb, tx);
#pragma unroll
                    for (int b = 0; b < Cfg::kNBoxes; ++b) tma_load_5d(sN + b * BOX, &tmI, fb, n0 + b * KC, w0, 0, h0, nimg0);
                    for (int a = 0; a < grp.nacc; ++a) {
                        const int4 ta = p.taps[grp.tapA[a]];
#pragma unroll
                        for (int b = 0; b < BPH; ++b)
                            tma_load_5d(sM + a * Cfg::kAccBytes + b * BOX, &tmG, fb, m0 + grp.offA[a] + b * KC + ta.x, w0 + ta.y, ta.z,
                                        h0 + ta.w, nimg0);
                        if (grp.tapB[a] >= 0) {
                            const int4 tb = p.taps[grp.tapB[a]];
#pragma unroll
                            for (int b = 0; b < BPH; ++b)
                                tma_load_5d(sM + a * Cfg::kAccBytes + (BPH + b) * BOX, &tmG, fb, m0 + grp.offB[a] + b * KC + tb.x,
                                            w0 + tb.y, tb.z, h0 + tb.w, nimg0);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc<Op>(128, BNW, 1, 1);   // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x, ++it) {
                const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
                const int ks = unit % p.ksplit;
                const int g = unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles);
                const int nacc = p.groups[g].nacc;
                const int px_begin = ks * p.px_tiles_per_split;
                const int px_end = min(px_begin + p.px_tiles_per_split, p.num_px_tiles);
                mbar_wait(bar_tempty, (it & 1) ^ 1);
                tc_fence_after();
                for (int pt = px_begin; pt < px_end; ++pt) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sN = base + stage * Cfg::kStageBytes, sM = sN + Cfg::kNBytes;
                    // MN-major operands: bf16 = 128-byte swizzle over 8 pixel rows; tf32 = 128-byte swizzle with 32-byte atoms over
                    // 4 pixel rows (the only MN-major layout the tensor core accepts for 32-bit elements)
                    constexpr uint32_t LT = Op::kTf32 ? kUmmaSwizzle128BAtom32B : kUmmaSwizzle128B;
                    constexpr uint32_t SBO = Op::kTf32 ? 512 : 1024;
                    const uint64_t db = umma_smem_desc(sN, BOX, SBO, LT);
                    for (int a = 0; a < nacc; ++a) {
                        const uint64_t da = umma_smem_desc(sM + a * Cfg::kAccBytes, BOX, SBO, LT);
#pragma unroll
                        for (int kk = 0; kk < Cfg::kKSteps; ++kk)   // 16 pixel rows = 2048 B (bf16) / 8 pixel rows = 1024 B (tf32) per MMA
                            umma<Op>(tmem_base + a * BNW, da + Cfg::kKStepUnits * kk, db + Cfg::kKStepUnits * kk, idesc,
                                     (pt > px_begin) || kk);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int unit_ = blockIdx.x; unit_ < num_units; unit_ += gridDim.x, ++it) {
            const int unit = wg_unit_index(unit_, num_units, p.ksplit, p.ks_slowest);
            const int nt = (unit / p.ksplit) % p.num_n_tiles;
            const int mt = (unit / (p.ksplit * p.num_n_tiles)) % p.num_m_tiles;
            const int g = unit / (p.ksplit * p.num_n_tiles * p.num_m_tiles);
            const WgGroup& grp = p.groups[g];
            const int m0 = mt * p.m_tile_channels, n0 = nt * BNW;
            mbar_wait(bar_tfull, it & 1);
            tc_fence_after();
            for (int a = 0; a < grp.nacc; ++a) {
                const int tap = (row < 64) ? grp.tapA[a] : grp.tapB[a];
                const int m = m0 + ((row < 64) ? grp.offA[a] + row : grp.offB[a] + row - 64);
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BNW;
#pragma unroll 1
                for (int ch = 0; ch < BNW / 32; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + ch * 32, r);
                    tmem_ld_wait();
                    if (tap >= 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + ch * 32 + j;
                            const long long idx = p.out_transposed
                                                      ? (static_cast<long long>(n) * p.m_total + m) * p.ntaps + tap
                                                      : (static_cast<long long>(m) * p.n_total + n) * p.ntaps + tap;
                            atomicAdd(p.out + idx, __uint_as_float(r[j]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace onet
