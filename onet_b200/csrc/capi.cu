// C ABI of libonet_b200.so: host-side launchers for the kernels in this directory (see include/onet_b200.h).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/onet_b200.h"
#include "elementwise.cuh"
#include "first_layer.cuh"
#include "first_layer_mma.cuh"
#include "head_mma.cuh"
#include "simt_conv.cuh"
#include "synth.cuh"
#include "tapgemm_tc.cuh"
#include "conv3x3_halo_tc.cuh"

using namespace onet;
typedef __nv_bfloat16 bf16;

static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
static long long g_launches = 0;
static thread_local char g_last_kernel[96] = "";     // kernel variant of the most recent launch (onet_last_kernel)
static int check_launch(const char* what, int variant = -1) {
    ++g_launches;
    if (variant >= 0) snprintf(g_last_kernel, sizeof(g_last_kernel), "%s<%d>", what, variant);
    else snprintf(g_last_kernel, sizeof(g_last_kernel), "%s", what);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
    return 0;
}
static inline int grid_for(long long items, int block, int cap = 148 * 16) {
    long long g = (items + block - 1) / block;
    return static_cast<int>(std::max(1LL, std::min<long long>(g, cap)));
}
static inline int p2floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}
static inline int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 5-D view (c, w, q, h, n) with 128-byte swizzle; strides in ELEMENTS for dims 1..4; elem_bytes = 2 (bf16) or 4 (fp32, which
// the tf32 kernels read as they are: the tensor core ignores the low 13 mantissa bits of each operand word)
static int make_map5(CUtensorMap* m, const void* base, const uint64_t dims[5], const uint64_t strides_el[4],
                     const uint32_t box[5], int elem_bytes, bool mn_major = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < 5; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < 4; ++i) gs[i] = strides_el[i] * elem_bytes;
    if (reinterpret_cast<uintptr_t>(base) % 16) return fail("tensor map base not 16-byte aligned");
    for (int i = 0; i < 4; ++i)
        if (gs[i] % 16) return fail("tensor map stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
    // MN-major (weight-gradient) operands of 32-bit elements must sit in shared memory in the "128-byte swizzle, 32-byte atom"
    // pattern (tc_common.cuh, umma_smem_desc); everything else uses the plain 128-byte swizzle
    const CUtensorMapSwizzle sw = (mn_major && elem_bytes == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                     const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(5d) failed with %d", static_cast<int>(r));
    return 0;
}
static int make_map2(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                     uint32_t box_outer, int elem_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[2] = {inner, outer}, gs[1] = {inner * elem_bytes};
    cuuint32_t bx[2] = {box_inner, box_outer}, es[2] = {1, 1};
    if (reinterpret_cast<uintptr_t>(base) % 16 || gs[0] % 16) return fail("weight tensor map misaligned");
    CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(2d) failed with %d", static_cast<int>(r));
    return 0;
}
// plain NHWC activation view: (C, W, 1, H, N)
template <typename T>
static int make_act_map(CUtensorMap* m, const T* base, int C, int N, int H, int W, long long ld, const uint32_t box[5],
                        bool mn_major = false) {
    const uint64_t dims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), 1, static_cast<uint64_t>(H),
                              static_cast<uint64_t>(N)};
    const uint64_t st[4] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(ld) * W, static_cast<uint64_t>(ld) * W,
                            static_cast<uint64_t>(ld) * W * H};
    return make_map5(m, base, dims, st, box, static_cast<int>(sizeof(T)), mn_major);
}
// fine grid [N,Ho,Wo,ld] (Ho >= 2H, Wo >= 2W) seen from the coarse grid: (c' = dx*ld + c, w, q = dy, h, n)
template <typename T>
static int make_up_map(CUtensorMap* m, const T* base, int C, int N, int H, int W, long long ld, const uint32_t box[5],
                       int Ho, int Wo, bool mn_major = false) {
    const uint64_t dims[5] = {static_cast<uint64_t>(ld + C), static_cast<uint64_t>(W), 2, static_cast<uint64_t>(H),
                              static_cast<uint64_t>(N)};
    const uint64_t st[4] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * Wo,
                            static_cast<uint64_t>(ld) * 2 * Wo, static_cast<uint64_t>(ld) * Wo * Ho};
    return make_map5(m, base, dims, st, box, static_cast<int>(sizeof(T)), mn_major);
}

// Function attributes (dynamic shared-memory opt-in, carve-out), co-resident cluster counts and the SM count are properties of
// ONE device: every cache below is indexed by the current device so that a process driving several GPUs stays correct.
constexpr int kMaxDevices = 64;
static int cur_dev() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}
static int sm_count() {
    static int count[kMaxDevices] = {};
    const int dev = cur_dev();
    if (count[dev] == 0) {
        cudaDeviceGetAttribute(&count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (count[dev] <= 0) count[dev] = 148;
    }
    return count[dev];
}

// Deterministic split-K (FP32 verification mode): when the caller has registered a workspace for the current device
// (onet_set_splitk_workspace), the CUDA-core weight-gradient / column-sum kernels write one slab of partial sums per pixel
// split and splitk_reduce_kernel adds the slabs in split order; without a (large enough) workspace they use fp32 atomics.
static float* g_splitk_ws[kMaxDevices] = {};
static long long g_splitk_floats[kMaxDevices] = {};
static float* splitk_ws(long long floats_needed) {
    const int dev = cur_dev();
    return (g_splitk_ws[dev] != nullptr && floats_needed <= g_splitk_floats[dev]) ? g_splitk_ws[dev] : nullptr;
}
static int splitk_reduce(const float* partial, int splits, long long numel, float* dst, cudaStream_t st) {
    const int grid = static_cast<int>(std::max(1LL, std::min<long long>((numel + 255) / 256, 148 * 8)));
    splitk_reduce_kernel<<<grid, 256, 0, st>>>(partial, splits, numel, dst);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pixel-major tcgen05 launcher
// ------------------------------------------------------------------------------------------------
template <int BN, class Op = OpBf16>
static int launch_px(const CUtensorMap& tA, const CUtensorMap& tB, const PxParams& p, cudaStream_t st) {
    using Cfg = PxCfg<BN>;
    static bool attr_set_[kMaxDevices] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_px_kernel<BN, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(px<%d>): %s", BN, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = std::min(tiles, sm_count());
    PxParams pf = p;
    px_set_fastdiv(pf);
    tapgemm_px_kernel<BN, Op><<<grid, kPxThreads, Cfg::kSmemBytes, st>>>(tA, tB, pf);
    return check_launch(Op::kTf32 ? "tapgemm_px_kernel/tf32" : "tapgemm_px_kernel", BN);
}

// Fill the pixel tiling of PxParams; TN is restricted to divide `group_images` so that a tile never straddles
// the two BatchNorm statistics groups.
static void px_tiling(PxParams& p, int N, int H, int W, int group_images) {
    p.N = N; p.H = H; p.W = W;
    p.TW = std::min(p2floor(W), 16);
    p.TH = std::min(p2floor(H), 128 / p.TW);
    int tn = 128 / (p.TW * p.TH);
    tn = std::min(tn, p2floor(N));
    if (group_images > 0)
        while (tn > 1 && (group_images % tn) != 0) tn >>= 1;
    p.TN = tn;
    p.log_tw = ilog2(p.TW);
    p.log_th = ilog2(p.TH);
    p.tiles_w = (W + p.TW - 1) / p.TW;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.tiles_n = (N + p.TN - 1) / p.TN;
    p.num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.valid_rows = p.TW * p.TH * p.TN;
}

template <int BN, class Op = OpBf16>
static int launch_halo_px(const CUtensorMap& tA, const CUtensorMap& tB, const PxParams& p, cudaStream_t st) {
    using Cfg = HaloCfg<BN>;
    static bool attr_set_[kMaxDevices] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_px_kernel<BN, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(halo_px<%d>): %s", BN, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    PxParams pf = p;
    px_set_fastdiv(pf);
    conv3x3_halo_px_kernel<BN, Op><<<std::min(tiles, sm_count()), kPxThreads, Cfg::kSmemBytes, st>>>(tA, tB, pf);
    return check_launch(Op::kTf32 ? "conv3x3_halo_px_kernel/tf32" : "conv3x3_halo_px_kernel", BN);
}

template <int BN, bool RED = false>
static int launch_halo_res_px(const CUtensorMap& tA, const CUtensorMap& tB, const PxParams& p, cudaStream_t st) {
    using Cfg = HaloResCfg<BN>;
    static bool attr_set_[kMaxDevices] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_res_px_kernel<BN, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(halo_res_px<%d>): %s", BN, cudaGetErrorString(e));
        attr_set = true;
    }
    PxParams pf = p;
    px_set_fastdiv(pf);
    conv3x3_halo_res_px_kernel<BN, RED><<<std::min(p.num_m_tiles, sm_count()), kPxThreads, Cfg::kSmemBytes, st>>>(tA, tB, pf);
    return check_launch(RED ? "conv3x3_halo_res_px_kernel+bnred" : "conv3x3_halo_res_px_kernel", BN);
}

// CTA-pair (cta_group::2) variant: grid = 2 x min(pair tiles, co-resident clusters)
template <int BN, bool RED = false, class Op = OpBf16>
static int launch_halo2_px(const CUtensorMap& tA, const CUtensorMap& tB, const PxParams& p, cudaStream_t st) {
    using Cfg = Halo2Cfg<BN>;
    static int max_clusters_[kMaxDevices] = {};
    int& max_clusters = max_clusters_[cur_dev()];
    if (max_clusters == 0) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_halo2_px_kernel<BN, RED, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(halo2_px<%d>): %s", BN, cudaGetErrorString(e));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(sm_count() & ~1);
        cfg.blockDim = dim3(kPxThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv3x3_halo2_px_kernel<BN, RED, Op>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = sm_count() / 2;
        }
        max_clusters = std::min(n, sm_count() / 2);
    }
    const int units = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int grid = 2 * std::min(units, max_clusters);
    PxParams pf = p;
    px_set_fastdiv(pf);
    conv3x3_halo2_px_kernel<BN, RED, Op><<<grid, kPxThreads, Cfg::kSmemBytes, st>>>(tA, tB, pf);
    return check_launch(RED ? "conv3x3_halo2_px_kernel+bnred" : (Op::kTf32 ? "conv3x3_halo2_px_kernel/tf32" : "conv3x3_halo2_px_kernel"), BN);
}

// Split-K factor for the weight-gradient kernels: minimise (waves x K-steps per unit + fixed per-unit epilogue cost).
static void pick_ksplit(int base_units, int num_px_tiles, int* ksplit, int* per_split) {
    const int sms = sm_count();
    long long best_cost = -1;
    int best_per = num_px_tiles;
    const int max_ks = std::min(num_px_tiles, 8 * sms);
    for (int ks = 1; ks <= max_ks; ++ks) {
        const int per = (num_px_tiles + ks - 1) / ks;
        const int ks_eff = (num_px_tiles + per - 1) / per;
        const long long units = static_cast<long long>(base_units) * ks_eff;
        const long long waves = (units + sms - 1) / sms;
        const long long cost = waves * (per + 6);           // ~6 K-steps worth of epilogue / pipeline fill per unit
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_per = per; }
    }
    *per_split = best_per;
    *ksplit = (num_px_tiles + best_per - 1) / best_per;
}

// BatchNorm-backward reduce of the previous layer folded into a dgrad launch (PxParams::red_y)
struct BnRedArgs {
    const bf16* y; const float* scale; const float* shift; const float* mean; const float* invstd;
    double* sums;                                                                              // [G][2][Cout]
};
static void apply_bnred(PxParams& p, const BnRedArgs* red, int Cout) {
    if (red == nullptr || red->y == nullptr) return;
    p.red_y = red->y; p.red_scale = red->scale; p.red_shift = red->shift; p.red_mean = red->mean; p.red_invstd = red->invstd;
    p.stat_sum = red->sums; p.stat_sq = red->sums + Cout; p.stat_gstride = 2 * Cout;
}
// The same sums from the standalone reduce pass (kernel variants without a fused-reduce instantiation: small images)
static int bnred_standalone(const BnRedArgs& red, const bf16* g, int N, int H, int W, int C, int group_images, cudaStream_t st) {
    BnBwdArgs<bf16> a;
    memset(&a, 0, sizeof(a));
    a.y = red.y; a.N = N; a.H = H; a.W = W; a.C = C;
    a.scale = red.scale; a.shift = red.shift; a.mean = red.mean; a.invstd = red.invstd;
    a.group_images = group_images > 0 ? group_images : N;
    a.g1 = g; a.ld1 = C; a.off1 = 0;
    a.sums = red.sums;
    const int G = std::min(2, (N + a.group_images - 1) / a.group_images);
    const int OC = C / 8, lanes = std::max(1, 256 / OC);
    if (C % 8 || 256 % OC != 0) return fail("conv3x3_dgrad_bnred: C/8 must divide 256");
    const long long px = static_cast<long long>(a.group_images) * H * W;
    const int gx = static_cast<int>(std::max(1LL, std::min<long long>((px + lanes * 4 - 1) / (lanes * 4), 148 * 4 / G)));
    bn_bwd_px_kernel<bf16, 4, false, false><<<dim3(gx, G), 256, 0, st>>>(a);
    return check_launch("bn_bwd_reduce");
}

static int pick_bn(int cout) { return (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 64); }

static int conv3x3_tc_then_reduce(const bf16* in, long long ldi, int ci_off, int N, int H, int W, int Cin, const bf16* wp,
                                  int Cout, bf16* out, long long ldo, int co_off, int group_images, cudaStream_t st,
                                  const BnRedArgs& red);

// Op = OpBf16: bf16 activations / packed weights; Op = OpTf32: fp32 activations / packed weights read as TF32 (no fused
// BatchNorm-backward reduce, no weight-resident variant: its 9 taps x 2 K chunks would not fit in shared memory).
template <class Op>
static int conv3x3_tc(const typename Op::T* in, long long ldi, int ci_off, int N, int H, int W, int Cin, const typename Op::T* wp,
                      int Cout, typename Op::T* out, long long ldo, int co_off, double* ssum, double* ssq, int group_images,
                      cudaStream_t st, const float* bn_scale = nullptr, const float* bn_shift = nullptr, const BnRedArgs* red = nullptr) {
    constexpr int KC = Op::kKC;
    constexpr int EB = static_cast<int>(sizeof(typename Op::T));
    if (Cin % 64 || Cout % 64) return fail("tc conv needs Cin, Cout multiples of 64 (got %d, %d)", Cin, Cout);
    if (ldo % 8 || co_off % 8) return fail("tc conv output channel stride/offset must be multiples of 8");
    if (Op::kTf32 && red) return fail("conv3x3_dgrad_bnred: bf16 path only");
    PxParams p;
    memset(&p, 0, sizeof(p));
    const int BN = pick_bn(Cout);
    if (H >= 16 && W >= 8 && !getenv("ONET_NO_HALO")) {
        // H-halo kernel: 16 x 8 pixel tiles, activation boxes shared by the three vertical taps
        p.N = N; p.H = H; p.W = W;
        p.TW = 8; p.TH = 16; p.TN = 1; p.log_tw = 3; p.log_th = 4;
        p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.tiles_n = N;
        p.num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
        p.valid_rows = 128;
        p.num_n_tiles = Cout / BN;
        p.ntaps = 9; p.k_chunks = Cin / KC; p.cin = Cin;
        p.epi_mode = EPI_STORE;
        p.out = out; p.ldo = ldo; p.out_coff = co_off;
        p.stat_sum = ssum; p.stat_sq = ssq; p.cout_total = Cout; p.stat_gstride = Cout; p.group_images = group_images > 0 ? group_images : N;
        p.scale = bn_scale; p.shift = bn_shift;
        apply_bnred(p, red, Cout);
        CUtensorMap tA, tB;
        const uint32_t hbox[5] = {KC, 8, 1, 18, 1};
        if (make_act_map(&tA, in + ci_off, Cin, N, H, W, ldi, hbox)) return 1;
        const char* min_kc_env = getenv("ONET_2CTA_MIN_KC");
        const int min_kc = min_kc_env ? atoi(min_kc_env) : 2;
        if (p.num_m_tiles >= 2 && Cin / 64 >= min_kc && !getenv("ONET_NO_2CTA")) {
            // CTA pairs: two pixel tiles per MMA, each CTA stages half of the weight tile
            if (make_map2(&tB, wp, 9ULL * Cin, Cout, KC, BN / 2, EB)) return 1;
            if constexpr (!Op::kTf32) {
                if (red && BN == 256) return launch_halo2_px<256, true>(tA, tB, p, st);
                if (red && BN == 128) return launch_halo2_px<128, true>(tA, tB, p, st);
                if (red) return conv3x3_tc_then_reduce(in, ldi, ci_off, N, H, W, Cin, wp, Cout, out, ldo, co_off, group_images, st, *red);
            }
            if (BN == 256) return launch_halo2_px<256, false, Op>(tA, tB, p, st);
            if (BN == 128) return launch_halo2_px<128, false, Op>(tA, tB, p, st);
            return launch_halo2_px<64, false, Op>(tA, tB, p, st);
        }
        if (make_map2(&tB, wp, 9ULL * Cin, Cout, KC, BN, EB)) return 1;
        if constexpr (!Op::kTf32) {
            if (p.k_chunks == 1 && p.num_n_tiles == 1 && BN <= 128 && p.num_m_tiles >= 4 * sm_count() && !getenv("ONET_NO_BRES")) {
                // Cin = 64, Cout <= 128, many tiles per CTA: weights stay resident in shared memory
                if (red && BN == 64) return launch_halo_res_px<64, true>(tA, tB, p, st);
                if (red) return conv3x3_tc_then_reduce(in, ldi, ci_off, N, H, W, Cin, wp, Cout, out, ldo, co_off, group_images, st, *red);
                if (BN == 128) return launch_halo_res_px<128>(tA, tB, p, st);
                return launch_halo_res_px<64>(tA, tB, p, st);
            }
            if (red) return conv3x3_tc_then_reduce(in, ldi, ci_off, N, H, W, Cin, wp, Cout, out, ldo, co_off, group_images, st, *red);
        }
        if (BN == 256) return launch_halo_px<256, Op>(tA, tB, p, st);
        if (BN == 128) return launch_halo_px<128, Op>(tA, tB, p, st);
        return launch_halo_px<64, Op>(tA, tB, p, st);
    }
    if constexpr (!Op::kTf32) {
        if (red) return conv3x3_tc_then_reduce(in, ldi, ci_off, N, H, W, Cin, wp, Cout, out, ldo, co_off, group_images, st, *red);
    }
    px_tiling(p, N, H, W, (ssum || bn_scale) ? group_images : 0);
    p.num_n_tiles = Cout / BN;
    p.ntaps = 9; p.k_chunks = Cin / KC; p.cin = Cin;
    // same accumulation order as the halo kernels (filter column outer, filter row inner): a pixel gets bit-identical
    // results whichever kernel variant its image size selects (tiled inference relies on it)
    for (int t = 0; t < 9; ++t) {
        const int kw = t / 3, kh = t % 3;
        p.taps[t] = make_int4(0, kw - 1, 0, kh - 1);
        p.tap_w[t] = kh * 3 + kw;
    }
    p.epi_mode = EPI_STORE;
    p.out = out; p.ldo = ldo; p.out_coff = co_off;
    p.stat_sum = ssum; p.stat_sq = ssq; p.cout_total = Cout; p.stat_gstride = Cout; p.group_images = group_images > 0 ? group_images : N;
    p.scale = bn_scale; p.shift = bn_shift;
    CUtensorMap tA, tB;
    const uint32_t box[5] = {KC, static_cast<uint32_t>(p.TW), 1, static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
    if (make_act_map(&tA, in + ci_off, Cin, N, H, W, ldi, box)) return 1;
    if (make_map2(&tB, wp, 9ULL * Cin, Cout, KC, BN, EB)) return 1;
    if (BN == 256) return launch_px<256, Op>(tA, tB, p, st);
    if (BN == 128) return launch_px<128, Op>(tA, tB, p, st);
    return launch_px<64, Op>(tA, tB, p, st);
}

// dgrad without a fused-reduce instantiation: the plain launch, then the standalone reduce pass over its output
static int conv3x3_tc_then_reduce(const bf16* in, long long ldi, int ci_off, int N, int H, int W, int Cin, const bf16* wp,
                                  int Cout, bf16* out, long long ldo, int co_off, int group_images, cudaStream_t st,
                                  const BnRedArgs& red) {
    if (ldo != Cout || co_off != 0) return fail("conv3x3_dgrad_bnred: dense output required");
    if (conv3x3_tc<OpBf16>(in, ldi, ci_off, N, H, W, Cin, wp, Cout, out, ldo, co_off, nullptr, nullptr, group_images, st)) return 1;
    return bnred_standalone(red, out, N, H, W, Cout, group_images, st);
}

// convT fwd on tensor cores: D[px][(tap,co)] = X[px][:] . wf[(tap,co)][:], scatter epilogue
template <class Op>
static int convT_fwd_tc(const typename Op::T* x, long long ldx, int xoff, int N, int H, int W, int Cin, const typename Op::T* wf,
                        const float* bias, int Co, typename Op::T* out, long long ldo, int ooff, int Ho, int Wo, cudaStream_t st) {
    constexpr int KC = Op::kKC;
    constexpr int EB = static_cast<int>(sizeof(typename Op::T));
    if (Cin % 64 || Co % 64) return fail("tc convT needs Cin, Co multiples of 64 (got %d, %d)", Cin, Co);
    if (ldo % 8 || ooff % 8) return fail("tc convT output channel stride/offset must be multiples of 8");
    PxParams p;
    memset(&p, 0, sizeof(p));
    px_tiling(p, N, H, W, 0);
    const int BN = 256;                       // 4*Co is a multiple of 256; a tile spans one or more of the 2x2 positions
    p.num_n_tiles = 4 * Co / BN;
    p.ntaps = 1; p.k_chunks = Cin / KC; p.cin = Cin;
    p.taps[0] = make_int4(0, 0, 0, 0);
    p.tap_w[0] = 0;
    p.epi_mode = EPI_CONVT;
    p.out = out; p.ldo = ldo; p.out_coff = ooff;
    p.bias = bias; p.co_per_tap = Co; p.cout_total = 4 * Co; p.stat_gstride = 4 * Co; p.group_images = N;
    p.Ho = Ho; p.Wo = Wo;
    CUtensorMap tA, tB;
    const uint32_t box[5] = {KC, static_cast<uint32_t>(p.TW), 1, static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
    if (make_act_map(&tA, x + xoff, Cin, N, H, W, ldx, box)) return 1;
    if (make_map2(&tB, wf, Cin, 4ULL * Co, KC, BN, EB)) return 1;
    return launch_px<256, Op>(tA, tB, p, st);
}

// convT dgrad on tensor cores: dX[px][ci] = sum_{tap,co} dO[2px+tap][co] * wd[ci][(tap,co)]
template <class Op>
static int convT_dgrad_tc(const typename Op::T* go, long long ldg, int goff, int N, int H, int W, int Cin, const typename Op::T* wd,
                          int Co, typename Op::T* dx, long long ldd, int doff, int Ho, int Wo, cudaStream_t st) {
    constexpr int KC = Op::kKC;
    constexpr int EB = static_cast<int>(sizeof(typename Op::T));
    if (Cin % 64 || Co % 64) return fail("tc convT dgrad needs Cin, Co multiples of 64 (got %d, %d)", Cin, Co);
    if (ldd % 8 || doff % 8) return fail("tc convT dgrad output channel stride/offset must be multiples of 8");
    PxParams p;
    memset(&p, 0, sizeof(p));
    px_tiling(p, N, H, W, 0);
    const int BN = pick_bn(Cin);
    p.num_n_tiles = Cin / BN;
    p.ntaps = 4; p.k_chunks = Co / KC; p.cin = Co;
    for (int t = 0; t < 4; ++t) {
        p.taps[t] = make_int4((t & 1) * static_cast<int>(ldg), 0, t >> 1, 0);
        p.tap_w[t] = t;
    }
    p.epi_mode = EPI_STORE;
    p.out = dx; p.ldo = ldd; p.out_coff = doff;
    p.cout_total = Cin; p.stat_gstride = Cin; p.group_images = N;
    CUtensorMap tA, tB;
    const uint32_t box[5] = {KC, static_cast<uint32_t>(p.TW), 1, static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
    if (make_up_map(&tA, go + goff, Co, N, H, W, ldg, box, Ho, Wo)) return 1;
    if (make_map2(&tB, wd, 4ULL * Co, Cin, KC, BN, EB)) return 1;
    if (BN == 256) return launch_px<256, Op>(tA, tB, p, st);
    if (BN == 128) return launch_px<128, Op>(tA, tB, p, st);
    return launch_px<64, Op>(tA, tB, p, st);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient tcgen05 launcher
//   dW[m][n][t] (or transposed) = sum_px G[px (-) t][m] * In[px][n]
//   g_is_up: the M-side operand lives on the 2x upsampled grid (transposed conv), taps = 2x2 positions.
// ------------------------------------------------------------------------------------------------
template <int BNW, class Op = OpBf16>
static int launch_wg(const CUtensorMap& tG, const CUtensorMap& tI, const WgParams& p, cudaStream_t st) {
    using Cfg = WgCfg<BNW, Op>;
    static bool attr_set_[kMaxDevices] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_wg_kernel<BNW, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(wg<%d>): %s", BNW, cudaGetErrorString(e));
        attr_set = true;
    }
    const int units = p.ngroups * p.num_m_tiles * p.num_n_tiles * p.ksplit;
    const int grid = std::min(units, sm_count());
    tapgemm_wg_kernel<BNW, Op><<<grid, 192, Cfg::kSmemBytes, st>>>(tG, tI, p);
    return check_launch(Op::kTf32 ? "tapgemm_wg_kernel/tf32" : "tapgemm_wg_kernel", BNW);
}

template <int BNW, int XS = 1>
static int launch_wh(const CUtensorMap& tG, const CUtensorMap& tI, const WhParams& p, cudaStream_t st) {
    using Cfg = WhCfg<BNW, XS>;
    static bool attr_set_[kMaxDevices] = {};
    bool& attr_set = attr_set_[cur_dev()];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad3x3_halo_kernel<BNW, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(wh<%d,%d>): %s", BNW, XS, cudaGetErrorString(e));
        attr_set = true;
    }
    const int units = p.ntypes * p.num_m_tiles * p.num_n_tiles * p.ksplit;
    wgrad3x3_halo_kernel<BNW, XS><<<std::min(units, sm_count()), 192, Cfg::kSmemBytes, st>>>(tG, tI, p);
    return check_launch(XS == 1 ? "wgrad3x3_halo_kernel" : "wgrad3x3_halo_kernel/xs", BNW);
}

// CTA-pair weight gradient (Mc, Nc multiples of 128)
static int wgrad3x3_halo2_tc(const bf16* g, long long ldg, int goff, int Mc, const bf16* in, long long ldi, int ioff, int Nc,
                             int N, int H, int W, float* dw, cudaStream_t st) {
    static int max_clusters_[kMaxDevices] = {};
    int& max_clusters = max_clusters_[cur_dev()];
    if (max_clusters == 0) {
        cudaError_t e = cudaFuncSetAttribute(wgrad3x3_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Wh2Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(wh2): %s", cudaGetErrorString(e));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(sm_count() & ~1);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = Wh2Cfg::kSmemBytes;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, wgrad3x3_halo2_kernel, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = sm_count() / 2;
        }
        max_clusters = std::min(n, sm_count() / 2);
    }
    Wh2Params p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = W;
    p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 7) / 8;
    p.num_px_tiles = p.tiles_w * p.tiles_h * N;
    p.num_m_units = (Mc / 128) * 3;
    p.num_m_pairs = (p.num_m_units + 1) / 2;
    p.num_n_tiles = Nc / 128;
    // split-K over the co-resident clusters
    {
        const int base_units = p.num_m_pairs * p.num_n_tiles;
        long long best_cost = -1;
        int best_per = p.num_px_tiles;
        const int max_ks = std::min(p.num_px_tiles, 8 * max_clusters);
        for (int ks = 1; ks <= max_ks; ++ks) {
            const int per = (p.num_px_tiles + ks - 1) / ks;
            const int ks_eff = (p.num_px_tiles + per - 1) / per;
            const long long units = static_cast<long long>(base_units) * ks_eff;
            const long long waves = (units + max_clusters - 1) / max_clusters;
            const long long cost = waves * (per + 8);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_per = per; }
        }
        p.px_tiles_per_split = best_per;
        p.ksplit = (p.num_px_tiles + best_per - 1) / best_per;
    }
    p.out = dw; p.m_total = Mc; p.n_total = Nc;
    p.ks_slowest = getenv("ONET_WG_KS_FASTEST") ? 0 : 1;
    CUtensorMap tG, tI;
    const uint32_t gbox[5] = {64, 8, 1, 10, 1}, ibox[5] = {64, 8, 1, 8, 1};
    if (make_act_map(&tG, g + goff, Mc, N, H, W, ldg, gbox)) return 1;
    if (make_act_map(&tI, in + ioff, Nc, N, H, W, ldi, ibox)) return 1;
    const int units = p.num_m_pairs * p.num_n_tiles * p.ksplit;
    wgrad3x3_halo2_kernel<<<2 * std::min(units, max_clusters), 192, Wh2Cfg::kSmemBytes, st>>>(tG, tI, p);
    return check_launch("wgrad3x3_halo2_kernel");
}

// 3x3 weight gradient, H-halo variant: dW[co][ci][kh][kw] += sum_px G[px - (kh-1,kw-1)][co] * In[px][ci]
static int wgrad3x3_halo_tc(const bf16* g, long long ldg, int goff, int Mc, const bf16* in, long long ldi, int ioff, int Nc,
                            int N, int H, int W, float* dw, cudaStream_t st) {
    // CTA pairs need an even number of (128-channel block, filter column) units to be worth it: Mc >= 256
    if (Mc % 256 == 0 && Nc % 128 == 0 && !getenv("ONET_NO_2CTA") && !getenv("ONET_NO_2CTA_WGRAD"))
        return wgrad3x3_halo2_tc(g, ldg, goff, Mc, in, ldi, ioff, Nc, N, H, W, dw, st);
    WhParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = W;
    p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 7) / 8;
    p.num_px_tiles = p.tiles_w * p.tiles_h * N;
    int BNW;
    static const bool no_xs = getenv("ONET_WG_NO_XSHIFT") != nullptr;       // A/B: the classic forms below
    const bool xshift = !no_xs && (Mc == 64 || Nc == 64);
    if (xshift) {
        // 64-channel G blocks x 64-channel In blocks, filter column carried by three shifted In boxes (WhCfg, XS = 3)
        BNW = 64;
        p.m_tile_channels = 64; p.num_m_tiles = Mc / 64; p.ntypes = 1;
        WhUnit& u = p.types[0];
        u.nbox = 1; u.box_kw[0] = 1; u.box_ch[0] = 0;                       // one h-halo box, no shift in w
        u.nacc = 2;
        u.acc[0].start_off = 0; u.acc[0].lbo = 1024;                        // rows 0..63 = filter row 2, rows 64..127 = filter row 1
        u.acc[0].tapA = 6; u.acc[0].tapB = 3; u.acc[0].chA = u.acc[0].chB = 0;
        u.acc[1].start_off = 2048; u.acc[1].lbo = 1024;                     // rows 0..63 = filter row 0, rows 64..127 unused
        u.acc[1].tapA = 0; u.acc[1].tapB = -1; u.acc[1].chA = u.acc[1].chB = 0;
    } else if (Mc >= 128) {
        if (Mc % 128) return fail("tc wgrad: M channels %d not a multiple of 128", Mc);
        BNW = (Nc % 128 == 0) ? 128 : 64;
        p.m_tile_channels = 128; p.num_m_tiles = Mc / 128; p.ntypes = 3;
        for (int kw = 0; kw < 3; ++kw) {
            WhUnit& u = p.types[kw];
            u.nbox = 2;
            u.box_kw[0] = u.box_kw[1] = kw; u.box_ch[0] = 0; u.box_ch[1] = 64;
            u.nacc = 3;
            for (int kh = 0; kh < 3; ++kh) {
                WhAcc& a = u.acc[kh];
                a.start_off = (2 - kh) * 1024; a.lbo = kWhBoxBytes;
                a.tapA = a.tapB = kh * 3 + kw; a.chA = 0; a.chB = 64;
            }
        }
    } else {
        BNW = 64;
        p.m_tile_channels = 64; p.num_m_tiles = 1; p.ntypes = 1;
        WhUnit& u = p.types[0];
        u.nbox = 3;
        for (int kw = 0; kw < 3; ++kw) { u.box_kw[kw] = kw; u.box_ch[kw] = 0; }
        u.nacc = 5;
        for (int kw = 0; kw < 3; ++kw) {       // rows 0..63 = tap (kh=2,kw), rows 64..127 = tap (kh=1,kw): 8 box rows apart
            WhAcc& a = u.acc[kw];
            a.start_off = kw * kWhBoxBytes; a.lbo = 1024;
            a.tapA = 6 + kw; a.tapB = 3 + kw; a.chA = a.chB = 0;
        }
        u.acc[3].start_off = 2048; u.acc[3].lbo = kWhBoxBytes;            // (kh=0,kw=0) | (kh=0,kw=1)
        u.acc[3].tapA = 0; u.acc[3].tapB = 1; u.acc[3].chA = u.acc[3].chB = 0;
        u.acc[4].start_off = 2 * kWhBoxBytes + 2048; u.acc[4].lbo = 1024;  // (kh=0,kw=2) | unused
        u.acc[4].tapA = 2; u.acc[4].tapB = -1; u.acc[4].chA = u.acc[4].chB = 0;
    }
    p.num_n_tiles = Nc / BNW;
    pick_ksplit(p.ntypes * p.num_m_tiles * p.num_n_tiles, p.num_px_tiles, &p.ksplit, &p.px_tiles_per_split);
    p.out = dw; p.m_total = Mc; p.n_total = Nc;
    p.ks_slowest = getenv("ONET_WG_KS_FASTEST") ? 0 : 1;
    CUtensorMap tG, tI;
    const uint32_t gbox[5] = {64, 8, 1, 10, 1}, ibox[5] = {64, 8, 1, 8, 1};
    if (make_act_map(&tG, g + goff, Mc, N, H, W, ldg, gbox)) return 1;
    if (make_act_map(&tI, in + ioff, Nc, N, H, W, ldi, ibox)) return 1;
    if (xshift) return launch_wh<64, 3>(tG, tI, p, st);
    if (BNW == 128) return launch_wh<128>(tG, tI, p, st);
    return launch_wh<64>(tG, tI, p, st);
}

template <class Op>
static int wgrad_tc(const typename Op::T* g, long long ldg, int goff, int Mc, bool g_is_up, const typename Op::T* in, long long ldi,
                    int ioff, int Nc, int N, int H, int W, int ntaps, float* dw, bool transposed, cudaStream_t st, int Ho = 0,
                    int Wo = 0) {
    constexpr int KC = Op::kKC;
    if (Mc % 64 || Nc % 64) return fail("tc wgrad needs channel counts multiples of 64 (got %d, %d)", Mc, Nc);
    if constexpr (!Op::kTf32) {
        if (!g_is_up && ntaps == 9 && !transposed && H >= 8 && W >= 8 && !getenv("ONET_NO_HALO"))
            return wgrad3x3_halo_tc(g, ldg, goff, Mc, in, ldi, ioff, Nc, N, H, W, dw, st);
    }
    constexpr int SLAB = Op::kTf32 ? 32 : 64;         // pixels per K slab (WgCfg::kSlabPx)
    WgParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = W;
    p.TW = std::min(p2floor(W), 8);
    p.TH = std::min(p2floor(H), SLAB / p.TW);
    p.TN = SLAB / (p.TW * p.TH);
    p.tiles_w = (W + p.TW - 1) / p.TW;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.tiles_n = (N + p.TN - 1) / p.TN;
    p.num_px_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
        if (g_is_up) p.taps[t] = make_int4((t & 1) * static_cast<int>(ldg), 0, t >> 1, 0);
        else p.taps[t] = make_int4(0, -(t % 3 - 1), 0, -(t / 3 - 1));
    }
    // accumulator groups
    int ng = 0;
    if (Mc >= 128) {
        if (Mc % 128) return fail("tc wgrad: M channels %d not a multiple of 128", Mc);
        p.m_tile_channels = 128;
        p.num_m_tiles = Mc / 128;
        for (int t0 = 0; t0 < ntaps; t0 += kWgMaxAcc) {
            WgGroup& gr = p.groups[ng++];
            gr.nacc = std::min(kWgMaxAcc, ntaps - t0);
            for (int a = 0; a < gr.nacc; ++a) { gr.tapA[a] = gr.tapB[a] = t0 + a; gr.offA[a] = 0; gr.offB[a] = 64; }
        }
    } else {
        p.m_tile_channels = 64;
        p.num_m_tiles = 1;
        const int npairs = (ntaps + 1) / 2;
        for (int p0 = 0; p0 < npairs; p0 += kWgMaxAcc) {
            WgGroup& gr = p.groups[ng++];
            gr.nacc = std::min(kWgMaxAcc, npairs - p0);
            for (int a = 0; a < gr.nacc; ++a) {
                const int ta = 2 * (p0 + a), tb = ta + 1;
                gr.tapA[a] = ta; gr.offA[a] = 0;
                gr.tapB[a] = tb < ntaps ? tb : -1; gr.offB[a] = 0;
            }
        }
    }
    p.ngroups = ng;
    const int BNW = (Nc % 128 == 0) ? 128 : 64;
    p.num_n_tiles = Nc / BNW;
    pick_ksplit(p.ngroups * p.num_m_tiles * p.num_n_tiles, p.num_px_tiles, &p.ksplit, &p.px_tiles_per_split);
    p.out = dw;
    p.m_total = Mc; p.n_total = Nc; p.out_transposed = transposed ? 1 : 0;
    p.ks_slowest = getenv("ONET_WG_KS_FASTEST") ? 0 : 1;
    CUtensorMap tG, tI;
    const uint32_t box[5] = {KC, static_cast<uint32_t>(p.TW), 1, static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
    if (g_is_up) { if (make_up_map(&tG, g + goff, Mc, N, H, W, ldg, box, Ho, Wo, true)) return 1; }
    else { if (make_act_map(&tG, g + goff, Mc, N, H, W, ldg, box, true)) return 1; }
    if (make_act_map(&tI, in + ioff, Nc, N, H, W, ldi, box, true)) return 1;
    if (BNW == 128) return launch_wg<128, Op>(tG, tI, p, st);
    return launch_wg<64, Op>(tG, tI, p, st);
}

// ------------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------------
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int onet_version(void) { return 100; }
int64_t onet_launch_count(void) { return g_launches; }
const char* onet_last_kernel(void) { return g_last_kernel; }
const char* onet_last_error(void) { return g_err; }

int onet_set_splitk_workspace(float* ws, int64_t nfloats) {
    const int dev = cur_dev();
    g_splitk_ws[dev] = (ws != nullptr && nfloats > 0) ? ws : nullptr;
    g_splitk_floats[dev] = g_splitk_ws[dev] ? nfloats : 0;
    return 0;
}

int onet_device_info(int* smc, int* major, int* minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("no CUDA device");
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
    if (smc) *smc = pr.multiProcessorCount;
    if (major) *major = pr.major;
    if (minor) *minor = pr.minor;
    return 0;
}

int onet_prep_input(const float* x, int B, int Cin, int H, int W, float bias, void* out, int dtype, void* stream) {
    const long long total = static_cast<long long>(B) * H * W * Cin;
    if (dtype == ONET_F32)
        prep_input_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(x, B, Cin, static_cast<long long>(H) * W, bias, static_cast<float*>(out));
    else
        prep_input_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(x, B, Cin, static_cast<long long>(H) * W, bias, static_cast<bf16*>(out));
    return check_launch("prep_input");
}

int onet_pack_conv_weights(const float* w, int Cout, int Cin, void* wf, void* wd, int dtype, void* stream) {
    const long long total = 9LL * Cout * Cin;
    if (dtype == ONET_F32)
        pack_conv_w_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(w, Cout, Cin, static_cast<float*>(wf), static_cast<float*>(wd));
    else
        pack_conv_w_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(w, Cout, Cin, static_cast<bf16*>(wf), static_cast<bf16*>(wd));
    return check_launch("pack_conv_weights");
}

int onet_pack_convT_weights(const float* w, int Cin, int Cout, void* wf, void* wd, int dtype, void* stream) {
    const long long total = 4LL * Cout * Cin;
    if (dtype == ONET_F32)
        pack_convT_w_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(w, Cin, Cout, static_cast<float*>(wf), static_cast<float*>(wd));
    else
        pack_convT_w_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(w, Cin, Cout, static_cast<bf16*>(wf), static_cast<bf16*>(wd));
    return check_launch("pack_convT_weights");
}

int onet_pack_all_weights(int n, const float* const* w, const int* d0, const int* d1, const int* taps, void* const* wf,
                          void* const* wd, int dtype, void* stream) {
    if (n < 1 || n > kPackMaxLayers) return fail("pack_all_weights: between 1 and %d layers per call", kPackMaxLayers);
    PackJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.n = n;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        if (taps[i] != 9 && taps[i] != 4) return fail("pack_all_weights: taps must be 9 (conv) or 4 (convT)");
        if (taps[i] == 4 && wd[i] == nullptr) return fail("pack_all_weights: convT needs both packed layouts");
        PackJob& j = jobs.job[i];
        j.w = w[i]; j.wf = wf[i]; j.wd = wd[i]; j.d0 = d0[i]; j.d1 = d1[i]; j.taps = taps[i];
        j.tile_begin = tiles;
        tiles += ((d0[i] + 31) / 32) * ((d1[i] + 31) / 32);
    }
    if (dtype == ONET_F32) pack_all_weights_kernel<float><<<tiles, 256, 0, ST(stream)>>>(jobs);
    else pack_all_weights_kernel<bf16><<<tiles, 256, 0, ST(stream)>>>(jobs);
    return check_launch("pack_all_weights");
}

int onet_conv3x3_fwd(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                     void* out, int64_t ldo, int co_off, double* stat_sum, double* stat_sq, int group_images,
                     int dtype, int engine, void* stream) {
    if (N <= 0 || H <= 0 || W <= 0) return fail("conv3x3_fwd: empty tensor");
    if (engine == ONET_ENGINE_TC) {     // tensor cores: bf16 operands, or (dtype ONET_F32) fp32 operands read as TF32
        if (dtype == ONET_F32)
            return conv3x3_tc<OpTf32>(static_cast<const float*>(in), ldi, ci_off, N, H, W, Cin, static_cast<const float*>(wp), Cout,
                                      static_cast<float*>(out), ldo, co_off, stat_sum, stat_sq, group_images, ST(stream));
        return conv3x3_tc<OpBf16>(static_cast<const bf16*>(in), ldi, ci_off, N, H, W, Cin, static_cast<const bf16*>(wp), Cout,
                                  static_cast<bf16*>(out), ldo, co_off, stat_sum, stat_sq, group_images, ST(stream));
    }
    const long long M = static_cast<long long>(N) * H * W;
    dim3 grid(static_cast<unsigned>((M + 63) / 64), (Cout + 63) / 64);
    const int gi = group_images > 0 ? group_images : N;
    if ((Cin == 1 || Cin == 3) && Cout == 64 && ldi == Cin && ci_off == 0 && ldo == 64 && co_off == 0 && W % 4 == 0) {
        // first layer of the U-Net: direct bandwidth-bound kernel
        constexpr int ROWS = 32;
        const int wgb = (W / 4 + 31) / 32, chunks = (H + ROWS - 1) / ROWS;
        const unsigned fg = static_cast<unsigned>(N) * chunks * wgb;
#define ONET_FIRST(TT, CC)                                                                                               \
    conv_first_fwd_rows_kernel<TT, CC, ROWS><<<fg, 256, 0, ST(stream)>>>(static_cast<const TT*>(in), N, H, W,            \
                                                                         static_cast<const TT*>(wp), static_cast<TT*>(out), \
                                                                         stat_sum, stat_sq, gi)
        if (dtype == ONET_F32) { if (Cin == 1) ONET_FIRST(float, 1); else ONET_FIRST(float, 3); }
        else { if (Cin == 1) ONET_FIRST(bf16, 1); else ONET_FIRST(bf16, 3); }
#undef ONET_FIRST
        return check_launch("conv_first_fwd");
    }
    if (dtype == ONET_F32)
        conv3x3_simt_kernel<float><<<grid, 256, 0, ST(stream)>>>(static_cast<const float*>(in), ldi, ci_off, N, H, W, Cin,
                                                                 static_cast<const float*>(wp), Cout, static_cast<float*>(out),
                                                                 ldo, co_off, stat_sum, stat_sq, gi);
    else
        conv3x3_simt_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16*>(in), ldi, ci_off, N, H, W, Cin,
                                                                static_cast<const bf16*>(wp), Cout, static_cast<bf16*>(out), ldo,
                                                                co_off, stat_sum, stat_sq, gi);
    return check_launch("conv3x3_simt");
}

// ---- first convolution of the U-Net without materialising its output (first_layer.cuh)
// ONET_NO_FIRST_MMA=1: A/B switch back to the CUDA-core form of the 1-channel bf16 first conv (first_layer.cuh)
static bool first_mma_disabled() {
    static const bool off = getenv("ONET_NO_FIRST_MMA") != nullptr;
    return off;
}

static int first_layer_check(const char* what, int N, int H, int W, int Cin) {
    if (N <= 0 || H <= 0 || W <= 0) return fail("%s: empty tensor", what);
    if (Cin != 1 && Cin != 3) return fail("%s: in_chns must be 1 or 3 (got %d)", what, Cin);
    if (W % 4) return fail("%s: width must be a multiple of 4 (got %d); use onet_conv3x3_fwd + onet_bn_relu_apply", what, W);
    return 0;
}

int onet_first_conv_stats(const void* x, int N, int H, int W, int Cin, const void* wp, double* gram, double* stat_sum,
                          double* stat_sq, int group_images, int dtype, void* stream) {
    if (first_layer_check("first_conv_stats", N, H, W, Cin)) return 1;
    if (stat_sum == nullptr || stat_sq == nullptr) return fail("first_conv_stats: stat_sum and stat_sq are required");
    constexpr int ROWS = 32;
    const int gi = group_images > 0 ? group_images : N;
    const int chunks = (H + ROWS - 1) / ROWS;
    if (gram != nullptr) {
        // closed form: patch moments S, G per statistics group, then sum y = w.S, sum y^2 = w^T G w (first_layer.cuh)
        const int G = std::min(2, (N + gi - 1) / gi);
        const int K = 9 * Cin;
        if (Cin == 1) {
            const int WG = W / 4;
            const int threads = WG <= 64 ? 64 : (WG <= 128 ? 128 : 256);
            const int wgb = (WG + 255) / 256;
            const unsigned fg = static_cast<unsigned>(N) * chunks * wgb;
            if (dtype == ONET_F32) first_gram_kernel<float, ROWS><<<fg, threads, 0, ST(stream)>>>(static_cast<const float*>(x), N, H, W, gi, gram);
            else first_gram_kernel<bf16, ROWS><<<fg, threads, 0, ST(stream)>>>(static_cast<const bf16*>(x), N, H, W, gi, gram);
        } else {
            // in_chns = 3: 27 + 27 x 27 moments on warp-level MMAs (first_layer_mma.cuh); bf16 storage only (exact operands)
            if (dtype == ONET_F32 || first_mma_disabled()) return fail("first_conv_stats: patch moments for in_chns = 3 exist for bf16 storage only");
            const unsigned mg = static_cast<unsigned>(N) * ((H + kFmRows - 1) / kFmRows) * ((W + kFmCols - 1) / kFmCols);
            first_gram_mma_kernel<3><<<mg, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), N, H, W, gi, gram);
        }
        if (check_launch("first_gram")) return 1;
        if (dtype == ONET_F32) first_stats_from_gram_kernel<float><<<G, 64, 0, ST(stream)>>>(static_cast<const float*>(wp), gram, G, K, stat_sum, stat_sq);
        else first_stats_from_gram_kernel<bf16><<<G, 64, 0, ST(stream)>>>(static_cast<const bf16*>(wp), gram, G, K, stat_sum, stat_sq);
        return check_launch("first_stats_from_gram");
    }
    const int wgb = (W / 4 + 31) / 32;
    const unsigned fg = static_cast<unsigned>(N) * chunks * wgb;
#define ONET_FIRST_FWD(TT, CC, MODE, RND, OUT)                                                                                     \
    first_conv_fwd_kernel<TT, CC, ROWS, MODE, RND><<<fg, 256, 0, ST(stream)>>>(static_cast<const TT*>(x), N, H, W, static_cast<const TT*>(wp), \
                                                                               stat_sum, stat_sq, scale_, shift_, gi, static_cast<TT*>(OUT))
    const float* scale_ = nullptr;
    const float* shift_ = nullptr;
    if (dtype == ONET_F32) { if (Cin == 1) ONET_FIRST_FWD(float, 1, FIRST_STATS, true, nullptr); else ONET_FIRST_FWD(float, 3, FIRST_STATS, true, nullptr); }
    else { if (Cin == 1) ONET_FIRST_FWD(bf16, 1, FIRST_STATS, true, nullptr); else ONET_FIRST_FWD(bf16, 3, FIRST_STATS, true, nullptr); }
    return check_launch("first_conv_stats");
}

int onet_first_conv_bn_relu(const void* x, int N, int H, int W, int Cin, const void* wp, const float* scale, const float* shift,
                            int group_images, void* out, int round_y, int dtype, void* stream) {
    if (first_layer_check("first_conv_bn_relu", N, H, W, Cin)) return 1;
    if (scale == nullptr || shift == nullptr || out == nullptr) return fail("first_conv_bn_relu: scale, shift and out are required");
    constexpr int ROWS = 32;
    const int wgb = (W / 4 + 31) / 32, chunks = (H + ROWS - 1) / ROWS;
    const unsigned fg = static_cast<unsigned>(N) * chunks * wgb;
    const int gi = group_images > 0 ? group_images : N;
    double* stat_sum = nullptr;
    double* stat_sq = nullptr;
    const float* scale_ = scale;
    const float* shift_ = shift;
    if (dtype == ONET_F32) {        // fp32 storage: rounding is the identity
        if (Cin == 1) ONET_FIRST_FWD(float, 1, FIRST_APPLY, false, out); else ONET_FIRST_FWD(float, 3, FIRST_APPLY, false, out);
    } else if (round_y) {
        if (Cin == 1) ONET_FIRST_FWD(bf16, 1, FIRST_APPLY, true, out); else ONET_FIRST_FWD(bf16, 3, FIRST_APPLY, true, out);
    } else if (!first_mma_disabled()) {
        // unrounded y: warp-level tensor-core form (first_layer_mma.cuh), any width
        const unsigned mg = static_cast<unsigned>(N) * ((H + kFmRows - 1) / kFmRows) * ((W + kFmCols - 1) / kFmCols);
        if (Cin == 1)
            first_mma_fwd_kernel<1><<<mg, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), N, H, W, static_cast<const bf16*>(wp), scale, shift, gi,
                                                                static_cast<bf16*>(out));
        else
            first_mma_fwd_kernel<3><<<mg, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), N, H, W, static_cast<const bf16*>(wp), scale, shift, gi,
                                                                static_cast<bf16*>(out));
    } else {
        if (Cin == 1) ONET_FIRST_FWD(bf16, 1, FIRST_APPLY, false, out); else ONET_FIRST_FWD(bf16, 3, FIRST_APPLY, false, out);
    }
#undef ONET_FIRST_FWD
    return check_launch("first_conv_bn_relu");
}

int onet_first_conv_bwd(const void* x, int N, int H, int W, int Cin, const void* wp, const float* scale, const float* shift,
                        const float* mean, const float* invstd, int group_images, const void* g, const double* gram, float* acc_a,
                        double* sums, double count, float* dw, float* dgamma0, float* dbeta0, float* dgamma1, float* dbeta1,
                        int dtype, void* stream) {
    if (first_layer_check("first_conv_bwd", N, H, W, Cin)) return 1;
    if (g == nullptr || sums == nullptr || dw == nullptr) return fail("first_conv_bwd: g, sums and dw are required");
    constexpr int ROWS = 32;
    if (gram != nullptr) {
        // single pass over g + closed-form assembly from the patch moments of the forward pass (first_layer.cuh)
        if (acc_a == nullptr) return fail("first_conv_bwd: acc_a (zeroed float [G][64][9 * in_chns]) is required with gram");
        if (Cin == 3 && (dtype == ONET_F32 || first_mma_disabled())) return fail("first_conv_bwd: the closed form for in_chns = 3 exists for bf16 storage only");
        FirstFusedArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.N = N; fa.H = H; fa.W = W; fa.group_images = group_images > 0 ? group_images : N;
        fa.scale = scale; fa.shift = shift; fa.mean = mean; fa.invstd = invstd; fa.sums = sums; fa.acc_a = acc_a;
        const int G = std::min(2, (N + fa.group_images - 1) / fa.group_images);
        const int lanes = 16, wgb = (W / 4 + lanes - 1) / lanes, chunks = (H + ROWS - 1) / ROWS;
        const unsigned gx = static_cast<unsigned>(N) * chunks * wgb;
        const long long numel = 64LL * 9 * Cin;
        fa.partial = (dtype == ONET_F32) ? splitk_ws(static_cast<long long>(gx) * numel) : nullptr;
        const bool mma = dtype != ONET_F32 && !first_mma_disabled();
        if (dtype == ONET_F32) {
            first_conv_bwd_fused_kernel<float, 4, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const float*>(x), static_cast<const float*>(wp), static_cast<const float*>(g), fa);
            if (check_launch("first_conv_bwd_fused")) return 1;
        } else if (!mma) {
            first_conv_bwd_fused_kernel<bf16, 4, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(wp), static_cast<const bf16*>(g), fa);
            if (check_launch("first_conv_bwd_fused")) return 1;
        } else {
            // bf16: A and s1 by warp-level MMAs, s2 derived from A in the assembly kernel (first_layer_mma.cuh)
            const unsigned mg = static_cast<unsigned>(N) * ((H + kFmRows - 1) / kFmRows) * ((W + kFmCols - 1) / kFmCols);
            if (Cin == 1) first_mma_bwd_kernel<1><<<mg, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(wp), static_cast<const bf16*>(g), fa);
            else first_mma_bwd_kernel<3><<<mg, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(wp), static_cast<const bf16*>(g), fa);
            if (check_launch("first_mma_bwd")) return 1;
        }
        if (fa.partial != nullptr) {       // deterministic: per-block partials added in block order, group by group
            const long long per_img = static_cast<long long>(chunks) * wgb;
            const long long b0 = std::min<long long>(fa.group_images, N) * per_img;
            splitk_reduce(fa.partial, static_cast<int>(b0), numel, acc_a, ST(stream));
            if (check_launch("splitk_reduce")) return 1;
            if (G > 1) {
                splitk_reduce(fa.partial + b0 * numel, static_cast<int>(gx - b0), numel, acc_a + numel, ST(stream));
                if (check_launch("splitk_reduce")) return 1;
            }
        }
        if (dtype == ONET_F32)
            first_bwd_assemble_kernel<float><<<(64 * 9 * Cin + 127) / 128, 128, 0, ST(stream)>>>(static_cast<const float*>(wp), gram, acc_a, sums, scale, mean, invstd, G, count, dw, 0, Cin);
        else
            first_bwd_assemble_kernel<bf16><<<(64 * 9 * Cin + 127) / 128, 128, 0, ST(stream)>>>(static_cast<const bf16*>(wp), gram, acc_a, sums, scale, mean, invstd, G, count, dw, mma ? 1 : 0, Cin);
        if (check_launch("first_bwd_assemble")) return 1;
        if (dgamma0 != nullptr) {
            bn_param_grad_kernel<<<1, 64, 0, ST(stream)>>>(sums, G, 64, dgamma0, dbeta0, dgamma1 ? dgamma1 : dgamma0, dbeta1 ? dbeta1 : dbeta0);
            if (check_launch("bn_param_grad")) return 1;
        }
        return 0;
    }
    FirstBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.H = H; a.W = W; a.group_images = group_images > 0 ? group_images : N;
    a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd;
    a.sums = sums; a.count = count; a.dw = dw;
    const int G = std::min(2, (N + a.group_images - 1) / a.group_images);
    const int lanes = Cin == 1 ? 16 : 4;                    // column groups per block = 256 / (64 / CPT)
    const int wgb = (W / 4 + lanes - 1) / lanes, chunks = (H + ROWS - 1) / ROWS;
    const unsigned gx = static_cast<unsigned>(N) * chunks * wgb;
    const long long numel = 64LL * 9 * Cin;
    a.partial = (dtype == ONET_F32) ? splitk_ws(static_cast<long long>(gx) * numel) : nullptr;
#define ONET_FIRST_BWD(KERNEL, TT, CC, CPT)                                                                                       \
    KERNEL<TT, CC, CPT, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const TT*>(x), static_cast<const TT*>(wp), static_cast<const TT*>(g), a)
    if (dtype == ONET_F32) { if (Cin == 1) ONET_FIRST_BWD(first_conv_bwd_reduce_kernel, float, 1, 4); else ONET_FIRST_BWD(first_conv_bwd_reduce_kernel, float, 3, 1); }
    else { if (Cin == 1) ONET_FIRST_BWD(first_conv_bwd_reduce_kernel, bf16, 1, 4); else ONET_FIRST_BWD(first_conv_bwd_reduce_kernel, bf16, 3, 1); }
    if (check_launch("first_conv_bwd_reduce")) return 1;
    if (dtype == ONET_F32) { if (Cin == 1) ONET_FIRST_BWD(first_conv_bwd_wgrad_kernel, float, 1, 4); else ONET_FIRST_BWD(first_conv_bwd_wgrad_kernel, float, 3, 1); }
    else { if (Cin == 1) ONET_FIRST_BWD(first_conv_bwd_wgrad_kernel, bf16, 1, 4); else ONET_FIRST_BWD(first_conv_bwd_wgrad_kernel, bf16, 3, 1); }
#undef ONET_FIRST_BWD
    if (check_launch("first_conv_bwd_wgrad")) return 1;
    if (a.partial != nullptr) {
        splitk_reduce(a.partial, static_cast<int>(gx), numel, dw, ST(stream));
        if (check_launch("splitk_reduce")) return 1;
    }
    if (dgamma0 != nullptr) {
        bn_param_grad_kernel<<<1, 64, 0, ST(stream)>>>(sums, G, 64, dgamma0, dbeta0, dgamma1 ? dgamma1 : dgamma0, dbeta1 ? dbeta1 : dbeta0);
        if (check_launch("bn_param_grad")) return 1;
    }
    return 0;
}

int onet_conv3x3_bn_relu_infer(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                               const float* scale, const float* shift, int group_images, void* out, int64_t ldo, int co_off,
                               int dtype, int engine, void* stream) {
    if (N <= 0 || H <= 0 || W <= 0) return fail("conv3x3_bn_relu_infer: empty tensor");
    if (engine != ONET_ENGINE_TC) return fail("conv3x3_bn_relu_infer: tensor-core engine only");
    if (scale == nullptr || shift == nullptr) return fail("conv3x3_bn_relu_infer: scale and shift are required");
    if (dtype == ONET_F32)
        return conv3x3_tc<OpTf32>(static_cast<const float*>(in), ldi, ci_off, N, H, W, Cin, static_cast<const float*>(wp), Cout,
                                  static_cast<float*>(out), ldo, co_off, nullptr, nullptr, group_images, ST(stream), scale, shift);
    return conv3x3_tc<OpBf16>(static_cast<const bf16*>(in), ldi, ci_off, N, H, W, Cin, static_cast<const bf16*>(wp), Cout,
                              static_cast<bf16*>(out), ldo, co_off, nullptr, nullptr, group_images, ST(stream), scale, shift);
}

int onet_maxpool2x2(const void* in, int64_t ldi, int ioff, int N, int H, int W, int C, void* out, int dtype, void* stream) {
    if (C % 8 || ldi % 8 || ioff % 8) return fail("maxpool2x2: channel counts/offsets must be multiples of 8");
    const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return 0;
    if (dtype == ONET_F32)
        maxpool2x2_kernel<float><<<grid_for(total, 256, 148 * 16), 256, 0, ST(stream)>>>(static_cast<const float*>(in), ldi, ioff, N, H, W, C, static_cast<float*>(out));
    else
        maxpool2x2_kernel<bf16><<<grid_for(total, 256, 148 * 16), 256, 0, ST(stream)>>>(static_cast<const bf16*>(in), ldi, ioff, N, H, W, C, static_cast<bf16*>(out));
    return check_launch("maxpool2x2");
}

int onet_conv3x3_wgrad(const void* g, int64_t ldg, int g_off, const void* in, int64_t ldi, int ci_off, int N, int H,
                       int W, int Cin, int Cout, float* dw, int dtype, int engine, void* stream) {
    if (engine == ONET_ENGINE_TC) {
        if (dtype == ONET_F32)
            return wgrad_tc<OpTf32>(static_cast<const float*>(g), ldg, g_off, Cout, false, static_cast<const float*>(in), ldi, ci_off,
                                    Cin, N, H, W, 9, dw, false, ST(stream));
        return wgrad_tc<OpBf16>(static_cast<const bf16*>(g), ldg, g_off, Cout, false, static_cast<const bf16*>(in), ldi, ci_off, Cin,
                                N, H, W, 9, dw, false, ST(stream));
    }
    const long long M = static_cast<long long>(N) * H * W;
    if ((Cin == 1 || Cin == 3) && Cout == 64 && ldi == Cin && ci_off == 0 && ldg == 64 && g_off == 0 && W % 4 == 0) {
        constexpr int ROWS = 32;
        const int lanes = Cin == 1 ? 32 : 8;       // column groups per block = 256 / (64 / CPT)
        const int wgb = (W / 4 + lanes - 1) / lanes, chunks = (H + ROWS - 1) / ROWS;
        const unsigned gx = static_cast<unsigned>(N) * chunks * wgb;
        const long long numel = 64LL * 9 * Cin;
        float* part = (dtype == ONET_F32) ? splitk_ws(static_cast<long long>(gx) * numel) : nullptr;
        if (dtype == ONET_F32) {
            if (Cin == 1) conv_first_wgrad_rows_kernel<float, 1, 8, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const float*>(g), static_cast<const float*>(in), N, H, W, dw, part);
            else conv_first_wgrad_rows_kernel<float, 3, 2, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const float*>(g), static_cast<const float*>(in), N, H, W, dw, part);
        } else {
            if (Cin == 1) conv_first_wgrad_rows_kernel<bf16, 1, 8, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(in), N, H, W, dw, part);
            else conv_first_wgrad_rows_kernel<bf16, 3, 2, ROWS><<<gx, 256, 0, ST(stream)>>>(static_cast<const bf16*>(g), static_cast<const bf16*>(in), N, H, W, dw, part);
        }
        if (check_launch("conv_first_wgrad")) return 1;
        if (part != nullptr) {
            splitk_reduce(part, static_cast<int>(gx), numel, dw, ST(stream));
            return check_launch("splitk_reduce");
        }
        return 0;
    }
    const int K = 9 * Cin;
    const int bx = (Cout + 63) / 64, by = (K + 63) / 64;
    int splits = std::max(1, (4 * sm_count()) / (bx * by));
    long long per = (M + splits - 1) / splits;
    per = std::max<long long>(16, (per + 15) / 16 * 16);
    splits = static_cast<int>((M + per - 1) / per);
    dim3 grid(bx, by, splits);
    const long long numel = 9LL * Cin * Cout;
    // one split: a single add per address, already deterministic
    float* part = (dtype == ONET_F32 && splits > 1) ? splitk_ws(static_cast<long long>(splits) * numel) : nullptr;
    if (dtype == ONET_F32)
        conv3x3_wgrad_simt_kernel<float><<<grid, 256, 0, ST(stream)>>>(static_cast<const float*>(g), ldg, g_off,
                                                                       static_cast<const float*>(in), ldi, ci_off, N, H, W, Cin, Cout, dw, per, part);
    else
        conv3x3_wgrad_simt_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16*>(g), ldg, g_off,
                                                                      static_cast<const bf16*>(in), ldi, ci_off, N, H, W, Cin, Cout, dw, per, part);
    if (check_launch("conv3x3_wgrad_simt")) return 1;
    if (part != nullptr) {
        splitk_reduce(part, splits, numel, dw, ST(stream));
        return check_launch("splitk_reduce");
    }
    return 0;
}

int onet_bn_finalize(const double* stat_sum, const double* stat_sq, int G, int C, double count, const float* gamma0,
                     const float* beta0, float* running_mean0, float* running_var0, const float* gamma1,
                     const float* beta1, float* running_mean1, float* running_var1, float momentum, float* mean,
                     float* invstd, float* scale, float* shift, void* stream) {
    if (G < 1 || G > 2) return fail("bn_finalize: G must be 1 or 2");
    BnGroupPtrs p;
    p.gamma[0] = gamma0; p.beta[0] = beta0; p.running_mean[0] = running_mean0; p.running_var[0] = running_var0;
    p.gamma[1] = gamma1; p.beta[1] = beta1; p.running_mean[1] = running_mean1; p.running_var[1] = running_var1;
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(stat_sum, stat_sq, G, C, count, p, momentum, mean, invstd, scale, shift);
    return check_launch("bn_finalize");
}

int onet_bn_eval_prepare(int G, int C, const float* gamma0, const float* beta0, const float* running_mean0,
                         const float* running_var0, const float* gamma1, const float* beta1,
                         const float* running_mean1, const float* running_var1, float* scale, float* shift,
                         void* stream) {
    if (G < 1 || G > 2) return fail("bn_eval_prepare: G must be 1 or 2");
    BnGroupPtrs p;
    p.gamma[0] = gamma0; p.beta[0] = beta0;
    p.running_mean[0] = const_cast<float*>(running_mean0); p.running_var[0] = const_cast<float*>(running_var0);
    p.gamma[1] = gamma1; p.beta[1] = beta1;
    p.running_mean[1] = const_cast<float*>(running_mean1); p.running_var[1] = const_cast<float*>(running_var1);
    bn_eval_prepare_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(G, C, p, scale, shift);
    return check_launch("bn_eval_prepare");
}

int onet_bn_relu_apply(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                       int group_images, void* out, int64_t ldo, int ooff, void* pool, void* pool_arg, int dtype, void* stream) {
    if (C % 8 || ldo % 8 || ooff % 8) return fail("bn_relu_apply: channel counts/offsets must be multiples of 8");
    const long long total = static_cast<long long>(N) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    if (total >= (1LL << 30)) return fail("bn_relu_apply: %lld (window, channel octet) items exceed the 32-bit index range", total);
    const int gi = group_images > 0 ? group_images : N;
    constexpr int UNR = 2;
    const int grid = grid_for((total + UNR - 1) / UNR, 256, 148 * 24);
    const FastDiv fd_oc = make_fastdiv(C / 8), fd_w2 = make_fastdiv((W + 1) / 2), fd_h2 = make_fastdiv((H + 1) / 2);
    if (dtype == ONET_F32)
        bn_relu_apply_kernel<float, UNR><<<grid, 256, 0, ST(stream)>>>(
            static_cast<const float*>(y), N, H, W, C, scale, shift, gi, static_cast<float*>(out), ldo, ooff, static_cast<float*>(pool),
            static_cast<unsigned short*>(pool_arg), fd_oc, fd_w2, fd_h2);
    else
        bn_relu_apply_kernel<bf16, UNR><<<grid, 256, 0, ST(stream)>>>(
            static_cast<const bf16*>(y), N, H, W, C, scale, shift, gi, static_cast<bf16*>(out), ldo, ooff, static_cast<bf16*>(pool),
            static_cast<unsigned short*>(pool_arg), fd_oc, fd_w2, fd_h2);
    return check_launch("bn_relu_apply");
}

}  // extern "C"

template <typename T>
static int bn_bwd_impl(const void* y, int N, int H, int W, int C, const float* scale, const float* shift, const float* mean,
                       const float* invstd, int group_images, const void* g1, int64_t ld1, int off1, const void* g2,
                       int64_t ld2, int off2, const void* gp, const void* gp_arg, double* sums, double count, void* dy,
                       float* dgamma0, float* dbeta0, float* dgamma1, float* dbeta1, cudaStream_t st, bool prereduced = false,
                       const float* g1_scale = nullptr, const float* dl_add = nullptr, void* dl_out = nullptr) {
    BnBwdArgs<T> a;
    a.g1_scale = g1_scale; a.dl_add = dl_add; a.dl_out = static_cast<T*>(dl_out);
    a.y = static_cast<const T*>(y); a.N = N; a.H = H; a.W = W; a.C = C;
    a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd;
    a.group_images = group_images > 0 ? group_images : N;
    a.g1 = static_cast<const T*>(g1); a.ld1 = ld1; a.off1 = off1;
    a.g2 = static_cast<const T*>(g2); a.ld2 = ld2; a.off2 = off2;
    a.gp = static_cast<const T*>(gp);
    a.gp_arg = static_cast<const unsigned short*>(gp_arg);
    a.sums = sums; a.count = count; a.dy = static_cast<T*>(dy);
    const int G = std::min(2, (N + a.group_images - 1) / a.group_images);
    const int OC = C / 8, lanes = std::max(1, 256 / OC);
    // The weight-gradient kernels of the previous layer run next to these kernels on a second stream (model.py): ask for the
    // largest shared-memory carve-out so that an SM already holding BatchNorm CTAs can still take a 171 KB wgrad CTA
    // (these kernels stream through L1 and do not need it).  ONET_BN_BWD_BLOCKS overrides the grid (A/B measurements).
    static bool carveout_set_[kMaxDevices] = {};
    bool& carveout_set = carveout_set_[cur_dev()];
    if (!carveout_set) {
        cudaFuncSetAttribute(bn_bwd_win_kernel<T, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_win_kernel<T, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_win_kernel<T, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_win_kernel<T, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_px_kernel<T, 4, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_px_kernel<T, 4, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_px_kernel<T, 4, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(bn_bwd_px_kernel<T, 4, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaGetLastError();
        carveout_set = true;
    }
    static int blocks_override = -1;
    if (blocks_override < 0) {
        const char* e = getenv("ONET_BN_BWD_BLOCKS");
        blocks_override = e ? atoi(e) : 0;
    }
    const int max_blocks = (blocks_override > 0 ? blocks_override : 148 * 4) / G;        // 2 resident blocks per SM, two rounds
    if (prereduced && gp != nullptr) return fail("bn_relu_bwd_apply: a pooled gradient source cannot be pre-reduced");
    if (gp != nullptr) {
        if (OC > kBnWinThreads) return fail("bn_relu_bwd: the pooled variant supports C <= %d", 8 * kBnWinThreads);
        const int wlanes = kBnWinThreads / OC;
        const long long wins = static_cast<long long>(a.group_images) * ((H + 1) / 2) * ((W + 1) / 2);
        if (static_cast<long long>(N) * ((H + 1) / 2) * ((W + 1) / 2) >= (1LL << 30)) return fail("bn_relu_bwd: too many windows for the 32-bit index range");
        a.fd_w2 = make_fastdiv((W + 1) / 2); a.fd_h2 = make_fastdiv((H + 1) / 2);
        // 3 resident blocks of 128 threads per SM; 8 blocks per SM in total: with exactly two full waves (148 * 6) every block of a
        // wave reached its block reduction at the same time and the memory pipes idled twice - 148 * 8 staggers them
        // (pooled C=64 @256x256 1.27 -> 1.13 ms, C=128 @128x128 0.66 -> 0.58 ms; tools/gpu_session_r2t.sh)
        const int wmax = (blocks_override > 0 ? blocks_override : 148 * 8) / G;
        const int gx = static_cast<int>(std::max(1LL, std::min<long long>((wins + wlanes - 1) / wlanes, wmax)));
        if (g2 != nullptr) {
            bn_bwd_win_kernel<T, true, false><<<dim3(gx, G), kBnWinThreads, 0, st>>>(a);
            if (check_launch("bn_bwd_reduce")) return 1;
            bn_bwd_win_kernel<T, true, true><<<dim3(gx, G), kBnWinThreads, 0, st>>>(a);
        } else {
            bn_bwd_win_kernel<T, false, false><<<dim3(gx, G), kBnWinThreads, 0, st>>>(a);
            if (check_launch("bn_bwd_reduce")) return 1;
            bn_bwd_win_kernel<T, false, true><<<dim3(gx, G), kBnWinThreads, 0, st>>>(a);
        }
    } else {
        constexpr int UNR = 4;
        const long long px = static_cast<long long>(a.group_images) * H * W;
        const int gx = static_cast<int>(std::max(1LL, std::min<long long>((px + lanes * UNR - 1) / (lanes * UNR), max_blocks)));
        if (prereduced) {
            // sums already reduced by the producing dgrad launch (PxParams::red_y): only the apply pass is left
            if (g2 != nullptr) return fail("bn_relu_bwd_apply: a second gradient source cannot be pre-reduced");
            bn_bwd_px_kernel<T, UNR, false, true><<<dim3(gx, G), 256, 0, st>>>(a);
        } else if (g2 != nullptr) {
            bn_bwd_px_kernel<T, UNR, true, false><<<dim3(gx, G), 256, 0, st>>>(a);
            if (check_launch("bn_bwd_reduce")) return 1;
            bn_bwd_px_kernel<T, UNR, true, true><<<dim3(gx, G), 256, 0, st>>>(a);
        } else {
            bn_bwd_px_kernel<T, UNR, false, false><<<dim3(gx, G), 256, 0, st>>>(a);
            if (check_launch("bn_bwd_reduce")) return 1;
            bn_bwd_px_kernel<T, UNR, false, true><<<dim3(gx, G), 256, 0, st>>>(a);
        }
    }
    if (check_launch("bn_bwd_apply")) return 1;
    if (dgamma0 != nullptr) {
        bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, G, C, dgamma0, dbeta0, dgamma1 ? dgamma1 : dgamma0,
                                                              dbeta1 ? dbeta1 : dbeta0);
        if (check_launch("bn_param_grad")) return 1;
    }
    return 0;
}

extern "C" {

int onet_bn_relu_bwd(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                     const float* mean, const float* invstd, int group_images, const void* g1, int64_t ld1, int off1,
                     const void* g2, int64_t ld2, int off2, const void* gp, const void* gp_arg, double* sums, double count,
                     void* dy, float* dgamma0, float* dbeta0, float* dgamma1, float* dbeta1, int dtype, void* stream) {
    if (static_cast<long long>(N) * H * W >= (1LL << 31)) return fail("bn_relu_bwd: more than 2^31 pixels");
    if (C % 8 || C > 2048) return fail("bn_relu_bwd: C must be a multiple of 8 and <= 2048");
    if (256 % (C / 8) != 0) return fail("bn_relu_bwd: C/8 must divide 256");
    if (gp != nullptr && C > 1024) return fail("bn_relu_bwd: the pooled variant supports C <= 1024");
    if (g1 == nullptr) return fail("bn_relu_bwd: g1 is required");
    if (dtype == ONET_F32)
        return bn_bwd_impl<float>(y, N, H, W, C, scale, shift, mean, invstd, group_images, g1, ld1, off1, g2, ld2, off2, gp, gp_arg,
                                  sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream));
    return bn_bwd_impl<bf16>(y, N, H, W, C, scale, shift, mean, invstd, group_images, g1, ld1, off1, g2, ld2, off2, gp, gp_arg,
                             sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream));
}

int onet_bn_relu_bwd_head(const void* y, int N, int H, int W, int C, const float* scale, const float* shift, const float* mean,
                          const float* invstd, int group_images, const void* L, int64_t ldl, int offl, const float* gv,
                          const float* gab, void* dL, double* sums, double count, void* dy, float* dgamma0, float* dbeta0,
                          float* dgamma1, float* dbeta1, int dtype, void* stream) {
    if (static_cast<long long>(N) * H * W >= (1LL << 31)) return fail("bn_relu_bwd_head: more than 2^31 pixels");
    if (C % 8 || C > 2048 || 256 % (C / 8) != 0) return fail("bn_relu_bwd_head: C must be a multiple of 8 with C/8 dividing 256");
    if (L == nullptr || gv == nullptr || gab == nullptr || dL == nullptr) return fail("bn_relu_bwd_head: L, gv, gab and dL are required");
    if (dtype == ONET_F32)
        return bn_bwd_impl<float>(y, N, H, W, C, scale, shift, mean, invstd, group_images, L, ldl, offl, nullptr, 0, 0, nullptr, nullptr,
                                  sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream), false, gv, gab, dL);
    return bn_bwd_impl<bf16>(y, N, H, W, C, scale, shift, mean, invstd, group_images, L, ldl, offl, nullptr, 0, 0, nullptr, nullptr,
                             sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream), false, gv, gab, dL);
}

int onet_bn_relu_bwd_apply(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                           const float* mean, const float* invstd, int group_images, const void* g1, int64_t ld1, int off1,
                           double* sums, double count, void* dy, float* dgamma0, float* dbeta0, float* dgamma1,
                           float* dbeta1, int dtype, void* stream) {
    if (static_cast<long long>(N) * H * W >= (1LL << 31)) return fail("bn_relu_bwd_apply: more than 2^31 pixels");
    if (C % 8 || C > 2048 || 256 % (C / 8) != 0) return fail("bn_relu_bwd_apply: C must be a multiple of 8 with C/8 dividing 256");
    if (g1 == nullptr || sums == nullptr) return fail("bn_relu_bwd_apply: g1 and sums are required");
    if (dtype == ONET_F32)
        return bn_bwd_impl<float>(y, N, H, W, C, scale, shift, mean, invstd, group_images, g1, ld1, off1, nullptr, 0, 0, nullptr, nullptr,
                                  sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream), true);
    return bn_bwd_impl<bf16>(y, N, H, W, C, scale, shift, mean, invstd, group_images, g1, ld1, off1, nullptr, 0, 0, nullptr, nullptr,
                             sums, count, dy, dgamma0, dbeta0, dgamma1, dbeta1, ST(stream), true);
}

int onet_conv3x3_dgrad_bnred(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                             void* out, const void* y_prev, const float* scale, const float* shift, const float* mean,
                             const float* invstd, double* sums, int group_images, int dtype, int engine, void* stream) {
    if (dtype != ONET_BF16 || engine != ONET_ENGINE_TC) return fail("conv3x3_dgrad_bnred: tcgen05 bf16 path only");
    if (y_prev == nullptr || scale == nullptr || shift == nullptr || mean == nullptr || invstd == nullptr || sums == nullptr)
        return fail("conv3x3_dgrad_bnred: y_prev, scale, shift, mean, invstd and sums are required");
    BnRedArgs red{static_cast<const bf16*>(y_prev), scale, shift, mean, invstd, sums};
    return conv3x3_tc<OpBf16>(static_cast<const bf16*>(in), ldi, ci_off, N, H, W, Cin, static_cast<const bf16*>(wp), Cout,
                              static_cast<bf16*>(out), Cout, 0, nullptr, nullptr, group_images, ST(stream), nullptr, nullptr, &red);
}

int onet_convT2x2_fwd(const void* x, int64_t ldx, int xoff, int N, int H, int W, int Cin, const void* w,
                      const float* bias, int Co, void* out, int64_t ldo, int ooff, int Ho, int Wo, int dtype, int engine,
                      void* stream) {
    if (Ho == 0) Ho = 2 * H;
    if (Wo == 0) Wo = 2 * W;
    if (Ho < 2 * H || Wo < 2 * W) return fail("convT2x2_fwd: output grid %dx%d smaller than 2H x 2W", Ho, Wo);
    if (engine == ONET_ENGINE_TC) {
        if (dtype == ONET_F32)
            return convT_fwd_tc<OpTf32>(static_cast<const float*>(x), ldx, xoff, N, H, W, Cin, static_cast<const float*>(w), bias, Co,
                                        static_cast<float*>(out), ldo, ooff, Ho, Wo, ST(stream));
        return convT_fwd_tc<OpBf16>(static_cast<const bf16*>(x), ldx, xoff, N, H, W, Cin, static_cast<const bf16*>(w), bias, Co,
                                    static_cast<bf16*>(out), ldo, ooff, Ho, Wo, ST(stream));
    }
    const long long total = 4LL * N * H * W * Co;
    if (dtype == ONET_F32)
        convT2x2_fwd_simt_kernel<float><<<grid_for(total, 256, 148 * 64), 256, 0, ST(stream)>>>(
            static_cast<const float*>(x), ldx, xoff, N, H, W, Cin, static_cast<const float*>(w), bias, Co, static_cast<float*>(out), ldo, ooff, 0, Ho, Wo);
    else
        convT2x2_fwd_simt_kernel<bf16><<<grid_for(total, 256, 148 * 64), 256, 0, ST(stream)>>>(
            static_cast<const bf16*>(x), ldx, xoff, N, H, W, Cin, static_cast<const float*>(w), bias, Co, static_cast<bf16*>(out), ldo, ooff, 1, Ho, Wo);
    return check_launch("convT2x2_fwd_simt");
}

int onet_convT2x2_dgrad(const void* go, int64_t ldg, int goff, int N, int H, int W, int Cin, const void* w, int Co,
                        void* dx, int64_t ldd, int doff, int Ho, int Wo, int dtype, int engine, void* stream) {
    if (Ho == 0) Ho = 2 * H;
    if (Wo == 0) Wo = 2 * W;
    if (Ho < 2 * H || Wo < 2 * W) return fail("convT2x2_dgrad: gradient grid %dx%d smaller than 2H x 2W", Ho, Wo);
    if (engine == ONET_ENGINE_TC) {
        if (dtype == ONET_F32)
            return convT_dgrad_tc<OpTf32>(static_cast<const float*>(go), ldg, goff, N, H, W, Cin, static_cast<const float*>(w), Co,
                                          static_cast<float*>(dx), ldd, doff, Ho, Wo, ST(stream));
        return convT_dgrad_tc<OpBf16>(static_cast<const bf16*>(go), ldg, goff, N, H, W, Cin, static_cast<const bf16*>(w), Co,
                                      static_cast<bf16*>(dx), ldd, doff, Ho, Wo, ST(stream));
    }
    const long long total = static_cast<long long>(N) * H * W * Cin;
    if (dtype == ONET_F32)
        convT2x2_dgrad_simt_kernel<float><<<grid_for(total, 256, 148 * 64), 256, 0, ST(stream)>>>(
            static_cast<const float*>(go), ldg, goff, N, H, W, Cin, static_cast<const float*>(w), Co, static_cast<float*>(dx), ldd, doff, Ho, Wo);
    else
        convT2x2_dgrad_simt_kernel<bf16><<<grid_for(total, 256, 148 * 64), 256, 0, ST(stream)>>>(
            static_cast<const bf16*>(go), ldg, goff, N, H, W, Cin, static_cast<const float*>(w), Co, static_cast<bf16*>(dx), ldd, doff, Ho, Wo);
    return check_launch("convT2x2_dgrad_simt");
}

int onet_zero_border(void* buf, int N, int Ho, int Wo, int64_t ld, int coff, int C, int Hv, int Wv, int dtype, void* stream) {
    if (Hv > Ho || Wv > Wo) return fail("zero_border: valid window larger than the buffer");
    const long long total = static_cast<long long>(N) * ((Ho - Hv) * Wo + Hv * (Wo - Wv)) * C;
    if (total == 0) return 0;
    if (dtype == ONET_F32)
        zero_border_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<float*>(buf), N, Ho, Wo, ld, coff, C, Hv, Wv);
    else
        zero_border_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<bf16*>(buf), N, Ho, Wo, ld, coff, C, Hv, Wv);
    return check_launch("zero_border");
}

int onet_add_colsums(const double* sums, int C, float* dst, void* stream) {
    add_colsums_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(sums, C, dst);
    return check_launch("add_colsums");
}

int onet_convT2x2_wgrad(const void* x, int64_t ldx, int xoff, const void* go, int64_t ldg, int goff, int N, int H,
                        int W, int Cin, int Co, float* dw, float* dbias, int Ho, int Wo, int dtype, int engine, void* stream) {
    if (Ho == 0) Ho = 2 * H;
    if (Wo == 0) Wo = 2 * W;
    if (Ho < 2 * H || Wo < 2 * W) return fail("convT2x2_wgrad: gradient grid %dx%d smaller than 2H x 2W", Ho, Wo);
    const long long M = static_cast<long long>(N) * H * W;
    if (dbias != nullptr) {
        // bias gradient = column sums of dO over the whole upsampled grid
        const int cpb = std::min(Co, 64);
        dim3 grid(static_cast<unsigned>(std::min<long long>(148 * 4, (4 * M + (256 / cpb) - 1) / (256 / cpb))), (Co + cpb - 1) / cpb);
        float* part = (dtype == ONET_F32 && grid.x > 1) ? splitk_ws(static_cast<long long>(grid.x) * Co) : nullptr;
        if (dtype == ONET_F32)
            colsum_kernel<float><<<grid, 256, 0, ST(stream)>>>(static_cast<const float*>(go), ldg, goff, 4 * M, Co, dbias, 2 * H, 2 * W, Ho, Wo, part);
        else
            colsum_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16*>(go), ldg, goff, 4 * M, Co, dbias, 2 * H, 2 * W, Ho, Wo, part);
        if (check_launch("colsum")) return 1;
        if (part != nullptr) {
            splitk_reduce(part, static_cast<int>(grid.x), Co, dbias, ST(stream));
            if (check_launch("splitk_reduce")) return 1;
        }
    }
    if (engine == ONET_ENGINE_TC) {
        // M-side = dO on the upsampled grid (m = co), N-side = X (n = ci); dW[ci][co][tap] -> transposed output
        if (dtype == ONET_F32)
            return wgrad_tc<OpTf32>(static_cast<const float*>(go), ldg, goff, Co, true, static_cast<const float*>(x), ldx, xoff, Cin,
                                    N, H, W, 4, dw, true, ST(stream), Ho, Wo);
        return wgrad_tc<OpBf16>(static_cast<const bf16*>(go), ldg, goff, Co, true, static_cast<const bf16*>(x), ldx, xoff, Cin, N, H, W,
                                4, dw, true, ST(stream), Ho, Wo);
    }
    const long long nthreads = 4LL * Cin * Co;
    int splits = static_cast<int>(std::max<long long>(1, std::min<long long>(M, (148LL * 2048) / std::max<long long>(1, nthreads))));
    const long long per = (M + splits - 1) / splits;
    splits = static_cast<int>((M + per - 1) / per);
    dim3 grid(static_cast<unsigned>((nthreads + 255) / 256), splits);
    float* part = (dtype == ONET_F32 && splits > 1) ? splitk_ws(static_cast<long long>(splits) * nthreads) : nullptr;
    if (dtype == ONET_F32)
        convT2x2_wgrad_simt_kernel<float><<<grid, 256, 0, ST(stream)>>>(static_cast<const float*>(x), ldx, xoff,
                                                                        static_cast<const float*>(go), ldg, goff, N, H, W, Cin, Co, dw, per, Ho, Wo, part);
    else
        convT2x2_wgrad_simt_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), ldx, xoff,
                                                                       static_cast<const bf16*>(go), ldg, goff, N, H, W, Cin, Co, dw, per, Ho, Wo, part);
    if (check_launch("convT2x2_wgrad_simt")) return 1;
    if (part != nullptr) {
        splitk_reduce(part, splits, nthreads, dw, ST(stream));
        return check_launch("splitk_reduce");
    }
    return 0;
}

}  // extern "C"

// bf16 head forward on warp-level MMAs (head_mma.cuh); ONET_NO_HEAD_MMA=1 = the CUDA-core kernel (A/B)
static bool head_mma_ok(long long npx) {
    static const bool off = getenv("ONET_NO_HEAD_MMA") != nullptr;
    return !off && npx < (1LL << 30);
}
static int head_mma_grid(long long npx) {
    const long long tiles = (npx + 15) / 16;
    return static_cast<int>(std::max(1LL, std::min<long long>((tiles + 7) / 8, 148 * 2 * 4)));
}

template <typename T>
static void fill_head(HeadArgs<T>& a, const void* L, int64_t ldl, int offl, const void* Hf, int64_t ldh, int offh, int B,
                      int H, int W) {
    memset(&a, 0, sizeof(a));
    a.L = static_cast<const T*>(L); a.ldl = ldl; a.offl = offl;
    a.Hf = static_cast<const T*>(Hf); a.ldh = ldh; a.offh = offh;
    a.B = B; a.HW = static_cast<long long>(H) * W;
}

extern "C" {

int onet_head_fwd(const void* L, int64_t ldl, int offl, const void* Hf, int64_t ldh, int offh, int B, int H, int W,
                  float* Vt, float* Vd, float* S, float* a_out, float* b_out, double* loss_acc, int dtype, void* stream) {
    if (ldl % 8 || offl % 8 || ldh % 8 || offh % 8) return fail("head: channel strides/offsets must be multiples of 8");
    const long long npx = static_cast<long long>(B) * H * W;
    const int grid = grid_for(npx * 8, 256, 148 * 8);
    if (dtype == ONET_F32) {
        HeadArgs<float> a;
        fill_head(a, L, ldl, offl, Hf, ldh, offh, B, H, W);
        a.Vt = Vt; a.Vd = Vd; a.S = S; a.a = a_out; a.b = b_out; a.loss_acc = loss_acc;
        head_fwd_kernel<float><<<grid, 256, 0, ST(stream)>>>(a);
    } else {
        HeadArgs<bf16> a;
        fill_head(a, L, ldl, offl, Hf, ldh, offh, B, H, W);
        a.Vt = Vt; a.Vd = Vd; a.S = S; a.a = a_out; a.b = b_out; a.loss_acc = loss_acc;
        if (head_mma_ok(npx)) head_fwd_mma_kernel<<<head_mma_grid(npx), 256, 0, ST(stream)>>>(a, make_fastdiv(H * W));
        else head_fwd_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(a);
    }
    return check_launch("head_fwd");
}

int onet_head_fwd_bn(const void* L, int64_t ldl, int offl, const void* Y, int64_t ldy, int offy, int B, int H, int W,
                     const float* hscale_t, const float* hshift_t, const float* hscale_d, const float* hshift_d,
                     float* Vt, float* Vd, float* S, float* a_out, float* b_out, double* loss_acc, int dtype, void* stream) {
    if (ldl % 8 || offl % 8 || ldy % 8 || offy % 8) return fail("head: channel strides/offsets must be multiples of 8");
    if (!hscale_t || !hshift_t || !hscale_d || !hshift_d) return fail("head_fwd_bn: the four scale / shift vectors are required");
    const long long npx = static_cast<long long>(B) * H * W;
    const int grid = grid_for(npx * 8, 256, 148 * 8);
    if (dtype == ONET_F32) {
        HeadArgs<float> a;
        fill_head(a, L, ldl, offl, Y, ldy, offy, B, H, W);
        a.Vt = Vt; a.Vd = Vd; a.S = S; a.a = a_out; a.b = b_out; a.loss_acc = loss_acc;
        a.hsc_t = hscale_t; a.hsh_t = hshift_t; a.hsc_d = hscale_d; a.hsh_d = hshift_d;
        head_fwd_kernel<float><<<grid, 256, 0, ST(stream)>>>(a);
    } else {
        HeadArgs<bf16> a;
        fill_head(a, L, ldl, offl, Y, ldy, offy, B, H, W);
        a.Vt = Vt; a.Vd = Vd; a.S = S; a.a = a_out; a.b = b_out; a.loss_acc = loss_acc;
        a.hsc_t = hscale_t; a.hsh_t = hshift_t; a.hsc_d = hscale_d; a.hsh_d = hshift_d;
        if (head_mma_ok(npx)) head_fwd_mma_kernel<<<head_mma_grid(npx), 256, 0, ST(stream)>>>(a, make_fastdiv(H * W));
        else head_fwd_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(a);
    }
    return check_launch("head_fwd+bn");
}

int onet_head_bwd_scalars(const float* Vt, const float* Vd, const float* a_in, const float* b_in, const float* gscale,
                          const float* gVt, const float* gVd, const float* gS, int B, int H, int W, float* gv, float* gab,
                          void* stream) {
    const long long npx = static_cast<long long>(B) * H * W;
    if (npx <= 0) return 0;
    if (gv == nullptr || gab == nullptr) return fail("head_bwd_scalars: gv and gab are required");
    head_bwd_scalars_kernel<<<grid_for(npx, 256, 148 * 8), 256, 0, ST(stream)>>>(Vt, Vd, a_in, b_in, gscale, gVt, gVd, gS, npx,
                                                                               static_cast<long long>(H) * W, gv, gab);
    return check_launch("head_bwd_scalars");
}

int onet_head_bwd(const void* L, int64_t ldl, int offl, const void* Hf, int64_t ldh, int offh, int B, int H, int W,
                  const float* Vt, const float* Vd, const float* a_in, const float* b_in, const float* gscale,
                  const float* gVt, const float* gVd, const float* gS, void* dL, void* dH, int dtype, void* stream) {
    const long long npx = static_cast<long long>(B) * H * W;
    const int grid = grid_for(npx * 8, 256, 148 * 8);
    if (dtype == ONET_F32) {
        HeadArgs<float> a;
        fill_head(a, L, ldl, offl, Hf, ldh, offh, B, H, W);
        a.Vt = const_cast<float*>(Vt); a.Vd = const_cast<float*>(Vd); a.a = const_cast<float*>(a_in); a.b = const_cast<float*>(b_in);
        a.gscale = gscale; a.gVt = gVt; a.gVd = gVd; a.gS = gS;
        a.dL = static_cast<float*>(dL); a.dH = static_cast<float*>(dH);
        head_bwd_kernel<float><<<grid, 256, 0, ST(stream)>>>(a);
    } else {
        HeadArgs<bf16> a;
        fill_head(a, L, ldl, offl, Hf, ldh, offh, B, H, W);
        a.Vt = const_cast<float*>(Vt); a.Vd = const_cast<float*>(Vd); a.a = const_cast<float*>(a_in); a.b = const_cast<float*>(b_in);
        a.gscale = gscale; a.gVt = gVt; a.gVd = gVd; a.gS = gS;
        a.dL = static_cast<bf16*>(dL); a.dH = static_cast<bf16*>(dH);
        head_bwd_kernel<bf16><<<grid, 256, 0, ST(stream)>>>(a);
    }
    return check_launch("head_bwd");
}

int onet_predict_label(const float* Vt, const float* Vd, int64_t n, int64_t* out, void* stream) {
    predict_label_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(Vt, Vd, n, reinterpret_cast<long long*>(out));
    return check_launch("predict_label");
}

int onet_predict_label_u8(const float* Vt, const float* Vd, int64_t n, unsigned char* out, void* stream) {
    if (n <= 0) return 0;
    predict_label_u8_kernel<<<grid_for((n + 15) / 16, 256, 148 * 8), 256, 0, ST(stream)>>>(Vt, Vd, n, out);
    return check_launch("predict_label_u8");
}

int onet_eval_confusion(const float* Vt, const float* Vd, const int64_t* gt, int64_t n, int64_t* counts, void* stream) {
    if (n <= 0) return 0;
    eval_confusion_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, ST(stream)>>>(Vt, Vd, reinterpret_cast<const long long*>(gt), n,
                                                                           reinterpret_cast<unsigned long long*>(counts));
    return check_launch("eval_confusion");
}

int onet_normalize_per_frame(const float* x, int frames, int64_t hw, int* work, float* out, void* stream) {
    if (frames <= 0 || hw <= 0) return 0;
    if (work == nullptr) return fail("normalize_per_frame: work (2 * frames ints) is required");
    if (frames > 65535) return fail("normalize_per_frame: at most 65535 frames per call");
    frame_minmax_init_kernel<<<(frames + 255) / 256, 256, 0, ST(stream)>>>(work, frames);
    if (check_launch("frame_minmax_init")) return 1;
    // enough blocks per frame to fill the GPU, not more than the frame has 1024-element chunks
    const long long chunks = std::max<long long>(1, (hw + 1023) / 1024);
    const int per_frame = static_cast<int>(std::min<long long>(chunks, std::max(1, 148 * 8 / frames)));
    frame_minmax_kernel<<<dim3(per_frame, frames), 256, 0, ST(stream)>>>(x, hw, work);
    if (check_launch("frame_minmax")) return 1;
    const float eps = static_cast<float>(2.220446049250313e-16);        // np.spacing(1)
    frame_normalize_kernel<<<dim3(per_frame, frames), 256, 0, ST(stream)>>>(x, hw, work, eps, out);
    return check_launch("frame_normalize");
}

int onet_synth_rayleigh(float* out, int64_t n, float sigma, int64_t seed, int stream_id, void* stream) {
    if (n <= 0) return 0;
    rayleigh_fill_kernel<<<grid_for((n + 3) / 4, 256, 148 * 8), 256, 0, ST(stream)>>>(
        out, n, sigma, static_cast<uint32_t>(seed), static_cast<uint32_t>(static_cast<uint64_t>(seed) >> 32), static_cast<uint32_t>(stream_id));
    return check_launch("synth_rayleigh");
}

int onet_synth_kclutter(float* out, int64_t n, int nu, int64_t seed, int stream_id, void* stream) {
    if (n <= 0) return 0;
    if (nu < 1 || nu > 64) return fail("synth_kclutter: integer texture shape nu in [1, 64]");
    kclutter_fill_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, ST(stream)>>>(
        out, n, nu, static_cast<uint32_t>(seed), static_cast<uint32_t>(static_cast<uint64_t>(seed) >> 32), static_cast<uint32_t>(stream_id));
    return check_launch("synth_kclutter");
}

int onet_synth_add_targets(float* frames, unsigned char* masks, int n_frames, int H, int W, const void* targets,
                           int targets_per_frame, float snr_db, float* erc_out, void* stream) {
    if (n_frames <= 0) return 0;
    if (frames == nullptr || masks == nullptr) return fail("synth_add_targets: frames and masks are required");
    if (targets_per_frame > 0 && targets == nullptr) return fail("synth_add_targets: targets table is null");
    static_assert(sizeof(SynthTarget) == 32, "SynthTarget is 8 x 4 bytes (onet_b200.h)");
    add_targets_kernel<<<n_frames, 256, 0, ST(stream)>>>(frames, masks, H, W, static_cast<const SynthTarget*>(targets),
                                                         targets_per_frame, powf(10.0f, snr_db / 20.0f), erc_out);
    return check_launch("synth_add_targets");
}

int onet_synth_normal(double* out, int64_t n, int64_t seed, int stream_id, void* stream) {
    if (n <= 0) return 0;
    normal_fill_kernel<<<grid_for((n + 1) / 2, 256, 148 * 8), 256, 0, ST(stream)>>>(
        out, n, static_cast<uint32_t>(seed), static_cast<uint32_t>(static_cast<uint64_t>(seed) >> 32), static_cast<uint32_t>(stream_id));
    return check_launch("synth_normal");
}

int onet_kfield_mnlt(const double* x, int64_t n, int v, double* y, void* stream) {
    if (n <= 0) return 0;
    if (v < 1 || v > 64) return fail("kfield_mnlt: integer gamma shape v in [1, 64]");
    kfield_mnlt_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, ST(stream)>>>(x, n, v, y);
    return check_launch("kfield_mnlt");
}

int onet_kfield_coeff_sums(const double* x, const double* g, int frames, int64_t per_frame, double* sums, void* stream) {
    if (frames <= 0 || per_frame <= 0) return 0;
    if (frames > 65535) return fail("kfield_coeff_sums: at most 65535 frames per call");
    const int per = static_cast<int>(std::min<long long>((per_frame + 1023) / 1024, std::max(1, 148 * 4 / frames)));
    kfield_coeff_sums_kernel<<<dim3(per, frames), 256, 0, ST(stream)>>>(x, g, per_frame, sums);
    return check_launch("kfield_coeff_sums");
}

int onet_kfield_acf_root(const double* coeffs, const double* acf, int frames, int64_t per_frame, double* out, void* stream) {
    if (frames <= 0 || per_frame <= 0) return 0;
    if (frames > 65535) return fail("kfield_acf_root: at most 65535 frames per call");
    const int per = static_cast<int>(std::min<long long>((per_frame + 255) / 256, std::max(1, 148 * 8 / frames)));
    kfield_acf_root_kernel<<<dim3(per, frames), 256, 0, ST(stream)>>>(coeffs, acf, per_frame, reinterpret_cast<double2*>(out));
    return check_launch("kfield_acf_root");
}

int onet_kfield_amplitude(const double* speckle, const double* texture, int64_t n, float* out, void* stream) {
    if (n <= 0) return 0;
    kfield_amplitude_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, ST(stream)>>>(reinterpret_cast<const double2*>(speckle), texture, n, out);
    return check_launch("kfield_amplitude");
}

int onet_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale, void* stream) {
    if (step < 1) return fail("adam: step must be >= 1");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
    adam_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
    return check_launch("adam");
}

int onet_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int* step,
                       float grad_scale, void* stream) {
    if (hyper == nullptr || step == nullptr) return fail("adam_step_dev: hyper and step must be device pointers");
    adam_tick_kernel<<<1, 1, 0, ST(stream)>>>(step);
    if (check_launch("adam_tick")) return 1;
    adam_dev_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, ST(stream)>>>(p, g, m, v, n, hyper, step, grad_scale);
    return check_launch("adam_dev");
}

}  // extern "C"
