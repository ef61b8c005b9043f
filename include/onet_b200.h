/* onet_b200 — C ABI of the B200-native Onet hot path (libonet_b200.so).
 *
 * The reference (joeyee/Onet) is pure Python/PyTorch and has no FFI of its own: its hot path is the ATen
 * operator sequence dispatched by `source_code/Onet_vanilla_20240606.py`.  Each entry point below replaces
 * one group of those ATen calls; the file:line it replaces is cited next to it.  The host-side mirror of the
 * reference interface (class `Onet` with forward / compute_loss / predict_label and identical state_dict
 * keys) lives in `onet_b200/model.py` and binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); no allocation happens inside.
 *   - activations are NHWC; `ld*` is the number of channels per pixel of the underlying buffer and `*off`
 *     the first channel used, so a view into a wider (skip-concat) buffer needs no copy.
 *   - the twin batch holds the top branch in images [0,B) and the down branch in [B,2B); `group_images` = B
 *     is the BatchNorm statistics group size (statistics are per branch, Onet_vanilla_20240606.py:175,181).
 *   - dtype: ONET_F32 (FP32 verification mode, CUDA-core FMA) or ONET_BF16 (bf16 storage, fp32 accumulate).
 *   - engine: ONET_ENGINE_SIMT (CUDA cores) or ONET_ENGINE_TC (tcgen05 + TMEM + TMA; bf16 only,
 *     channel counts multiples of 64).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value 0 = ok; non-zero = error, message via onet_last_error().  No exceptions, no global state
 *     except the last-error string (thread-local) and the cached driver entry point for tensor-map encoding.
 */
#ifndef ONET_B200_H
#define ONET_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONET_F32 0
#define ONET_BF16 1
#define ONET_ENGINE_SIMT 0
#define ONET_ENGINE_TC 1

int onet_version(void);
/* number of kernels this process has launched through the library so far */
int64_t onet_launch_count(void);
/* name of the kernel variant the most recent call on this thread launched, e.g. "conv3x3_halo2_px_kernel<256>"
 * (measurement only: bench.py groups its per-launch CUDA-event times by it) */
const char* onet_last_kernel(void);
const char* onet_last_error(void);
int onet_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* FP32 verification mode, reproducible gradients: registers (per device; NULL unregisters) a caller-owned fp32 workspace.
 * While one is registered the CUDA-core weight-gradient kernels (onet_conv3x3_wgrad, onet_convT2x2_wgrad incl. its bias
 * column sums; dtype ONET_F32, engine ONET_ENGINE_SIMT) write per-split partial sums into it and add them in split order,
 * instead of accumulating with fp32 atomics; a call whose partials do not fit falls back to atomics.  The workspace must
 * stay alive, and calls that use it must be ordered on one stream, until it is unregistered.  The reference's CPU
 * convolution backward (ATen, torch>=1.7.0) is deterministic; this is what lets the fp32 parity tests run once. */
int onet_set_splitk_workspace(float* ws, int64_t nfloats);

/* X (B,Cin,H,W) fp32 NCHW -> twin NHWC batch [2B,H,W,Cin] = (X, clip(1-X+bias,0,1)).
 * Replaces Onet.forward's input handling, Onet_vanilla_20240606.py:175,180-181. */
int onet_prep_input(const float* x, int B, int Cin, int H, int W, float bias, void* out, int dtype, void* stream);

/* Pack fp32 master weights into kernel operand layouts (once per optimizer step).
 * conv  w[Co][Ci][3][3] -> wf[Co][tap][Ci], wd[Ci][8-tap][Co] (wd may be NULL);
 * convT w[Ci][Co][2][2] -> wf[(tap,co)][ci], wd[ci][(tap,co)]. */
int onet_pack_conv_weights(const float* w, int Cout, int Cin, void* wf, void* wd, int dtype, void* stream);
int onet_pack_convT_weights(const float* w, int Cin, int Cout, void* wf, void* wd, int dtype, void* stream);

/* The same packing for up to 24 layers in one launch (the per-step refresh after the optimizer update).  Host arrays
 * of length n: w[i] fp32 master weight, d0/d1 its outer/inner channel counts (conv: Cout, Cin; convT: Cin, Cout),
 * taps[i] = 9 (conv) or 4 (convT), wf[i]/wd[i] the packed destinations (wd[i] may be NULL for a conv). */
int onet_pack_all_weights(int n, const float* const* w, const int* d0, const int* d1, const int* taps, void* const* wf,
                          void* const* wd, int dtype, void* stream);

/* 3x3 / pad 1 / no-bias convolution on packed weights wp[Cout][9][Cin]; writes the RAW output and, when
 * stat_sum != NULL, accumulates per-group per-channel sum / sum-of-squares (doubles, [groups][Cout]) for the
 * training-mode BatchNorm that follows (with the tensor-core engine stat_sq may be NULL: column sums only).  Used for forward (nn.Conv2d, Onet_vanilla_20240606.py:47,51) and,
 * with wd, for the data gradient (autograd of the same lines). */
int onet_conv3x3_fwd(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                     void* out, int64_t ldo, int co_off, double* stat_sum, double* stat_sq, int group_images,
                     int dtype, int engine, void* stream);

/* The FIRST convolution of each U-Net (in_chns 1 or 3 -> 64 at full resolution, Onet_vanilla_20240606.py:111,47-49) without
 * materialising its raw output: 64 outputs per pixel are recomputed from the 3 x 3 x in_chns patch wherever they are needed
 * (x [N,H,W,Cin] dense, wp = packed wf [64][9][Cin], W % 4 == 0).  Results are those of onet_conv3x3_fwd + onet_bn_relu_apply
 * (bit for bit) and of onet_bn_relu_bwd + onet_conv3x3_wgrad (same expressions, sums in a different order):
 *   onet_first_conv_stats   : BatchNorm partial sums of conv(x) as stored (rounded to dtype) -> onet_bn_finalize
 *   onet_first_conv_bn_relu : out [N,H,W,64] = relu(conv(x) * scale + shift)   (training: batch statistics; eval: running)
 *   onet_first_conv_bwd     : g [N,H,W,64] = gradient w.r.t. out; reduces the BatchNorm-backward sums (sums: zeroed double
 *                             [G][2][64]), forms dY in registers and accumulates dw [64][Cin][3][3] (+ dgamma / dbeta).
 * With `gram` (zeroed double [G][K + K*K], K = 9 * in_chns: patch moments S[K], G[K][K] per statistics group, patch element
 * k = tap * in_chns + channel): the conv output is linear in the 3 x 3 x in_chns patch, so the statistics are w.S and w^T G w (no
 * pass over 64 channels), and the backward is ONE pass over g (s1 and A[c][k] = sum dz v[k] into `acc_a`, zeroed float [G][64][K])
 * followed by a closed-form assembly of dW (and s2) from A, S, G; y is then used unrounded everywhere (round_y = 0 in
 * onet_first_conv_bn_relu).  Available for in_chns = 1 (any dtype) and for in_chns = 3 with bf16 storage (moments and backward on
 * warp-level MMAs).  gram = NULL: the two-pass form whose values are those of the stored path (round_y = 1). */
int onet_first_conv_stats(const void* x, int N, int H, int W, int Cin, const void* wp, double* gram, double* stat_sum,
                          double* stat_sq, int group_images, int dtype, void* stream);
int onet_first_conv_bn_relu(const void* x, int N, int H, int W, int Cin, const void* wp, const float* scale, const float* shift,
                            int group_images, void* out, int round_y, int dtype, void* stream);
int onet_first_conv_bwd(const void* x, int N, int H, int W, int Cin, const void* wp, const float* scale, const float* shift,
                        const float* mean, const float* invstd, int group_images, const void* g, const double* gram, float* acc_a,
                        double* sums, double count, float* dw, float* dgamma0, float* dbeta0, float* dgamma1, float* dbeta1,
                        int dtype, void* stream);

/* Inference: the same convolution with BatchNorm(eval) + ReLU folded into the epilogue, out = relu(conv * scale + shift)
 * (scale / shift [G][Cout] from onet_bn_eval_prepare), written straight to its destination (e.g. a concat-buffer slice):
 * nn.Conv2d + nn.BatchNorm2d(eval) + nn.ReLU, Onet_vanilla_20240606.py:47-53.  Tensor-core engine (bf16, or fp32 storage with tf32 operands). */
int onet_conv3x3_bn_relu_infer(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                               const float* scale, const float* shift, int group_images, void* out, int64_t ldo, int co_off,
                               int dtype, int engine, void* stream);
/* nn.MaxPool2d(2) (floor) of channels [ioff, ioff+C) of an NHWC buffer -> dense [N,H/2,W/2,C] (inference path, :67) */
int onet_maxpool2x2(const void* in, int64_t ldi, int ioff, int N, int H, int W, int C, void* out, int dtype, void* stream);

/* Weight gradient of the same convolution, accumulated (atomics) into dw fp32 [Cout][Cin][3][3]. */
int onet_conv3x3_wgrad(const void* g, int64_t ldg, int g_off, const void* in, int64_t ldi, int ci_off, int N, int H,
                       int W, int Cin, int Cout, float* dw, int dtype, int engine, void* stream);

/* BatchNorm2d training statistics -> mean / invstd / scale / shift ([G][C] floats) + running-buffer update
 * (momentum, unbiased variance), groups folded sequentially (nn.BatchNorm2d, Onet_vanilla_20240606.py:48,52).
 * Per-group parameter pointers; for the weight-shared twin both groups pass the same pointers. */
int onet_bn_finalize(const double* stat_sum, const double* stat_sq, int G, int C, double count,
                     const float* gamma0, const float* beta0, float* running_mean0, float* running_var0,
                     const float* gamma1, const float* beta1, float* running_mean1, float* running_var1,
                     float momentum, float* mean, float* invstd, float* scale, float* shift, void* stream);
/* eval mode: scale / shift from the running statistics */
int onet_bn_eval_prepare(int G, int C, const float* gamma0, const float* beta0, const float* running_mean0,
                         const float* running_var0, const float* gamma1, const float* beta1,
                         const float* running_mean1, const float* running_var1, float* scale, float* shift,
                         void* stream);

/* y -> relu(y*scale+shift) written to out (+ooff, ld ldo; e.g. the skip half of a concat buffer) and, when
 * pool != NULL, the 2x2 max-pooled map [N,H/2,W/2,C] in the same pass (nn.ReLU :49,53 + nn.MaxPool2d(2) :67
 * + the skip half of torch.cat :100).  Nothing else is saved: the backward pass recomputes the window arg-max. */
int onet_bn_relu_apply(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                       int group_images, void* out, int64_t ldo, int ooff, void* pool, void* pool_arg, int dtype, void* stream);

/* Backward of BN -> ReLU (-> skip / max-pool): gradient sources g1 (+ optional g2, optional pooled gp routed to
 * the first maximum of each 2x2 window: read from gp_arg, the uint16 [N,H/2,W/2,C/8] arg-max words (2 bits per channel)
 * onet_bn_relu_apply wrote into its pool_arg, or recomputed from y when gp_arg is NULL), produces dy [N,H,W,C] and
 * accumulates dgamma/dbeta.
 * `sums` is a zero-initialised [G][2][C] double workspace. */
int onet_bn_relu_bwd(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                     const float* mean, const float* invstd, int group_images, const void* g1, int64_t ld1, int off1,
                     const void* g2, int64_t ld2, int off2, const void* gp, const void* gp_arg, double* sums, double count,
                     void* dy, float* dgamma0, float* dbeta0, float* dgamma1, float* dbeta1, int dtype, void* stream);

/* Backward fusion (tcgen05 bf16 path): the 3x3 data gradient whose OUTPUT is the gradient g w.r.t. the post-ReLU activation
 * of the previous conv layer (the two convs of a DoubleConv, Onet_vanilla_20240606.py:46-53).  Same product as
 * onet_conv3x3_fwd with the flipped weights `wp`; in addition the epilogue reduces that previous layer's BatchNorm-backward
 * sums from g (as stored, bf16) and its raw conv output y_prev [N,H,W,Cout]:
 *   sums[grp][0][c] += sum dz, sums[grp][1][c] += sum dz * (y - mean[grp][c]) * invstd[grp][c], dz = (relu(bn(y)) > 0) ? g : 0
 * (sums: zero-initialised double [G][2][Cout]), so the previous layer's BatchNorm backward only needs its apply pass:
 * onet_bn_relu_bwd_apply, same arguments as onet_bn_relu_bwd without the second / pooled gradient sources.  Layer shapes
 * whose kernel variant has no fused epilogue (images smaller than 16 x 8, few tiles) run the plain launch followed by the
 * standalone reduce pass: same results. */
int onet_conv3x3_dgrad_bnred(const void* in, int64_t ldi, int ci_off, int N, int H, int W, int Cin, const void* wp, int Cout,
                             void* out, const void* y_prev, const float* scale, const float* shift, const float* mean,
                             const float* invstd, double* sums, int group_images, int dtype, int engine, void* stream);
int onet_bn_relu_bwd_apply(const void* y, int N, int H, int W, int C, const float* scale, const float* shift,
                           const float* mean, const float* invstd, int group_images, const void* g1, int64_t ld1, int off1,
                           double* sums, double count, void* dy, float* dgamma0, float* dbeta0, float* dgamma1,
                           float* dbeta1, int dtype, void* stream);

/* ConvTranspose2d(Cin, Co, 2, 2) + bias written directly into channels [ooff, ooff+Co) of the concat buffer
 * (nn.ConvTranspose2d :86 + F.pad :92-96 + the up half of torch.cat :100).  The buffer holds a fine grid of Ho x Wo
 * pixels per image (0 = exactly 2H x 2W); when the skip tensor is one pixel larger (odd sizes after floor pooling) the
 * reference's F.pad puts the up-sampled map at offset (0,0) and zero-fills the last row / column: the caller zeroes that
 * border with onet_zero_border and passes Ho, Wo.
 * SIMT engine takes the fp32 master weight w[Cin][Co][2][2]; TC engine takes the packed wf / wd. */
int onet_convT2x2_fwd(const void* x, int64_t ldx, int xoff, int N, int H, int W, int Cin, const void* w,
                      const float* bias, int Co, void* out, int64_t ldo, int ooff, int Ho, int Wo, int dtype, int engine,
                      void* stream);
int onet_convT2x2_dgrad(const void* go, int64_t ldg, int goff, int N, int H, int W, int Cin, const void* w, int Co,
                        void* dx, int64_t ldd, int doff, int Ho, int Wo, int dtype, int engine, void* stream);
/* dbias (optional) = column sums of go over the valid 2H x 2W window only */
int onet_convT2x2_wgrad(const void* x, int64_t ldx, int xoff, const void* go, int64_t ldg, int goff, int N, int H,
                        int W, int Cin, int Co, float* dw, float* dbias, int Ho, int Wo, int dtype, int engine, void* stream);
/* zero channels [coff, coff+C) of every pixel with h >= Hv or w >= Wv of a [N,Ho,Wo,ld] buffer (the F.pad border) */
int onet_zero_border(void* buf, int N, int Ho, int Wo, int64_t ld, int coff, int C, int Hv, int Wv, int dtype, void* stream);

/* dst[c] += sums[c] for c < C.  The bias gradient of the transposed convolution is the column sum of d(concat)'s
 * up half; onet_conv3x3_fwd accumulates it (stat_sum) while it writes d(concat), this folds it into the fp32 .grad,
 * and onet_convT2x2_wgrad is then called with dbias = NULL (no separate pass over d(concat)). */
int onet_add_colsums(const double* sums, int C, float* dst, void* stream);

/* Head + loss (Onet.forward :176-189 and compute_loss/jensen_shannon_divergence/log1pexp :221-267) in one
 * bandwidth-bound pass: Vt, Vd, S=softmax([Vt,Vd]), a=sum_p Lt_p, b=sum_p Ld_p, and the sum over pixels of the
 * four piecewise-softplus terms added to *loss_acc (loss = *loss_acc / (2*B*H*W)). */
int onet_head_fwd(const void* L, int64_t ldl, int offl, const void* Hf, int64_t ldh, int offh, int B, int H, int W,
                  float* Vt, float* Vd, float* S, float* a, float* b, double* loss_acc, int dtype, void* stream);
/* Backward: closed-form JSD gradient scaled by *gscale (NULL = none) plus optional external gradients w.r.t.
 * Vt, Vd, S; writes dL and dH as dense [2B,H,W,64]. */
int onet_head_bwd(const void* L, int64_t ldl, int offl, const void* Hf, int64_t ldh, int offh, int B, int H, int W,
                  const float* Vt, const float* Vd, const float* a, const float* b, const float* gscale,
                  const float* gVt, const float* gVd, const float* gS, void* dL, void* dH, int dtype, void* stream);

/* The same head with the LAST layer's BatchNorm + ReLU folded in: Y is that layer's raw conv output [2B,H,W,ldy], the head
 * applies h = relu(y * hscale + hshift) per branch itself (top / down constants, [64] each), so the separate BatchNorm pass
 * over the global feature and the tensor it would write do not exist.  Backward counterpart: onet_head_bwd_scalars (the
 * per-pixel part of onet_head_bwd: gv[0..B*H*W) = dV of the top branch, gv[B*H*W..) of the down branch, gab = the gradient of the
 * loss's channel sums, same layout) followed by onet_bn_relu_bwd_head: the last layer's BatchNorm backward whose incoming
 * gradient is formed on the fly as g = gv * L (never stored) and whose reduce pass also writes dL = gv * relu(bn(y)) + gab.
 * Results are those of onet_head_bwd + onet_bn_relu_bwd bit for bit up to the order of the sums. */
int onet_head_fwd_bn(const void* L, int64_t ldl, int offl, const void* Y, int64_t ldy, int offy, int B, int H, int W,
                     const float* hscale_t, const float* hshift_t, const float* hscale_d, const float* hshift_d,
                     float* Vt, float* Vd, float* S, float* a_out, float* b_out, double* loss_acc, int dtype, void* stream);
int onet_head_bwd_scalars(const float* Vt, const float* Vd, const float* a_in, const float* b_in, const float* gscale,
                          const float* gVt, const float* gVd, const float* gS, int B, int H, int W, float* gv, float* gab,
                          void* stream);
int onet_bn_relu_bwd_head(const void* y, int N, int H, int W, int C, const float* scale, const float* shift, const float* mean,
                          const float* invstd, int group_images, const void* L, int64_t ldl, int offl, const float* gv,
                          const float* gab, void* dL, double* sums, double count, void* dy, float* dgamma0, float* dbeta0,
                          float* dgamma1, float* dbeta1, int dtype, void* stream);

/* argmax of the 2-way softmax: 1 iff Vd > Vt (Onet.predict_label :193-202) */
int onet_predict_label(const float* Vt, const float* Vd, int64_t n, int64_t* out, void* stream);
/* the same decision as one byte per pixel: the mask format of the tiled inference path (onet_b200/infer.py), 8x fewer
 * bytes over PCIe than the int64 argmax of the reference's predict_label */
int onet_predict_label_u8(const float* Vt, const float* Vd, int64_t n, unsigned char* out, void* stream);

/* Evaluation next to the path: counts[pred * 2 + gt] += 1 over n pixels, pred = (Vd > Vt) (predict_label :193-202), gt != 0
 * -> 1.  counts is a zero-initialised int64[4]; accuracy / mIoU / detection rate / false-alarm rate / target IoU of
 * utils_20231218.py:100-234 (evaluate_nau_segmentation_v2, re_assign_label) are functions of these four numbers
 * (onet_b200/evaluate.py), so test_simclutter's per-batch host syncs collapse into one 32-byte read-back. */
int onet_eval_confusion(const float* Vt, const float* Vd, const int64_t* gt, int64_t n, int64_t* counts, void* stream);

/* tensor_normal_per_frame (utils_20231218.py:673-689): x [frames][hw] fp32 -> out[f][i] = (x - min_f) / (max_f - min_f +
 * np.spacing(1)); out may alias x.  work: 2 * frames ints of device scratch (order-preserving min / max keys).  Used by the
 * two-stage cascade (test_2nd_stage_simclutter, Train_Onet_on_simclutter_20250407.py:296-390) between the two Onets. */
int onet_normalize_per_frame(const float* x, int frames, int64_t hw, int* work, float* out, void* stream);

/* ---- on-device synthesis of training frames (Rayleigh_bg_Gaussian_EOT_generator_20230208.py) ------------
 * Counter-based Philox4x32-10: element i of stream `stream_id` under `seed` is the same whatever the launch geometry. */
/* Rayleigh(sigma) amplitudes: the background of get_rayleigh_frame (:219-221, scipy.stats.rayleigh.rvs(scale=sigma)). */
int onet_synth_rayleigh(float* out, int64_t n, float sigma, int64_t seed, int stream_id, void* stream);
/* K-distributed amplitudes as a compound Gaussian: Rayleigh(1) speckle x sqrt(Gamma(nu, 1/nu)) texture, integer nu.  The
 * marginal of the reference's K generator (K_distributed_SeaClutter_Simulation_20210919.py:469-526), uncorrelated. */
int onet_synth_kclutter(float* out, int64_t n, int nu, int64_t seed, int stream_id, void* stream);
/* add_gaussian_template_on_clutter_v3 (:62-176, swerling type 0) for every frame: frames [n_frames][H][W] fp32 updated in
 * place, masks [n_frames][H][W] bytes OR-ed with the target masks (caller zeroes them).  targets: device table of
 * n_frames x targets_per_frame records of 8 x 4 bytes {int lx, ly, wr, hr; float a, b, c, thr} prepared on the host
 * (onet_b200/synth.py) from (cx, cy, w, h, theta); applied in order, as the reference's loop does.  thr = the mask
 * threshold kgauss.max() - 2 * kgauss.std() (:142), or negative to have the kernel reduce it over the window.  erc_out (optional):
 * the average clutter energy mean(bg^2) per frame that scales the target amplitude, kcoef = sqrt(10^(snr/10) * erc). */
int onet_synth_add_targets(float* frames, unsigned char* masks, int n_frames, int H, int W, const void* targets,
                           int targets_per_frame, float snr_db, float* erc_out, void* stream);

/* ---- correlated K-distributed clutter field (K_distributed_SeaClutter_Simulation_20210919.py:469-526), element-wise stages in
 * double precision; the FFTs between them are the caller's (onet_b200/synth.py uses torch.fft). */
/* standard normal white noise (the fields the reference draws with np.random.normal, :483 and :287) */
int onet_synth_normal(double* out, int64_t n, int64_t seed, int stream_id, void* stream);
/* mnlt(x, v) (:83-91): Gamma(shape v, scale 1) quantile of Phi(x), integer v */
int onet_kfield_mnlt(const double* x, int64_t n, int v, double* y, void* stream);
/* coeff_acf_polyn (:121-139): sums[f][n] += sum_i exp(-x^2) H_n(x) g, n = 0, 1, 2, per frame; sums zero-initialised */
int onet_kfield_coeff_sums(const double* x, const double* g, int frames, int64_t per_frame, double* sums, void* stream);
/* solve_acf_polyn (:141-164): out[f][i] = np.roots([a_f, b_f, 1 - acf[i]])[0] as interleaved (re, im); coeffs [frames][2] */
int onet_kfield_acf_root(const double* coeffs, const double* acf, int frames, int64_t per_frame, double* out, void* stream);
/* amplitude (:517-518): out[i] = | speckle[i] | * sqrt(texture[i]); speckle interleaved complex128 */
int onet_kfield_amplitude(const double* speckle, const double* texture, int64_t n, float* out, void* stream);

/* torch.optim.Adam step (no weight decay, amsgrad=False; Train_Onet_on_simclutter_20250407.py:181-182) over a
 * flat fp32 arena; `step` is the 1-based step count, gradients are multiplied by grad_scale first. */
int onet_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale, void* stream);

/* The same update with the step count and the hyper-parameters in DEVICE memory (hyper = {lr, beta1, beta2, eps};
 * *step is incremented first and then used as the 1-based step), so that a launch captured in a CUDA graph stays
 * valid across steps and learning-rate schedules. */
int onet_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int* step,
                       float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONET_B200_H */
