"""Per-kernel parity tests on the GPU (-m gpu): every C-ABI kernel against the matching torch op evaluated in
FP32 (TF32 off) on the same inputs.  The tcgen05 kernels are additionally compared with the CUDA-core kernels."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _imports():
    import gpu_util as U
    return U


def _bf16r(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 1, 64), (3, 12, 20, 3, 64), (2, 8, 8, 64, 128), (1, 5, 7, 16, 32)])
def test_conv3x3_simt_fp32(n, h, w, cin, cout):
    U = _imports()
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    wf, wd = U.pack_conv(wt, U.F32)
    y, st = U.conv3x3(U.to_nhwc(x, torch.float32), wf, cout, U.F32, U.ENGINE_SIMT, group_images=max(n // 2, 1), stats=True)
    ref = F.conv2d(x, wt, padding=1)
    assert U.rel_l2(U.from_nhwc(y), ref) < 1e-6
    # statistics: group 0 = first group_images images
    g = max(n // 2, 1)
    s0 = ref[:g].double().sum(dim=(0, 2, 3))
    q0 = (ref[:g].double() ** 2).sum(dim=(0, 2, 3))
    assert torch.allclose(st[0, 0], s0, rtol=1e-5, atol=1e-4)
    assert torch.allclose(st[1, 0], q0, rtol=1e-5, atol=1e-4)
    if n > g:
        s1 = ref[g:2 * g].double().sum(dim=(0, 2, 3))
        assert torch.allclose(st[0, 1], s1 + (ref[2 * g:].double().sum(dim=(0, 2, 3)) if n > 2 * g else 0), rtol=1e-5, atol=1e-4)
    # dgrad = conv with flipped/transposed weights
    gy = torch.randn_like(ref)
    dx, _ = U.conv3x3(U.to_nhwc(gy, torch.float32), wd, cin, U.F32, U.ENGINE_SIMT)
    dref = torch.nn.grad.conv2d_input(x.shape, wt, gy, padding=1)
    assert U.rel_l2(U.from_nhwc(dx), dref) < 1e-6
    dw = U.conv3x3_wgrad(U.to_nhwc(gy, torch.float32), U.to_nhwc(x, torch.float32), U.F32, U.ENGINE_SIMT)
    wref = torch.nn.grad.conv2d_weight(x, wt.shape, gy, padding=1)
    assert U.rel_l2(dw, wref) < 1e-5


TC_SHAPES = [(2, 16, 16, 64, 64), (4, 16, 16, 128, 256), (2, 32, 32, 64, 128), (8, 4, 4, 256, 256), (4, 2, 2, 128, 64),
             (2, 24, 40, 128, 64), (6, 8, 8, 192, 384), (1, 19, 21, 64, 128),
             (1, 16, 24, 64, 64), (3, 48, 40, 64, 256), (5, 32, 16, 128, 128),   # odd pixel-tile counts for the CTA-pair kernels
             (2, 16, 16, 256, 256), (3, 8, 24, 128, 384),
             (8, 128, 96, 64, 64), (7, 112, 104, 64, 128)]   # >= 4 tiles per SM with Cin = 64: weight-resident kernel


@pytest.mark.parametrize("n,h,w,cin,cout", TC_SHAPES)
def test_conv3x3_tc_fwd(n, h, w, cin, cout):
    U = _imports()
    torch.manual_seed(1)
    x = _bf16r(torch.randn(n, cin, h, w, device="cuda"))
    wt = _bf16r(torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5)
    wf, wd = U.pack_conv(wt, U.BF16)
    xb = U.to_nhwc(x, torch.bfloat16)
    g = max(n // 2, 1)
    y, st = U.conv3x3(xb, wf, cout, U.BF16, U.ENGINE_TC, group_images=g, stats=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, padding=1)
    err = U.rel_l2(U.from_nhwc(y), ref)
    assert err < 4e-3, err                       # bf16 output rounding only (inputs are bf16-exact)
    yf = U.from_nhwc(y).double()
    assert torch.allclose(st[0, 0], yf[:g].sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[1, 0], (yf[:g] ** 2).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    if n > g:
        assert torch.allclose(st[0, 1], yf[g:].sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    # same kernel as dgrad
    gy = _bf16r(torch.randn(n, cout, h, w, device="cuda"))
    dx, _ = U.conv3x3(U.to_nhwc(gy, torch.bfloat16), wd, cin, U.BF16, U.ENGINE_TC)
    dref = torch.nn.grad.conv2d_input(x.shape, wt, gy, padding=1)
    err = U.rel_l2(U.from_nhwc(dx), dref)
    assert err < 4e-3, err


@pytest.mark.parametrize("dt_name", ["fp32", "bf16"])
@pytest.mark.parametrize("n,h,w,cin", [(4, 32, 32, 1), (2, 40, 24, 3), (6, 16, 132, 1), (2, 67, 20, 1), (2, 35, 520, 1), (4, 16, 256, 1),
                                       (2, 19, 276, 3), (4, 32, 32, 3)])
def test_first_layer_recompute_matches_stored_path(dt_name, n, h, w, cin):
    """csrc/first_layer.cuh (the first convolution without materialising Y / dY) against the stored path of the same library:
    conv_first_fwd -> bn_finalize -> bn_relu_apply and bn_relu_bwd -> conv_first_wgrad.  Forward: bit-identical activations,
    equal statistics; backward: same weight / BatchNorm gradients up to summation order."""
    U = _imports()
    call, ptr = U.call, U.ptr
    dt = U.F32 if dt_name == "fp32" else U.BF16
    tdt = U.TDT[dt]
    rnd = (lambda t: t) if dt == U.F32 else _bf16r
    torch.manual_seed(21)
    g = n // 2
    x = rnd(torch.rand(n, cin, h, w, device="cuda"))
    wt = rnd(torch.randn(64, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5)
    gamma = 1.0 + 0.2 * torch.randn(64, device="cuda")
    beta = 0.1 * torch.randn(64, device="cuda")
    xn = U.to_nhwc(x, tdt)
    wf, _ = U.pack_conv(wt, dt)
    count = float(g * h * w)

    def finalize(st):
        aff = torch.empty(4, 2, 64, device="cuda")
        rm, rv = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
        call("onet_bn_finalize", ptr(st[0]), ptr(st[1]), 2, 64, count, ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(gamma), ptr(beta),
             ptr(rm), ptr(rv), 0.1, ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), U.stream())
        return aff
    # stored path
    y, st_ref = U.conv3x3(xn, wf, 64, dt, U.ENGINE_SIMT, group_images=g, stats=True)
    assert U._lib.lib().onet_last_kernel().decode().startswith("conv_first_fwd")
    aff_ref = finalize(st_ref)
    act_ref = torch.empty(n, h, w, 64, dtype=tdt, device="cuda")
    call("onet_bn_relu_apply", ptr(y), n, h, w, 64, ptr(aff_ref[2]), ptr(aff_ref[3]), g, ptr(act_ref), 64, 0, None, None, dt, U.stream())
    # recompute path
    st = torch.zeros(2, 2, 64, dtype=torch.float64, device="cuda")
    call("onet_first_conv_stats", ptr(xn), n, h, w, cin, ptr(wf), None, ptr(st[0]), ptr(st[1]), g, dt, U.stream())
    assert torch.allclose(st, st_ref, rtol=1e-12, atol=1e-9)
    aff = finalize(st)
    act = torch.empty_like(act_ref)
    call("onet_first_conv_bn_relu", ptr(xn), n, h, w, cin, ptr(wf), ptr(aff_ref[2]), ptr(aff_ref[3]), g, ptr(act), 1, dt, U.stream())
    assert torch.equal(act, act_ref)
    assert torch.allclose(aff, aff_ref, rtol=1e-6, atol=1e-7)
    # backward
    gr = rnd(torch.randn(n, 64, h, w, device="cuda"))
    gn = U.to_nhwc(gr, tdt)
    sums_ref = torch.zeros(2, 2, 64, dtype=torch.float64, device="cuda")
    dy = torch.empty(n, h, w, 64, dtype=tdt, device="cuda")
    dgam_ref, dbet_ref = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    call("onet_bn_relu_bwd", ptr(y), n, h, w, 64, ptr(aff_ref[2]), ptr(aff_ref[3]), ptr(aff_ref[0]), ptr(aff_ref[1]), g, ptr(gn), 64, 0,
         None, 0, 0, None, None, ptr(sums_ref), count, ptr(dy), ptr(dgam_ref), ptr(dbet_ref), ptr(dgam_ref), ptr(dbet_ref), dt, U.stream())
    dw_ref = U.conv3x3_wgrad(dy, xn, dt, U.ENGINE_SIMT)
    sums = torch.zeros_like(sums_ref)
    dw = torch.zeros(64, cin, 3, 3, device="cuda")
    dgam, dbet = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    call("onet_first_conv_bwd", ptr(xn), n, h, w, cin, ptr(wf), ptr(aff_ref[2]), ptr(aff_ref[3]), ptr(aff_ref[0]), ptr(aff_ref[1]), g,
         ptr(gn), None, None, ptr(sums), count, ptr(dw), ptr(dgam), ptr(dbet), ptr(dgam), ptr(dbet), dt, U.stream())
    torch.cuda.synchronize()
    assert torch.allclose(sums, sums_ref, rtol=1e-5, atol=1e-4)
    assert U.rel_l2(dgam, dgam_ref) < 1e-5 and U.rel_l2(dbet, dbet_ref) < 1e-5
    # bf16: a 1e-7 difference of a sum can flip the bf16 rounding of single dY elements
    assert U.rel_l2(dw, dw_ref) < (1e-5 if dt == U.F32 else 2e-3), U.rel_l2(dw, dw_ref)
    if cin != 1 and dt != U.BF16:
        return
    # closed form (in_chns = 1: every dtype; in_chns = 3: bf16, warp-level MMAs): statistics from the patch moments, ONE backward
    # pass + assembly; y and dY unrounded, so against the stored bf16 path only up to the rounding it does
    K = 9 * cin
    gram = torch.zeros(2, K + K * K, dtype=torch.float64, device="cuda")
    st2 = torch.zeros(2, 2, 64, dtype=torch.float64, device="cuda")
    call("onet_first_conv_stats", ptr(xn), n, h, w, cin, ptr(wf), ptr(gram), ptr(st2[0]), ptr(st2[1]), g, dt, U.stream())
    yf = F.conv2d(x, wt, padding=1).double()              # unrounded conv output (operands are exact in the storage type)
    ya = F.conv2d(x, wt.abs(), padding=1).double()        # the un-cancelled magnitude: the patch moments are fp32 sums per block,
    for gi, sl in enumerate((slice(0, g), slice(g, n))):  # so the error is relative to sum |w|.v, not to the cancelled sum w.v
        assert ((st2[0, gi] - yf[sl].sum(dim=(0, 2, 3))).abs() <= 2e-6 * ya[sl].sum(dim=(0, 2, 3)) + 1e-4).all()
        assert ((st2[1, gi] - (yf[sl] ** 2).sum(dim=(0, 2, 3))).abs() <= 2e-6 * (ya[sl] ** 2).sum(dim=(0, 2, 3)) + 1e-4).all()
    aff2 = finalize(st2)
    act2 = torch.empty_like(act_ref)
    call("onet_first_conv_bn_relu", ptr(xn), n, h, w, cin, ptr(wf), ptr(aff2[2]), ptr(aff2[3]), g, ptr(act2), 0, dt, U.stream())
    act2_ref = torch.relu(yf.float() * torch.repeat_interleave(aff2[2], torch.tensor([g, n - g], device="cuda"), dim=0)[:, :, None, None]
                          + torch.repeat_interleave(aff2[3], torch.tensor([g, n - g], device="cuda"), dim=0)[:, :, None, None])
    assert U.rel_l2(U.from_nhwc(act2), act2_ref) < (1e-6 if dt == U.F32 else 3e-3)
    if dt == U.BF16:
        # warp-level tensor-core form (first_layer_mma.cuh): products exact, fp32 accumulation -> the bf16 results are the rounded
        # reference except where a 1e-7 difference crosses a rounding boundary (then by one bf16 step)
        got, want = U.from_nhwc(act2), _bf16r(act2_ref)
        diff = got != want
        assert diff.float().mean().item() < 1e-3, diff.float().mean().item()
        assert ((got - want).abs() <= want.abs() * 2.0 ** -7 + 1e-5).all()      # near zero: y * sc + sh cancels
    sums2 = torch.zeros_like(sums_ref)
    dw2 = torch.zeros(64, cin, 3, 3, device="cuda")
    acc_a = torch.zeros(2, 64, K, device="cuda")
    dgam2, dbet2 = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    call("onet_first_conv_bwd", ptr(xn), n, h, w, cin, ptr(wf), ptr(aff2[2]), ptr(aff2[3]), ptr(aff2[0]), ptr(aff2[1]), g,
         ptr(gn), ptr(gram), ptr(acc_a), ptr(sums2), count, ptr(dw2), ptr(dgam2), ptr(dbet2), ptr(dgam2), ptr(dbet2), dt, U.stream())
    # torch reference of the same layer: conv -> BatchNorm (batch statistics per group) -> ReLU, gradient gr
    xr = x.clone()
    wr = wt.clone().requires_grad_(True)
    gam, bet = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = []
    for sl in (slice(0, g), slice(g, n)):
        yy = F.conv2d(xr[sl], wr, padding=1)
        out.append(torch.relu(F.batch_norm(yy, None, None, gam, bet, training=True, eps=1e-5)))
    (torch.cat(out) * gr).sum().backward()
    torch.cuda.synchronize()
    tol = 2e-5 if dt == U.F32 else 2e-3
    assert U.rel_l2(dw2, wr.grad) < tol, U.rel_l2(dw2, wr.grad)
    assert U.rel_l2(dgam2, gam.grad) < tol and U.rel_l2(dbet2, bet.grad) < tol


def _tf32t(t):
    """fp32 -> tf32 by truncation (the low 13 mantissa bits cleared): exactly representable tensor-core operands"""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def _tf32r(t):
    """fp32 -> tf32 by round-to-nearest (ties away), cvt.rna.tf32.f32"""
    return ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


TF32_SHAPES = [(2, 16, 16, 64, 64), (4, 16, 16, 128, 256), (2, 32, 32, 64, 128), (8, 4, 4, 256, 256), (4, 2, 2, 128, 64),
               (2, 24, 40, 128, 64), (1, 19, 21, 64, 128), (3, 48, 40, 64, 256), (5, 32, 16, 128, 128), (2, 16, 16, 256, 256),
               (3, 8, 24, 128, 384), (2, 64, 64, 64, 64)]


@pytest.mark.parametrize("n,h,w,cin,cout", TF32_SHAPES)
def test_conv3x3_tf32_fwd_dgrad_wgrad(n, h, w, cin, cout):
    """tcgen05 kind::tf32 kernels (C-ABI: dtype ONET_F32 + engine ONET_ENGINE_TC) on tf32-exact inputs: products are exact in
    fp32, so only the accumulation order differs from the FP32 torch op."""
    U = _imports()
    torch.manual_seed(11)
    x = _tf32t(torch.randn(n, cin, h, w, device="cuda"))
    wt = _tf32t(torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5)
    wf, wd = U.pack_conv(wt, U.F32)
    g = max(n // 2, 1)
    y, st = U.conv3x3(U.to_nhwc(x, torch.float32), wf, cout, U.F32, U.ENGINE_TC, group_images=g, stats=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, padding=1)
    err = U.rel_l2(U.from_nhwc(y), ref)
    assert err < 2e-5, err
    yf = U.from_nhwc(y).double()
    assert torch.allclose(st[0, 0], yf[:g].sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[1, 0], (yf[:g] ** 2).sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-3)
    if n > g:
        assert torch.allclose(st[0, 1], yf[g:].sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-3)
    gy = _tf32t(torch.randn(n, cout, h, w, device="cuda"))
    dx, _ = U.conv3x3(U.to_nhwc(gy, torch.float32), wd, cin, U.F32, U.ENGINE_TC)
    dref = torch.nn.grad.conv2d_input(x.shape, wt, gy, padding=1)
    err = U.rel_l2(U.from_nhwc(dx), dref)
    assert err < 2e-5, err
    dw = U.conv3x3_wgrad(U.to_nhwc(gy, torch.float32), U.to_nhwc(x, torch.float32), U.F32, U.ENGINE_TC)
    torch.cuda.synchronize()
    wref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), gy, padding=1)
    err = U.rel_l2(dw, wref)
    assert err < 1e-5, err


def test_tf32_operand_semantics():
    """What the tensor core does with fp32 operands that are NOT tf32-exact: the oracle's tf32 emulation (`_Policy`, truncation
    of the low 13 mantissa bits) must be the arithmetic of the hardware.  Compares one convolution on full-precision random
    inputs with the FP32 op on truncated and on round-to-nearest operands; records both and requires the truncation model."""
    U = _imports()
    from parity_record import record
    torch.manual_seed(12)
    n, h, w, cin, cout = 2, 32, 32, 128, 128
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * (2.0 / (9 * cin)) ** 0.5
    wf, _ = U.pack_conv(wt, U.F32)
    y, _ = U.conv3x3(U.to_nhwc(x, torch.float32), wf, cout, U.F32, U.ENGINE_TC)
    got = U.from_nhwc(y)
    e_trunc = U.rel_l2(got, F.conv2d(_tf32t(x), _tf32t(wt), padding=1))
    e_rna = U.rel_l2(got, F.conv2d(_tf32r(x), _tf32r(wt), padding=1))
    e_fp32 = U.rel_l2(got, F.conv2d(x, wt, padding=1))
    print(f"tf32 operands: vs truncated {e_trunc:.2e}, vs round-to-nearest {e_rna:.2e}, vs full fp32 {e_fp32:.2e}")
    record("tf32_operand_semantics", vs_truncated=e_trunc, vs_round_to_nearest=e_rna, vs_fp32=e_fp32)
    assert e_fp32 < 2e-3
    assert e_trunc < 1e-5 and e_trunc < 0.05 * e_rna, (e_trunc, e_rna)      # measured: 3.5e-6 vs 8.8e-4


def test_conv3x3_tc_strided_input():
    """input taken from channels [64,128) of a wider (concat-like) buffer"""
    U = _imports()
    torch.manual_seed(2)
    n, h, w, cin, cout = 2, 16, 16, 64, 64
    buf = _bf16r(torch.randn(n, h, w, 192, device="cuda")).to(torch.bfloat16)
    wt = _bf16r(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
    wf, _ = U.pack_conv(wt, U.BF16)
    y, _ = U.conv3x3(buf, wf, cout, U.BF16, U.ENGINE_TC, ld_in=192, off_in=64, cin=64)
    ref = F.conv2d(buf[..., 64:128].float().permute(0, 3, 1, 2), wt, padding=1)
    assert U.rel_l2(U.from_nhwc(y), ref) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,cout", TC_SHAPES)
def test_conv3x3_tc_wgrad(n, h, w, cin, cout):
    U = _imports()
    torch.manual_seed(3)
    x = _bf16r(torch.randn(n, cin, h, w, device="cuda"))
    gy = _bf16r(torch.randn(n, cout, h, w, device="cuda"))
    dw = U.conv3x3_wgrad(U.to_nhwc(gy, torch.bfloat16), U.to_nhwc(x, torch.bfloat16), U.BF16, U.ENGINE_TC)
    torch.cuda.synchronize()
    wref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), gy, padding=1)
    err = U.rel_l2(dw, wref)
    assert err < 1e-4, err                       # bf16-exact inputs, fp32 accumulate


@pytest.mark.parametrize("pad", [(0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("engine_name", ["simt_fp32", "simt_bf16", "tc", "tc_tf32"])
@pytest.mark.parametrize("n,h,w,cin", [(2, 4, 4, 128), (3, 8, 8, 256), (2, 16, 16, 128), (4, 2, 2, 1024),
                                       (2, 20, 24, 128), (1, 32, 48, 256)])     # partial 16 x 8 tiles: the full-line store path
def test_convT2x2(engine_name, n, h, w, cin, pad):
    """ConvTranspose2d(C, C/2, 2, 2) + bias written into the up half of a concat buffer whose fine grid is
    (2h + ph) x (2w + pw): the reference's F.pad (Onet_vanilla_20240606.py:92-96) for odd skip sizes puts the map at (0,0)
    and zero-fills the last row / column; backward ignores the border."""
    U = _imports()
    call, ptr = U.call, U.ptr
    dt, eng = {"simt_fp32": (U.F32, U.ENGINE_SIMT), "simt_bf16": (U.BF16, U.ENGINE_SIMT), "tc": (U.BF16, U.ENGINE_TC),
               "tc_tf32": (U.F32, U.ENGINE_TC)}[engine_name]
    tdt = U.TDT[dt]
    rnd = _tf32t if engine_name == "tc_tf32" else ((lambda t: t) if dt == U.F32 else _bf16r)
    ph, pw = pad
    ho, wo = 2 * h + ph, 2 * w + pw
    co = cin // 2
    torch.manual_seed(4)
    x = rnd(torch.randn(n, cin, h, w, device="cuda"))
    wt = rnd(torch.randn(cin, co, 2, 2, device="cuda") * (1.0 / cin) ** 0.5)
    b = torch.randn(co, device="cuda") * 0.1
    # forward into the upper half of a concat buffer (border pre-filled with garbage, skip half must stay untouched)
    cat = torch.zeros(n, ho, wo, 2 * co, dtype=tdt, device="cuda")
    cat[..., co:] = 7.0
    xn = U.to_nhwc(x, tdt)
    wf, wd = U.pack_convT(wt, dt)
    call("onet_zero_border", ptr(cat, co), n, ho, wo, 2 * co, 0, co, 2 * h, 2 * w, dt, U.stream())
    call("onet_convT2x2_fwd", ptr(xn), cin, 0, n, h, w, cin, ptr(wf) if eng == U.ENGINE_TC else ptr(wt), ptr(b), co,
         ptr(cat, co), 2 * co, 0, ho, wo, dt, eng, U.stream())
    ref = F.pad(F.conv_transpose2d(x, wt, b, stride=2), [0, pw, 0, ph])
    tol = (5e-6 if engine_name == "tc_tf32" else 2e-6) if dt == U.F32 else 4e-3
    assert U.rel_l2(cat[..., co:].float().permute(0, 3, 1, 2), ref) < tol
    assert float(cat[..., :co].float().abs().max()) == 0.0
    if ph:
        assert float(cat[:, 2 * h:, :, co:].float().abs().max()) == 0.0
    if pw:
        assert float(cat[:, :, 2 * w:, co:].float().abs().max()) == 0.0
    # backward: go lives in the upper half of a [n,ho,wo,2co] gradient buffer; the border gradient is dropped by F.pad
    go_full = rnd(torch.randn(n, co, ho, wo, device="cuda"))
    go = go_full[:, :, :2 * h, :2 * w].contiguous()
    gbuf = torch.zeros(n, ho, wo, 2 * co, dtype=tdt, device="cuda")
    gbuf[..., co:] = go_full.permute(0, 2, 3, 1).to(tdt)
    dx = torch.empty(n, h, w, cin, dtype=tdt, device="cuda")
    call("onet_convT2x2_dgrad", ptr(gbuf, co), 2 * co, 0, n, h, w, cin, ptr(wd) if eng == U.ENGINE_TC else ptr(wt), co,
         ptr(dx), cin, 0, ho, wo, dt, eng, U.stream())
    dref = F.conv2d(go, wt, stride=2)            # adjoint of conv_transpose2d
    assert U.rel_l2(U.from_nhwc(dx), dref) < tol
    dw = torch.zeros(cin, co, 2, 2, device="cuda")
    db = torch.zeros(co, device="cuda")
    call("onet_convT2x2_wgrad", ptr(xn), cin, 0, ptr(gbuf, co), 2 * co, 0, n, h, w, cin, co, ptr(dw), ptr(db), ho, wo, dt, eng,
         U.stream())
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, br, stride=2).backward(go)
    assert U.rel_l2(dw, wr.grad) < (1e-5 if dt == U.F32 else 1e-4)
    assert U.rel_l2(db, br.grad) < 1e-5


@pytest.mark.parametrize("dt_name", ["fp32", "bf16"])
@pytest.mark.parametrize("n,h,w,c,pool", [(4, 8, 8, 64, True), (2, 6, 10, 128, True), (2, 4, 4, 1024, False), (4, 16, 16, 64, False),
                                          (2, 7, 9, 64, True), (6, 12, 20, 256, True)])
def test_bn_relu_fwd_bwd(dt_name, n, h, w, c, pool):
    """conv-output statistics -> finalize -> apply(+pool) and the fused backward, against torch autograd through
    batch_norm(training) -> relu -> (identity skip + max_pool2d)."""
    U = _imports()
    call, ptr = U.call, U.ptr
    dt = U.F32 if dt_name == "fp32" else U.BF16
    tdt = U.TDT[dt]
    rnd = (lambda t: t) if dt == U.F32 else _bf16r
    torch.manual_seed(5)
    g = n // 2
    y = rnd(torch.randn(n, c, h, w, device="cuda") * 1.5 + 0.3)
    gamma = (1 + 0.2 * torch.randn(c, device="cuda"))
    beta = 0.2 * torch.randn(c, device="cuda")
    rm = torch.zeros(c, device="cuda")
    rv = torch.ones(c, device="cuda")
    yn = U.to_nhwc(y, tdt)
    yd = y.double()
    ssum = torch.stack([yd[:g].sum(dim=(0, 2, 3)), yd[g:].sum(dim=(0, 2, 3))])
    ssq = torch.stack([(yd[:g] ** 2).sum(dim=(0, 2, 3)), (yd[g:] ** 2).sum(dim=(0, 2, 3))])
    aff = torch.empty(4, 2, c, device="cuda")
    count = float(g * h * w)
    call("onet_bn_finalize", ptr(ssum), ptr(ssq), 2, c, count, ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(gamma), ptr(beta),
         ptr(rm), ptr(rv), 0.1, ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), U.stream())
    out = torch.zeros(n, h, w, 2 * c, dtype=tdt, device="cuda")         # skip half of a concat buffer
    pl = torch.empty(n, h // 2, w // 2, c, dtype=tdt, device="cuda") if pool else None
    parg = torch.empty(n, h // 2, w // 2, c // 8, dtype=torch.int16, device="cuda") if pool else None   # stored window arg-max
    call("onet_bn_relu_apply", ptr(yn), n, h, w, c, ptr(aff[2]), ptr(aff[3]), g, ptr(out), 2 * c, 0, ptr(pl), ptr(parg), dt,
         U.stream())
    # torch reference, branch by branch (shared BN module called twice)
    bn = torch.nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    yr = y.clone().requires_grad_(True)
    a0 = torch.relu(bn(yr[:g]))
    a1 = torch.relu(bn(yr[g:]))
    act = torch.cat([a0, a1])
    if dt == U.BF16:   # the kernel stores (and pools) the bf16-rounded activation: straight-through rounding in the reference
        act = act + (_bf16r(act) - act).detach()
    tol = 2e-6 if dt == U.F32 else 4e-3
    assert U.rel_l2(out[..., :c].float().permute(0, 3, 1, 2), act) < tol
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    g1 = rnd(torch.randn(n, c, h, w, device="cuda"))
    g2 = rnd(torch.randn(n, c, h, w, device="cuda"))
    loss = (act * g1).sum() + (act * g2).sum()
    gp = None
    if pool:
        pr = F.max_pool2d(act, 2)
        assert U.rel_l2(pl.float().permute(0, 3, 1, 2), pr) < tol
        gp = rnd(torch.randn_like(pr))
        loss = loss + (pr * gp).sum()
    loss.backward()
    sums = torch.zeros(2, 2, c, dtype=torch.float64, device="cuda")
    dy = torch.empty(n, h, w, c, dtype=tdt, device="cuda")
    dgam = torch.zeros(c, device="cuda")
    dbet = torch.zeros(c, device="cuda")
    g1n, g2n = U.to_nhwc(g1, tdt), U.to_nhwc(g2, tdt)
    gpn = U.to_nhwc(gp, tdt) if pool else None
    call("onet_bn_relu_bwd", ptr(yn), n, h, w, c, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(g1n), c, 0,
         ptr(g2n), c, 0, ptr(gpn), ptr(parg), ptr(sums), count, ptr(dy), ptr(dgam), ptr(dbet), ptr(dgam), ptr(dbet), dt,
         U.stream())
    btol = 2e-5 if dt == U.F32 else 8e-3
    assert U.rel_l2(U.from_nhwc(dy), yr.grad) < btol
    assert U.rel_l2(dgam, bn.weight.grad) < btol
    assert U.rel_l2(dbet, bn.bias.grad) < btol
    if pool:        # the stored arg-max and the arg-max recomputed from y route the pooled gradient identically (bit for bit)
        sums2 = torch.zeros_like(sums)
        dy2 = torch.empty_like(dy)
        dg2, db2 = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        call("onet_bn_relu_bwd", ptr(yn), n, h, w, c, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(g1n), c, 0,
             ptr(g2n), c, 0, ptr(gpn), None, ptr(sums2), count, ptr(dy2), ptr(dg2), ptr(db2), ptr(dg2), ptr(db2), dt, U.stream())
        assert torch.equal(dy2, dy)


@pytest.mark.parametrize("B,H,W", [(2, 8, 12), (1, 17, 17), (3, 5, 7)])     # 289 and 105 pixels: B*H*W % 4 != 0
@pytest.mark.parametrize("dt_name", ["fp32", "bf16"])
def test_head_fwd_bwd_vs_numpy_oracle(dt_name, B, H, W):
    U = _imports()
    from oracle import onet_oracle as orc
    call, ptr = U.call, U.ptr
    dt = U.F32 if dt_name == "fp32" else U.BF16
    tdt = U.TDT[dt]
    rnd = (lambda t: t) if dt == U.F32 else _bf16r
    torch.manual_seed(6)
    # local features up to ~1.5 so that a = sum_p L_p reaches the |x| > 37 and > 18 softplus branches
    Lt, Ht, Ld, Hd = [rnd(torch.rand(B, 64, H, W, device="cuda") * s) for s in (1.5, 0.05, 1.2, 0.05)]
    cat0 = torch.zeros(2 * B, H, W, 128, dtype=tdt, device="cuda")
    cat0[:B, ..., :64] = Lt.permute(0, 2, 3, 1).to(tdt)
    cat0[B:, ..., :64] = Ld.permute(0, 2, 3, 1).to(tdt)
    Hf = torch.cat([Ht, Hd]).permute(0, 2, 3, 1).contiguous().to(tdt)
    f32 = dict(dtype=torch.float32, device="cuda")
    Vt, Vd, S = torch.empty(B, 1, H, W, **f32), torch.empty(B, 1, H, W, **f32), torch.empty(B, 2, H, W, **f32)
    a, b = torch.empty(B, H, W, **f32), torch.empty(B, H, W, **f32)
    acc = torch.zeros((), dtype=torch.float64, device="cuda")
    call("onet_head_fwd", ptr(cat0), 128, 0, ptr(Hf), 64, 0, B, H, W, ptr(Vt), ptr(Vd), ptr(S), ptr(a), ptr(b), ptr(acc), dt,
         U.stream())
    r = orc.head_loss_numpy(Lt.cpu().numpy(), Ht.cpu().numpy(), Ld.cpu().numpy(), Hd.cpu().numpy())
    n = B * H * W
    loss = float(acc.item()) / (2 * n)
    assert abs(loss - float(r["loss"])) <= 1e-5 * abs(float(r["loss"]))
    assert U.rel_l2(Vt[:, 0].cpu(), torch.from_numpy(r["Vt"])) < 1e-6
    assert U.rel_l2(S[:, 0].cpu(), torch.from_numpy(r["St"])) < 1e-5
    assert U.rel_l2(S[:, 1].cpu(), torch.from_numpy(r["Sd"])) < 1e-5
    gscale = torch.ones((), **f32)
    dL = torch.empty(2 * B, H, W, 64, dtype=tdt, device="cuda")
    dH = torch.empty(2 * B, H, W, 64, dtype=tdt, device="cuda")
    call("onet_head_bwd", ptr(cat0), 128, 0, ptr(Hf), 64, 0, B, H, W, ptr(Vt), ptr(Vd), ptr(a), ptr(b), ptr(gscale), None,
         None, None, ptr(dL), ptr(dH), dt, U.stream())
    tol = 2e-5 if dt == U.F32 else 4e-3
    for got, key in ((dL[:B], "dLt"), (dH[:B], "dHt"), (dL[B:], "dLd"), (dH[B:], "dHd")):
        assert U.rel_l2(got.float().permute(0, 3, 1, 2).cpu(), torch.from_numpy(r[key])) < tol, key
    # the fused form: the head applies the last layer's BatchNorm + ReLU itself (onet_head_fwd_bn) and the per-pixel part of the
    # backward (onet_head_bwd_scalars) reproduces the dV / d(a) that onet_head_bwd folds into dL and dH
    Y = rnd(torch.randn(2 * B, H, W, 64, device="cuda")).to(tdt)
    sc = 0.5 + torch.rand(2, 64, device="cuda")
    sh = 0.2 * torch.randn(2, 64, device="cuda")
    Hb = torch.empty_like(Y)
    call("onet_bn_relu_apply", ptr(Y), 2 * B, H, W, 64, ptr(sc), ptr(sh), B, ptr(Hb), 64, 0, None, None, dt, U.stream())
    outs = []
    for fused in (False, True):
        o = [torch.empty(B, 1, H, W, **f32), torch.empty(B, 1, H, W, **f32), torch.empty(B, 2, H, W, **f32), torch.empty(B, H, W, **f32),
             torch.empty(B, H, W, **f32), torch.zeros((), dtype=torch.float64, device="cuda")]
        if fused:
            call("onet_head_fwd_bn", ptr(cat0), 128, 0, ptr(Y), 64, 0, B, H, W, ptr(sc[0]), ptr(sh[0]), ptr(sc[1]), ptr(sh[1]),
                 *[ptr(t) for t in o], dt, U.stream())
        else:
            call("onet_head_fwd", ptr(cat0), 128, 0, ptr(Hb), 64, 0, B, H, W, *[ptr(t) for t in o], dt, U.stream())
        outs.append(o)
    for t0, t1 in zip(outs[0][:5], outs[1][:5]):
        assert torch.equal(t0, t1)
    assert abs(float(outs[0][5]) - float(outs[1][5])) <= 1e-9 * abs(float(outs[0][5]))      # block sums meet in double atomics
    Vt2, Vd2, _, a2, b2, _ = outs[1]
    dL2, dH2 = torch.empty_like(dL), torch.empty_like(dH)
    call("onet_head_bwd", ptr(cat0), 128, 0, ptr(Hb), 64, 0, B, H, W, ptr(Vt2), ptr(Vd2), ptr(a2), ptr(b2), ptr(gscale), None,
         None, None, ptr(dL2), ptr(dH2), dt, U.stream())
    gv = torch.empty(2 * n, **f32)
    gab = torch.empty(2 * n, **f32)
    call("onet_head_bwd_scalars", ptr(Vt2), ptr(Vd2), ptr(a2), ptr(b2), ptr(gscale), None, None, None, B, H, W, ptr(gv), ptr(gab),
         U.stream())
    Lall = cat0[..., :64].float()
    dH_from_scalars = (gv.view(2 * B, H, W, 1) * Lall).to(tdt)
    dL_from_scalars = (gv.view(2 * B, H, W, 1) * Hb.float() + gab.view(2 * B, H, W, 1)).to(tdt)
    assert U.rel_l2(dH_from_scalars.float(), dH2.float()) < (1e-6 if dt == U.F32 else 3e-3)
    assert U.rel_l2(dL_from_scalars.float(), dL2.float()) < (1e-6 if dt == U.F32 else 3e-3)


def test_adam_kernel_matches_torch():
    U = _imports()
    call, ptr = U.call, U.ptr
    torch.manual_seed(7)
    n = 100003
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-6, betas=(0.9, 0.999), eps=1e-8)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        call("onet_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), n, 5e-6, 0.9, 0.999, 1e-8, step, 1.0, U.stream())
    assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("dt_name", ["fp32", "bf16"])
def test_pack_all_weights_matches_per_layer_packing(dt_name):
    """One-launch packing of several conv / up-conv weights (ragged channel counts included) is bit-identical to the
    per-layer packing kernels."""
    import ctypes
    U = _imports()
    call, ptr = U.call, U.ptr
    dt = U.F32 if dt_name == "fp32" else U.BF16
    tdt = U.TDT[dt]
    torch.manual_seed(3)
    convs = [torch.randn(co, ci, 3, 3, device="cuda") for co, ci in ((64, 1), (64, 3), (64, 64), (128, 64), (96, 40), (256, 128))]
    ups = [torch.randn(ci, ci // 2, 2, 2, device="cuda") for ci in (128, 256, 72)]
    jobs, want = [], []
    for w in convs:
        co, ci = w.shape[:2]
        wf, wd = torch.zeros(co, 9, ci, dtype=tdt, device="cuda"), torch.zeros(ci, 9, co, dtype=tdt, device="cuda")
        jobs.append((ptr(w), co, ci, 9, ptr(wf), ptr(wd)))
        want.append((wf, wd) + U.pack_conv(w, dt))
    for w in ups:
        ci, co = w.shape[:2]
        wf, wd = torch.zeros(4 * co, ci, dtype=tdt, device="cuda"), torch.zeros(ci, 4 * co, dtype=tdt, device="cuda")
        jobs.append((ptr(w), ci, co, 4, ptr(wf), ptr(wd)))
        want.append((wf, wd) + U.pack_convT(w, dt))
    n = len(jobs)
    vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
    call("onet_pack_all_weights", n, vp(*[j[0] for j in jobs]), ip(*[j[1] for j in jobs]), ip(*[j[2] for j in jobs]),
         ip(*[j[3] for j in jobs]), vp(*[j[4] for j in jobs]), vp(*[j[5] for j in jobs]), dt, U.stream())
    torch.cuda.synchronize()
    for wf, wd, rf, rd in want:
        assert torch.equal(wf.view(-1), rf.view(-1)) and torch.equal(wd.view(-1), rd.view(-1))


def test_dgrad_epilogue_column_sums_feed_upconv_bias_grad():
    """onet_conv3x3_fwd's statistics output used as column sums of d(concat) + onet_add_colsums == the bias gradient
    that a separate reduction over the up half of d(concat) gives."""
    U = _imports()
    call, ptr = U.call, U.ptr
    torch.manual_seed(9)
    n, h, w, cin, cout = 2, 16, 16, 64, 128
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, _ = U.pack_conv(wt, U.BF16)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    cs = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
    # sums only: the tcgen05 epilogue accepts stat_sq = NULL
    call("onet_conv3x3_fwd", ptr(x), cin, 0, n, h, w, cin, ptr(wf), cout, ptr(out), cout, 0, ptr(cs[0]), None, n, U.BF16,
         U.ENGINE_TC, U.stream())
    dbias = torch.full((64,), 0.5, device="cuda")
    call("onet_add_colsums", ptr(cs[0], 64), 64, ptr(dbias), U.stream())
    want = 0.5 + out[..., 64:].float().sum(dim=(0, 1, 2))
    assert torch.allclose(dbias, want, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("n,h,w,c", [(4, 32, 32, 256),      # CTA-pair kernel, 256-wide tiles, fused epilogue reduce
                                      (6, 32, 40, 128),      # CTA-pair kernel, 128-wide tiles (odd tile count per image row)
                                      (150, 32, 32, 64),     # weight-resident kernel (>= 4 tiles per SM), per-row running sums
                                      (2, 8, 8, 64),         # no fused instantiation for this variant: plain launch + standalone reduce
                                      (2, 4, 4, 1024)])      # images smaller than a 16 x 8 tile: generic tap-GEMM + standalone reduce
def test_dgrad_with_fused_bn_backward_reduce(n, h, w, c):
    """onet_conv3x3_dgrad_bnred + onet_bn_relu_bwd_apply against the unfused pair onet_conv3x3_fwd (flipped weights) +
    onet_bn_relu_bwd on the same inputs: identical data gradient (bit for bit), the same BatchNorm-backward sums (the fused
    epilogue reduces the bf16 gradient it stores, exactly what the standalone pass reads back), the same dY / dgamma / dbeta."""
    U = _imports()
    call, ptr = U.call, U.ptr
    torch.manual_seed(11)
    bf = torch.bfloat16
    g = n // 2
    dy_next = (torch.randn(n, h, w, c, device="cuda") * 0.5).to(bf)               # dY of the block's second conv
    wt = torch.randn(c, c, 3, 3, device="cuda") * (0.5 / (3 * c ** 0.5))
    _, wd = U.pack_conv(wt, U.BF16)
    y_prev = (torch.randn(n, h, w, c, device="cuda") * 1.5 + 0.3).to(bf)          # raw output of the block's first conv
    aff = torch.empty(4, 2, c, device="cuda")                                      # mean, invstd, scale, shift per group
    aff[0] = 0.3 + 0.1 * torch.randn(2, c, device="cuda")
    aff[1] = 0.66 + 0.05 * torch.rand(2, c, device="cuda")
    gamma = 1 + 0.2 * torch.randn(2, c, device="cuda")
    aff[2] = gamma * aff[1]
    aff[3] = 0.1 * torch.randn(2, c, device="cuda") - aff[0] * aff[2]
    count = float(g * h * w)
    # unfused
    g_ref, _ = U.conv3x3(dy_next, wd, c, U.BF16, U.ENGINE_TC, group_images=g, stats=False)
    sums_ref = torch.zeros(2, 2, c, dtype=torch.float64, device="cuda")
    dy_ref = torch.empty(n, h, w, c, dtype=bf, device="cuda")
    dgam_ref, dbet_ref = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    call("onet_bn_relu_bwd", ptr(y_prev), n, h, w, c, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(g_ref), c, 0,
         None, 0, 0, None, None, ptr(sums_ref), count, ptr(dy_ref), ptr(dgam_ref), ptr(dbet_ref), ptr(dgam_ref), ptr(dbet_ref),
         U.BF16, U.stream())
    # fused
    g_fused = torch.empty(n, h, w, c, dtype=bf, device="cuda")
    sums = torch.zeros(2, 2, c, dtype=torch.float64, device="cuda")
    call("onet_conv3x3_dgrad_bnred", ptr(dy_next), c, 0, n, h, w, c, ptr(wd), c, ptr(g_fused), ptr(y_prev), ptr(aff[2]), ptr(aff[3]),
         ptr(aff[0]), ptr(aff[1]), ptr(sums), g, U.BF16, U.ENGINE_TC, U.stream())
    dy = torch.empty(n, h, w, c, dtype=bf, device="cuda")
    dgam, dbet = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    call("onet_bn_relu_bwd_apply", ptr(y_prev), n, h, w, c, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), g, ptr(g_fused), c, 0,
         ptr(sums), count, ptr(dy), ptr(dgam), ptr(dbet), ptr(dgam), ptr(dbet), U.BF16, U.stream())
    torch.cuda.synchronize()
    assert torch.equal(g_fused, g_ref)
    scale = sums_ref.abs().max(dim=2, keepdim=True)[0].clamp_min(1e-30)
    assert float(((sums - sums_ref).abs() / scale).max()) < 2e-5, ((sums - sums_ref).abs() / scale).max()
    assert U.rel_l2(dy.float(), dy_ref.float()) < 5e-4          # sums differ by ~1e-6: a few bf16 roundings flip
    assert U.rel_l2(dgam, dgam_ref) < 1e-5 and U.rel_l2(dbet, dbet_ref) < 1e-5
