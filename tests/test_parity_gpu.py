"""Parity of the reduced-precision CUDA modes, held TIGHTLY (-m gpu).

The FP32 reference cannot pin a bf16 / tf32 run tightly: this network's gradient at random init is ill-conditioned (the FP32
oracle with inputs perturbed by 1e-6 / 1e-4 relative already disagrees with itself by 2.7e-3 / 5e-2 in gradient rel-L2), so
rounding noise is amplified and a loose bound would also pass a wiring bug.  These tests therefore compare each mode with the
oracle's restatement of the SAME algorithm with the SAME roundings (`oracle.onet_oracle.onet_forward(..., emulate=mode)`,
`_Policy`).  Free-running, that pins the forward tightly (loss 4e-6, activations 2e-3); the GRADIENT it does not - measured:
two CUDA implementations with identical rounding points (tcgen05 vs CUDA cores) still differ by 0.16, because one flipped bf16
rounding is a 4e-3 perturbation of an ill-conditioned forward.  The gradient is therefore pinned TEACHER-FORCED: the oracle is
run on the CUDA run's own stored convolution outputs (`_forced`), which shares every forward decision (BatchNorm statistics,
ReLU masks, pool arg-max) and leaves the backward pass as the linear map it is - there a wrong sign, scale, stream dependency
or a missed rounding point shows up as a gross difference, and every layer's forward is checked on identical inputs.  Shapes: B=4 at 128x128 and the BASELINE shape 1x256x256 (B=2).  The
north_star tolerances against the FP32 oracle (loss / activations 1e-2, masks >= 99.9 %, gradients 2e-2) are asserted
literally for the tf32 mode where they hold, and reported with the measured conditioning floor where they cannot (bf16
gradients).  Every measured value is written to gpurun_out/r2_parity.json (-> profiles/).
"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from parity_record import record

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _build(st, mode, use_tc=True, st_d=None):
    import onet_b200
    cin = st["inc.double_conv.0.weight"].shape[1]
    net = onet_b200.Onet(cin, True, st_d is None, mode=mode, use_tc=use_tc)
    sd = OrderedDict()
    for k, v in st.items():
        sd["topu." + k] = v.clone()
    for k, v in (st if st_d is None else st_d).items():
        sd["dwnu." + k] = v.clone()
    net.load_state_dict(sd)
    return net.cuda()


def _step(net, x):
    net.train()
    net.zero_grad()
    Lt, Vt, Ld, Vd, S = net(x)
    loss = net.compute_loss(Lt, S[:, 0, :, :].unsqueeze(dim=1), Ld, S[:, 1, :, :].unsqueeze(dim=1))
    loss.backward()
    torch.cuda.synchronize()
    return dict(Lt=Lt, Vt=Vt, Ld=Ld, Vd=Vd, S=S, loss=loss)


def _compare(net, got, ref_out, ref_grads):
    from oracle import onet_oracle as orc
    m = dict(loss=abs(got["loss"].item() - ref_out["loss"].item()) / abs(ref_out["loss"].item()))
    for n in ("Lt", "Ld", "Vt", "Vd", "S"):
        m[n] = _rel(got[n].float(), ref_out[n])
    gerr = {k: _rel(p.grad, ref_grads[k[len("topu."):]]) for k, p in net.named_parameters()}
    worst = max(gerr, key=gerr.get)
    m.update(grad_median=float(np.median(list(gerr.values()))), grad_worst=gerr[worst], grad_worst_tensor=worst)
    lab = net.predict_label(got["S"]).cpu()
    m["mask_agreement"] = float((lab == orc.predict_label(ref_out["S"])).float().mean())
    return m


def _cuda_taps(net, B):
    """What the CUDA forward stored, keyed like the oracle's taps: every raw convolution output Y and every up-conv output, per
    branch, as (B,C,h,w) fp32 on the CPU.  Weight-shared twin: images [0,B) of the twin batch are the top branch."""
    rec = net._last["rec"]
    from oracle import onet_oracle as orc
    names = [orc._dc_prefix(b) + f".{i}.raw" for b in ["inc"] + [n for n, _, _ in orc.ENCODER] + [n for n, _, _ in orc.DECODER]
             for i in (0, 3)]
    out = {"top": {}, "dwn": {}}
    for li, key in enumerate(names):
        sv = rec.saved[(0, li)]
        Y = sv["Y"]
        if Y is None and sv.get("gram") is not None:
            continue         # in_chns = 1: y is recomputed in fp32 from the patch and never rounded - the oracle's own value is the same
        if Y is None:        # first convolution (two-pass form): its raw output is not stored - the stored-path kernel gives it
            import gpu_util as U
            src = sv["src"][0]
            dt = U.BF16 if src.dtype == torch.bfloat16 else U.F32
            Y, _ = U.conv3x3(src, sv["wf"], 64, dt, U.ENGINE_SIMT)
        Y = Y.float().permute(0, 3, 1, 2).cpu()
        out["top"][key], out["dwn"][key] = Y[:B].contiguous(), Y[B:].contiguous()
    cs = [64, 128, 256, 512]
    for j, (name, _, _) in enumerate(orc.DECODER):
        k = 3 - j
        up = rec.cat[k][:, :2 * rec.hs[k + 1], :2 * rec.ws[k + 1], cs[k]:].float().permute(0, 3, 1, 2).cpu()
        out["top"][f"{name}.up.out"], out["dwn"][f"{name}.up.out"] = up[:B].contiguous(), up[B:].contiguous()
    return out


def _teacher_forced(net, st, x, got, mode, tag):
    """Oracle (with the mode's roundings) teacher-forced with the CUDA run's stored Y / up-conv outputs: per-layer forward
    error on IDENTICAL inputs, and the gradient of the same linearised backward pass."""
    from oracle import onet_oracle as orc
    B = x.shape[0]
    force = _cuda_taps(net, B)
    taps = {}
    out, grads, _ = orc.train_step_outputs(st, x, emulate=None if mode == "fp32" else mode, force=force, taps=taps)
    layer = {}
    for br in ("top", "dwn"):
        for key, forced in force[br].items():
            layer[f"{br}.{key}"] = _rel(forced, taps[br][key + ".own"])
    m = _compare(net, got, out, grads)
    wl = max(layer, key=layer.get)
    m.update(layer_fwd_worst=layer[wl], layer_fwd_worst_name=wl, layer_fwd_median=float(np.median(list(layer.values()))))
    print(f"{tag} teacher-forced: {m}")
    record(f"{tag}_teacher_forced", **m)
    return m


SHAPES = [(4, 128, 31), (2, 256, 77)]       # (B, H = W, seed); 256 x 256 is the BASELINE shape

_CACHE = {}


def _oracle(B, HW, seed, emulate):
    """(state, x, outputs, grads) of the CPU oracle, computed once per (shape, emulation)."""
    from oracle import onet_oracle as orc
    key = (B, HW, seed, emulate)
    if key not in _CACHE:
        torch.set_num_threads(os.cpu_count() or 1)
        st = orc.perturb_bn_affine(orc.init_state(1, seed=seed), seed=seed + 100)
        x = orc.rayleigh_frames(B, 1, HW, HW, seed=seed)
        out, grads, _ = orc.train_step_outputs(st, x, emulate=emulate)
        _CACHE[key] = (st, x, out, grads)
    return _CACHE[key]


@pytest.mark.parametrize("B,HW,seed", SHAPES)
def test_fp32_mode_against_oracle_at_full_size(B, HW, seed):
    """FP32 verification mode vs the FP32 oracle at 128 x 128 and at the BASELINE shape: the north_star 1e-5 / 2e-2 / 99.9 %."""
    st, x, out, grads = _oracle(B, HW, seed, None)
    net = _build(st, "fp32")
    m = _compare(net, _step(net, x.cuda()), out, grads)
    print(f"fp32 B={B} {HW}x{HW}: {m}")
    record(f"fp32_vs_fp32_oracle[B{B}_{HW}]", **m)
    assert m["loss"] < 1e-5
    assert max(m[n] for n in ("Lt", "Ld", "Vt", "Vd", "S")) < 2e-5
    assert m["grad_worst"] < 2e-2 and m["grad_median"] < 5e-3
    assert m["mask_agreement"] >= 0.999


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("B,HW,seed", SHAPES)
def test_bf16_mode_against_bf16_emulating_oracle(B, HW, seed, use_tc):
    """The benchmarked mode (bf16 storage + operands, tcgen05 kernels; use_tc=False = the same roundings on CUDA cores) vs the
    oracle restated with bf16 roundings at the same points.  Also reports the same run against the FP32 oracle."""
    st, x, out_e, grads_e = _oracle(B, HW, seed, "bf16")
    _, _, out_f, grads_f = _oracle(B, HW, seed, None)
    net = _build(st, "bf16", use_tc=use_tc)
    got = _step(net, x.cuda())
    m = _compare(net, got, out_e, grads_e)
    mf = _compare(net, got, out_f, grads_f)
    floor = {k: _rel(grads_e[k], grads_f[k]) for k in grads_f}        # bf16 rounding alone, on the CPU
    print(f"bf16 tc={use_tc} B={B} {HW}x{HW} vs emulated-bf16 oracle: {m}")
    print(f"   the same run vs the FP32 oracle: {mf}; emulated-bf16 vs FP32 oracle gradient median "
          f"{np.median(list(floor.values())):.2e}")
    record(f"bf16_vs_emulated_bf16_oracle[B{B}_{HW}_tc{int(use_tc)}]", **m)
    record(f"bf16_vs_fp32_oracle[B{B}_{HW}_tc{int(use_tc)}]", emulated_floor_grad_median=float(np.median(list(floor.values()))), **mf)
    # (1) free-running forward against the bf16-emulating oracle: tight
    assert m["loss"] < 1e-4
    assert max(m[n] for n in ("Lt", "Ld", "Vt", "Vd")) < 5e-3 and m["S"] < 1e-2
    assert m["mask_agreement"] >= 0.997       # measured 0.9980-0.9981: pixels with |Vt - Vd| below one bf16 ulp of their terms
    # (2) free-running gradient: NOT a tight quantity for bf16 in this network.  Two CUDA implementations with the SAME rounding
    # points (tcgen05 vs CUDA-core kernels, different fp32 accumulation order only) differ by 0.16 median / 0.27 worst, because
    # an accumulation-order flip of one bf16 rounding is a 4e-3 relative perturbation of an ill-conditioned forward pass (module
    # docstring).  The bound below only guards against gross errors; the wiring is pinned by (3).
    assert m["grad_median"] < 0.3 and m["grad_worst"] < 0.5, (m["grad_worst_tensor"], m["grad_worst"])
    # (3) teacher-forced: the oracle runs on the CUDA run's own stored conv outputs, so every forward decision is shared and the
    # backward is compared as a linear map - tight per tensor; plus every layer's forward on identical inputs
    tf = _teacher_forced(net, st, x, got, "bf16", f"bf16[B{B}_{HW}_tc{int(use_tc)}]")
    assert tf["layer_fwd_worst"] < 5e-3, (tf["layer_fwd_worst_name"], tf["layer_fwd_worst"])
    assert tf["loss"] < 1e-5 and max(tf[n] for n in ("Lt", "Ld", "Vt", "Vd")) < 2e-3
    # (measured: median 6e-3..7e-3; worst 2.0e-2..2.4e-2, always an up-conv bias = a sum of bf16-rounded values with cancellation)
    assert tf["grad_median"] < 1.5e-2 and tf["grad_worst"] < 5e-2, (tf["grad_worst_tensor"], tf["grad_worst"])
    assert tf["mask_agreement"] >= 0.999
    # north_star's bf16 tolerance on loss and activations against the FP32 oracle
    assert mf["loss"] < 1e-2 and max(mf[n] for n in ("Lt", "Vt", "Vd")) < 1e-2 and mf["Ld"] < 2e-2


@pytest.mark.parametrize("B,HW,seed", SHAPES)
def test_tf32_mode_meets_the_stated_tolerances(B, HW, seed):
    """The tf32 mode (fp32 storage, tcgen05 kind::tf32 operands) against the FP32 oracle with the north_star tolerances asserted
    LITERALLY - loss and activations 1e-2 relative, masks >= 99.9 % - and against the tf32-emulating oracle, free-running and
    teacher-forced.  Free-running gradients: the stated 2e-2 is below this network's conditioning floor at random init (the
    tf32-emulating CPU oracle itself is 8e-2 from the FP32 oracle); asserted against that measured floor, and tightly
    (teacher-forced) against the emulating oracle."""
    st, x, out_f, grads_f = _oracle(B, HW, seed, None)
    _, _, out_e, grads_e = _oracle(B, HW, seed, "tf32")
    net = _build(st, "tf32")
    got = _step(net, x.cuda())
    mf = _compare(net, got, out_f, grads_f)
    me = _compare(net, got, out_e, grads_e)
    floor = {k: _rel(grads_e[k], grads_f[k]) for k in grads_f}
    fl_med, fl_max = float(np.median(list(floor.values()))), max(floor.values())
    print(f"tf32 B={B} {HW}x{HW} vs FP32 oracle: {mf}")
    print(f"   vs tf32-emulating oracle: {me}; emulated-tf32 vs FP32 oracle gradient median {fl_med:.2e} worst {fl_max:.2e}")
    record(f"tf32_vs_fp32_oracle[B{B}_{HW}]", emulated_floor_grad_median=fl_med, emulated_floor_grad_worst=fl_max, **mf)
    record(f"tf32_vs_emulated_tf32_oracle[B{B}_{HW}]", **me)
    # north_star, literally
    assert mf["loss"] < 1e-2 and max(mf[n] for n in ("Lt", "Ld", "Vt", "Vd", "S")) < 1e-2
    assert mf["mask_agreement"] >= 0.999
    # gradient against the FP32 oracle: within 1.5x of what tf32 operand rounding alone does to the oracle (the conditioning floor)
    assert mf["grad_median"] < 1.5 * fl_med + 1e-3 and mf["grad_worst"] < 1.5 * fl_max + 1e-3
    tf = _teacher_forced(net, st, x, got, "tf32", f"tf32[B{B}_{HW}]")
    assert tf["layer_fwd_worst"] < 1e-4, (tf["layer_fwd_worst_name"], tf["layer_fwd_worst"])
    assert tf["loss"] < 1e-5 and max(tf[n] for n in ("Lt", "Ld", "Vt", "Vd")) < 1e-4
    assert tf["grad_worst"] < 5e-3, (tf["grad_worst_tensor"], tf["grad_worst"])
    assert tf["mask_agreement"] >= 0.9999


def test_bf16_tensor_core_path_matches_cuda_core_path():
    """Wiring of the tcgen05 composition (two-stream order, fused BatchNorm reduce, column-sum bias gradient) against the
    CUDA-core kernels with the same bf16 roundings, per parameter tensor."""
    st, x, _, _ = _oracle(4, 128, 31, "bf16")
    grads = {}
    for use_tc in (True, False):
        net = _build(st, "bf16", use_tc=use_tc)
        got = _step(net, x.cuda())
        grads[use_tc] = ({k: p.grad.clone() for k, p in net.named_parameters()}, got["loss"].item())
    errs = {k: _rel(grads[True][0][k], grads[False][0][k]) for k in grads[True][0]}
    worst = max(errs, key=errs.get)
    print(f"tc vs simt (bf16): loss {grads[True][1]:.7f} / {grads[False][1]:.7f}, gradient median "
          f"{np.median(list(errs.values())):.2e}, worst {worst} {errs[worst]:.2e}")
    record("bf16_tc_vs_bf16_simt[B4_128]", grad_median=float(np.median(list(errs.values()))), grad_worst=errs[worst],
           grad_worst_tensor=worst, loss_rel=abs(grads[True][1] - grads[False][1]) / abs(grads[False][1]))
    assert abs(grads[True][1] - grads[False][1]) <= 1e-4 * abs(grads[False][1])
    # same rounding points, different fp32 accumulation order: this IS the noise floor of a free-running bf16 gradient of this
    # network (measured 0.16 median / 0.27 worst); the tight per-tensor check is the teacher-forced one above
    assert np.median(list(errs.values())) < 0.3 and errs[worst] < 0.5, (worst, errs[worst])


def test_fused_bn_reduce_matches_separate_reduce(monkeypatch):
    """ONET_NO_BNRED_FUSION=1 (separate BatchNorm-backward reduce pass) vs the reduce folded into the dgrad epilogue: same
    sums up to fp32 summation order."""
    st, x, _, _ = _oracle(4, 128, 31, "bf16")
    out = {}
    for fused in (False, True):
        if not fused:
            monkeypatch.setenv("ONET_NO_BNRED_FUSION", "1")
        else:
            monkeypatch.delenv("ONET_NO_BNRED_FUSION", raising=False)
        net = _build(st, "bf16")
        _step(net, x.cuda())
        out[fused] = {k: p.grad.clone() for k, p in net.named_parameters()}
    errs = {k: _rel(out[True][k], out[False][k]) for k in out[True]}
    worst = max(errs, key=errs.get)
    print(f"fused vs separate BN reduce: worst {worst} {errs[worst]:.2e}")
    record("bnred_fused_vs_separate[B4_128]", grad_worst=errs[worst], grad_worst_tensor=worst)
    assert errs[worst] < 2e-2, (worst, errs[worst])


def test_fp32_mode_is_reproducible():
    """Two FP32-mode steps on the same input give bit-identical losses and gradients: BatchNorm statistics are summed in a
    fixed order and the weight-gradient split-K partial sums are added in split order (onet_set_splitk_workspace)."""
    from oracle import onet_oracle as orc
    st = orc.perturb_bn_affine(orc.init_state(1, seed=5), seed=105)
    x = orc.rayleigh_frames(3, 1, 48, 64, seed=5).cuda()
    net = _build(st, "fp32")
    a = _step(net, x)
    ga = {k: p.grad.clone() for k, p in net.named_parameters()}
    b = _step(net, x)
    assert a["loss"].item() == b["loss"].item()
    diff = {k: float((p.grad - ga[k]).abs().max()) for k, p in net.named_parameters()}
    conv = {k: v for k, v in diff.items() if "double_conv.0.weight" in k or "double_conv.3.weight" in k or ".up." in k}
    print("fp32 reproducibility: max |g1 - g2| over conv / up-conv tensors", max(conv.values()), "over all", max(diff.values()))
    record("fp32_reproducible", conv_max_abs_diff=max(conv.values()), all_max_abs_diff=max(diff.values()))
    assert max(conv.values()) == 0.0, conv
    assert max(diff.values()) <= 1e-9


def test_eval_mode_forward_is_differentiable():
    """eval mode with autograd enabled (frozen-BatchNorm fine-tuning) returns differentiable outputs like the
    reference; gradients = the oracle's with BatchNorm in eval mode, and the running buffers do not move."""
    from oracle import onet_oracle as orc
    st = orc.perturb_bn_affine(orc.init_state(1, seed=9), seed=109)
    g = torch.Generator().manual_seed(3)
    for k in st:                                     # non-trivial running statistics
        if k.endswith("running_mean"):
            st[k] = 0.1 * torch.randn(st[k].shape, generator=g)
        if k.endswith("running_var"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=g)
    x = orc.rayleigh_frames(2, 1, 32, 32, seed=9)
    ref = OrderedDict((k, v.clone()) for k, v in st.items())
    leaves = [k for k, v in ref.items() if v.dtype.is_floating_point and "running" not in k]
    for k in leaves:
        ref[k].requires_grad_(True)
    Lt, Vt, Ld, Vd, S = orc.onet_forward(ref, x, training=False)
    loss_ref = orc.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss_ref.backward()
    net = _build(st, "fp32")
    net.eval()
    before = {k: v.clone() for k, v in net.state_dict().items() if "running" in k}
    net.zero_grad()
    out = net(x.cuda())
    loss = net.compute_loss(out[0], out[4][:, 0:1], out[2], out[4][:, 1:2])
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    errs = {k: _rel(p.grad, ref[k[len("topu."):]].grad) for k, p in net.named_parameters()}
    worst = max(errs, key=errs.get)
    print(f"eval-mode gradients vs oracle: median {np.median(list(errs.values())):.2e} worst {worst} {errs[worst]:.2e}")
    record("eval_mode_grad_vs_oracle", grad_median=float(np.median(list(errs.values()))), grad_worst=errs[worst])
    assert errs[worst] < 2e-2, (worst, errs[worst])
    for k, v in net.state_dict().items():
        if "running" in k:
            assert torch.equal(v, before[k]), k


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_eval_between_graph_replays_sees_current_weights(mode):
    """a replayed graph step updates the weights through raw pointers; every eval forward that follows
    must run on the updated weights, not on operand copies packed before the update."""
    import onet_b200
    from onet_b200.data import rayleigh_target_frames
    from onet_b200.model import invalidate_packed_weights
    from onet_b200.trainer import OnetTrainer
    torch.manual_seed(4)
    net = onet_b200.Onet(1, True, True, mode=mode).cuda()
    tr = OnetTrainer(net, lr=1e-3, graph=True)         # large steps: stale weights would be visible
    xs = [rayleigh_target_frames(2, 1, 32, 32, seed=60 + i).cuda() for i in range(3)]
    probe = rayleigh_target_frames(2, 1, 32, 32, seed=99).cuda()

    def eval_vt():
        net.eval()
        with torch.no_grad():
            return net(probe)[1].clone()
    outs = []
    for x in xs:
        tr.step(x)
        v = eval_vt()
        invalidate_packed_weights()
        v_fresh = eval_vt()
        assert torch.equal(v, v_fresh)
        outs.append(v)
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])


def test_cached_eval_affine_follows_running_statistics_and_weights():
    """The eval-mode BatchNorm affine (scale, shift) is cached per module between pure-inference forwards.  The cache must notice
    what does not bump a tensor version: running statistics updated by a training-mode forward (raw-pointer kernel writes), and
    must notice in-place parameter updates (torch optimizers) through the version counters."""
    import onet_b200
    from onet_b200 import model as M
    from onet_b200.data import rayleigh_target_frames
    torch.manual_seed(8)
    net = onet_b200.Onet(1, True, True, mode="bf16").cuda()
    x = rayleigh_target_frames(2, 1, 32, 32, seed=5).cuda()

    def eval_vt(fresh=False):
        if fresh:
            M._EVAL_AFFINE.clear()
        net.eval()
        with torch.no_grad():
            return net(x)[1].clone()
    v0 = eval_vt()
    assert torch.equal(v0, eval_vt()) and len(M._EVAL_AFFINE) > 0           # second call served from the cache
    net.train()
    with torch.no_grad():
        net(x)                                                              # updates the running statistics only
    v1 = eval_vt()
    assert not torch.equal(v1, v0) and torch.equal(v1, eval_vt(fresh=True))
    with torch.no_grad():
        net.topu.inc.double_conv["1"].weight.mul_(1.5)                        # in-place parameter update: version counter
    v2 = eval_vt()
    assert not torch.equal(v2, v1) and torch.equal(v2, eval_vt(fresh=True))
