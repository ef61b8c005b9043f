"""Pin the CPU oracle (oracle/onet_oracle.py) against golden vectors produced by the UNMODIFIED
reference module (tests/golden/make_golden.py) and, when /root/reference is present, against the
live reference.  CPU only."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import onet_oracle as orc
from oracle.ref_import import reference_available

CASES = ["c1_b2_32x32", "c3_b1_48x32", "c1_b2_32x32_noshare", "c1_b2_40x56_pad"]
SAMPLE_STRIDE = 9973


def _states(meta):
    cin, b, h, w, bshare, seed = (int(v) for v in meta)
    st = orc.perturb_bn_affine(orc.init_state(cin, seed=seed), seed=seed + 100)
    st_d = None if bshare else orc.perturb_bn_affine(orc.init_state(cin, seed=seed + 50), seed=seed + 150)
    return st, st_d


def _check_summary(name, t, g, rtol):
    a = t.detach().numpy().astype(np.float64)
    norm = np.sqrt((a ** 2).sum())
    assert abs(norm - g[f"{name}.norm"]) <= rtol * max(g[f"{name}.norm"], 1e-12), name
    if f"{name}.full" in g:
        ref = g[f"{name}.full"].astype(np.float64)
        assert np.linalg.norm(a - ref) <= rtol * max(np.linalg.norm(ref), 1e-12), name
    else:
        ref = g[f"{name}.sample"].astype(np.float64)
        got = a.reshape(-1)[::SAMPLE_STRIDE]
        assert np.linalg.norm(got - ref) <= rtol * max(np.linalg.norm(ref), 1e-12), name


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    st, st_d = _states(g["meta"])
    x = torch.from_numpy(g["x"])
    # the input generator itself is part of the fixture contract
    cin, b, h, w, _, seed = (int(v) for v in g["meta"])
    assert torch.equal(orc.rayleigh_frames(b, cin, h, w, seed=seed), x)
    res = orc.train_step_outputs(st, x, st_d)
    out, grads, new_state = res[0], res[1], res[2]
    # same ATen CPU kernels underneath -> tight tolerance (summation order may differ slightly)
    assert abs(out["loss"].item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    for k in ("Vt", "Vd", "S"):
        ref = torch.from_numpy(g[k])
        assert torch.linalg.norm(out[k] - ref) <= 2e-5 * torch.linalg.norm(ref), k
    _check_summary("Lt", out["Lt"], g, 2e-5)
    _check_summary("Ld", out["Ld"], g, 2e-5)
    for k, gr in grads.items():
        _check_summary("grad.topu." + k, gr, g, 5e-3)   # fp32 noise floor: 18 BN layers, deepest BN sees only B*2*2 values
    if st_d is not None:
        for k, gr in res[3].items():
            _check_summary("grad.dwnu." + k, gr, g, 5e-3)
    for k, v in new_state.items():
        if "running" in k:
            ref = g["buf.topu." + k]
            assert np.allclose(v.numpy(), ref, rtol=2e-5, atol=1e-6), k
        if "num_batches" in k:
            assert int(v) == int(g["buf.topu." + k]), k   # shared twin: += 2 per step
    # eval-mode forward with the updated running statistics, and labels
    st_eval = new_state
    sd_eval = res[4] if st_d is not None else None
    with torch.no_grad():
        Lt, Vt, Ld, Vd, S = orc.onet_forward(st_eval, x, training=False, st_dwn=sd_eval)
    for k, t in (("eval.Vt", Vt), ("eval.Vd", Vd), ("eval.S", S)):
        ref = torch.from_numpy(g[k])
        assert torch.linalg.norm(t - ref) <= 2e-5 * torch.linalg.norm(ref), k
    lab = orc.predict_label(S).numpy().astype(np.uint8)
    assert (lab == g["eval.label"]).mean() >= 0.9995


def test_log1pexp_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "log1pexp_kat.npz"))
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = orc.log1pexp(x)
    y.sum().backward()
    assert np.array_equal(y.detach().numpy(), g["y"])          # bit-exact, incl. ln2 below -37
    assert np.allclose(x.grad.numpy(), g["dy"], rtol=1e-6, atol=1e-12)
    # numpy restatement (value + closed-form derivative)
    v, d = orc._sp_np(g["x"])
    assert np.allclose(v, g["y"], rtol=1e-6, atol=1e-12)
    assert np.allclose(d, g["dy"], rtol=1e-5, atol=1e-12)


def test_numpy_head_matches_autograd():
    torch.manual_seed(3)
    B, C, H, W = 2, 64, 8, 12
    # scale so that a*S spans all softplus branches (a up to ~70 as at random init)
    Lt, Ht, Ld, Hd = [(torch.rand(B, C, H, W) * s).requires_grad_(True) for s in (1.5, 0.05, 1.2, 0.05)]
    Vt = (Lt * Ht).sum(1, keepdim=True)
    Vd = (Ld * Hd).sum(1, keepdim=True)
    S = torch.softmax(torch.cat([Vt, Vd], 1), 1)
    loss = orc.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss.backward()
    r = orc.head_loss_numpy(Lt.detach().numpy(), Ht.detach().numpy(), Ld.detach().numpy(), Hd.detach().numpy())
    assert abs(r["loss"] - loss.item()) <= 2e-6 * abs(loss.item())
    for k, t in (("dLt", Lt), ("dHt", Ht), ("dLd", Ld), ("dHd", Hd)):
        ref = t.grad.numpy()
        assert np.linalg.norm(r[k] - ref) <= 2e-5 * np.linalg.norm(ref), k


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    from oracle.ref_import import import_reference
    ref = import_reference()
    st = orc.perturb_bn_affine(orc.init_state(1, seed=21), seed=22)
    x = orc.rayleigh_frames(2, 1, 32, 48, seed=23)
    onet = ref.Onet(1, True, True)
    sd = OrderedDict()
    for k, v in st.items():
        sd["topu." + k] = v.clone()
        sd["dwnu." + k] = v.clone()
    onet.load_state_dict(sd)
    onet.train()
    Lt, Vt, Ld, Vd, S = onet(x)
    loss = onet.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss.backward()
    out, grads, _ = orc.train_step_outputs(st, x)
    assert abs(out["loss"].item() - loss.item()) <= 2e-6 * abs(loss.item())
    for k, p in onet.named_parameters():
        gk = grads[k[len("topu."):]]
        assert torch.linalg.norm(gk - p.grad) <= 5e-3 * torch.linalg.norm(p.grad) + 1e-12, k


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(100)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-6, betas=(0.9, 0.999), eps=1e-8)
    m = torch.zeros(100)
    v = torch.zeros(100)
    for step in range(1, 4):
        g = torch.randn(100)
        ref.grad = g.clone()
        opt.step()
        p, m, v = orc.adam_step(p, g, m, v, step, 5e-6)
    assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-9)


def _eval_cases(golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, "eval_kat.npz"))
    return [(torch.from_numpy(z[f"pred{i}"]), torch.from_numpy(z[f"gt{i}"]), torch.from_numpy(z[f"re{i}"]), z[f"metrics{i}"])
            for i in z["cases"]]


def test_eval_oracle_matches_reference_golden(golden_dir):
    """oracle/eval_oracle.py against re_assign_label + evaluate_nau_segmentation_v2 of the unmodified reference
    (tests/golden/make_eval_golden.py)."""
    from oracle import eval_oracle as ev
    for pred, gt, re, metrics in _eval_cases(golden_dir):
        got_re = ev.re_assign_label(pred, gt)
        assert torch.equal(got_re, re)
        assert np.allclose(np.array(ev.evaluate(got_re, gt)), metrics, rtol=1e-6, atol=1e-7)


def test_metrics_from_confusion_counts_match_reference_golden(golden_dir):
    """Host arithmetic of onet_b200.evaluate on the four confusion counts (the counts themselves come from the CUDA kernel,
    tested with -m gpu; here they are formed with torch on CPU) against the reference's metrics."""
    from onet_b200.evaluate import segmentation_metrics
    for pred, gt, re, metrics in _eval_cases(golden_dir):
        counts = [int(((pred == p) & (gt == g)).sum()) for p in (0, 1) for g in (0, 1)]
        m = segmentation_metrics(counts, reassign=True)
        assert m["flipped"] == (not torch.equal(re, pred))
        assert np.allclose([m["acc"], m["miou"], m["dr"], m["far"], m["t_iou"]], metrics, rtol=1e-6, atol=1e-7)


def _cascade_states(z):
    """(st1, st2, loader) of tests/golden/cascade.npz: seeded weights + the reference's running buffers + its batches."""
    from oracle import onet_oracle as orc
    b, h, w, s1, s2 = (int(v) for v in z["meta"])
    states = []
    for tag, seed in (("net1", s1), ("net2", s2)):
        st = orc.perturb_bn_affine(orc.init_state(1, seed=seed), seed=seed + 100)
        for k in list(st):
            if "running" in k or "num_batches" in k:
                st[k] = torch.from_numpy(z[f"{tag}.{k}"]).clone()
        states.append(st)
    loader = [(torch.from_numpy(z[f"x{i}"]), torch.from_numpy(z[f"label{i}"]).long(), torch.zeros(b)) for i in range(2)]
    return states[0], states[1], loader


def test_cascade_oracle_matches_reference_golden(golden_dir):
    """oracle/eval_oracle.py's restatement of test_simclutter / test_2nd_stage_simclutter against the tuples the unmodified
    reference functions returned (tests/golden/make_cascade_golden.py), and tensor_normal_per_frame bit for bit."""
    import os
    from oracle import eval_oracle as ev
    from oracle import onet_oracle as orc
    z = np.load(os.path.join(golden_dir, "cascade.npz"))
    assert np.array_equal(orc.tensor_normal_per_frame(torch.from_numpy(z["norm_in"])).numpy(), z["norm_out"])
    st1, st2, loader = _cascade_states(z)
    one = ev.test_simclutter(st1, loader)
    assert np.allclose(one, z["one_stage"], rtol=0, atol=1e-7), (one, z["one_stage"])
    s1, s2 = ev.test_2nd_stage_simclutter(st1, st2, loader)
    got = (s2[0], s2[1], s2[2], s2[3], s1[4])
    assert np.allclose(got, z["two_stage"], rtol=0, atol=1e-7), (got, z["two_stage"])


def test_synth_oracle_matches_reference_golden(golden_dir):
    """oracle/synth_oracle.py (gaussian_kernel2d, add_gaussian_template_on_clutter_v3 swerling 0, the target loop of
    get_rayleigh_frame) against outputs of the unmodified reference generator (tests/golden/make_synth_golden.py), and the
    host-side target table of onet_b200.synth against the same formulas."""
    import os
    from oracle import synth_oracle as so
    from onet_b200 import synth
    z = np.load(os.path.join(golden_dir, "synth.npz"))
    assert np.allclose(so.gaussian_kernel2d(2.25, 4.25, 37.0), z["kg"], rtol=1e-14, atol=0)
    for f in range(3):
        cx, cy, w, h, theta = z[f"par{f}"]
        out, mask, erc = so.composite_frame(z[f"bg{f}"], cx, cy, w, h, theta, int(z[f"snr{f}"]))
        assert np.allclose(out, z[f"out{f}"], rtol=1e-13, atol=0)
        assert np.array_equal(mask.astype(np.uint8), z[f"mask{f}"])
        tab = synth.target_table(cx[None], cy[None], w[None], h[None], theta[None], 128, 128, host_threshold=True)[0]
        for i in range(len(cx)):
            kg = so.gaussian_kernel2d((w[i] / 2 - 0.5) / 2, (h[i] / 2 - 0.5) / 2, theta[i])
            assert kg.shape == (2 * tab["hr"][i] + 1, 2 * tab["wr"][i] + 1)
            assert abs(tab["thr"][i] - (kg.max() - 2 * kg.std())) < 1e-6
    with pytest.raises(ValueError):
        so.composite_frame(z["bg0"], [3.0], [64.0], [10.0], [18.0], [0.0], 4)          # window leaves the frame
    with pytest.raises(ValueError):
        synth.target_table([[3.0]], [[64.0]], [[10.0]], [[18.0]], [[0.0]], 128, 128)
    with pytest.raises(ValueError):
        so.composite_frame(z["bg0"], [64.0], [64.0], [10.0], [18.0], [0.0], 13)         # snr outside the reference's table


def test_kclutter_oracle_matches_reference_golden(golden_dir):
    """oracle/kclutter_oracle.py against the unmodified reference's generate_K_distributed_noise on the reference's own
    replayed draws (tests/golden/make_kclutter_golden.py), its mnlt() known answers, and the closed-form restatement of
    np.roots(quadratic)[0] against np.roots itself."""
    import os
    from oracle import kclutter_oracle as ko
    z = np.load(os.path.join(golden_dir, "kclutter.npz"))
    for i in range(3):
        size, v, _ = (int(t) for t in z[f"meta_{i}"])
        amp, tex = ko.k_field(z[f"w1_{i}"], z[f"w2_{i}"], v)
        assert np.allclose(amp, z[f"amp_{i}"], rtol=1e-11, atol=1e-12) and np.allclose(tex, z[f"tex_{i}"], rtol=1e-11, atol=1e-12)
    for v in (1, 3, 5, 8):
        assert np.allclose(ko.mnlt(z["mnlt_x"], v), z[f"mnlt_v{v}"], rtol=1e-13, atol=0)
    rs = np.random.RandomState(0)
    for _ in range(2000):
        a, b, c = rs.uniform(0.01, 0.2), rs.uniform(0.01, 0.2) * rs.choice([-1, 1]), rs.uniform(-0.05, 0.05)
        want = np.roots([a, b, c])[0]
        got = ko.first_root(a, b, np.array([c]))[0]
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (a, b, c, got, want)


def test_numpy_restatement_anchors_the_aten_oracle():
    """oracle/onet_numpy_oracle.py (plain numpy, float64, no torch ops) against the ATen-based oracle on a 1 x 32 x 48 batch of
    two frames (ragged: exercises nothing divisible-by-16 specific) and a 40 x 24 one that goes through the F.pad branch:
    forward maps and loss agree to fp32 rounding, and the ATen oracle's autograd gradient matches central finite differences
    of the numpy loss on randomly chosen weights of every kind (first conv, deep conv, BatchNorm scale / shift, up-conv weight
    and bias)."""
    from oracle import onet_numpy_oracle as onp
    for (b, h, w, seed) in ((2, 32, 48, 21), (1, 40, 24, 22)):
        st = orc.perturb_bn_affine(orc.init_state(1, seed=seed), seed=seed + 1)
        x = orc.rayleigh_frames(b, 1, h, w, seed=seed)
        st64 = {k: v.double().numpy() for k, v in st.items() if v.dtype.is_floating_point}
        ref = onp.onet_forward_loss(st64, x.numpy())
        stc = {k: v.clone() for k, v in st.items()}
        leaves = [k for k, v in stc.items() if v.dtype.is_floating_point and "running" not in k]
        for k in leaves:
            stc[k].requires_grad_(True)
        Lt, Vt, Ld, Vd, S = orc.onet_forward(stc, x, training=True)
        loss = orc.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
        rel = lambda a, bb: float(np.linalg.norm(a - bb) / np.linalg.norm(bb))
        assert abs(float(loss) - ref["loss"]) <= 2e-5 * abs(ref["loss"]), (float(loss), ref["loss"])
        assert rel(Lt.detach().numpy(), ref["Lt"]) < 1e-5 and rel(Vt.detach().numpy(), ref["Vt"]) < 1e-4
        assert rel(S[:, 0:1].detach().numpy(), ref["St"]) < 1e-3
        if seed != 21:
            continue
        loss.backward()
        rs = np.random.RandomState(3)
        picks = ["inc.double_conv.0.weight", "down4.maxpool_conv.1.double_conv.3.weight", "up1.up.weight", "up4.up.bias",
                 "down2.maxpool_conv.1.double_conv.1.weight", "up3.conv.double_conv.4.bias", "up4.conv.double_conv.3.weight"]
        for k in picks:
            g = stc[k].grad.numpy().reshape(-1)
            idx = int(np.argmax(np.abs(g))) if rs.rand() < 0.5 else int(rs.randint(g.size))
            eps = 1e-4 * max(1.0, float(np.abs(st64[k]).max()))
            fd = []
            for sgn in (+1, -1):
                pert = dict(st64)
                arr = st64[k].copy().reshape(-1)
                arr[idx] += sgn * eps
                pert[k] = arr.reshape(st64[k].shape)
                fd.append(onp.onet_forward_loss(pert, x.numpy())["loss"])
            num = (fd[0] - fd[1]) / (2 * eps)
            tol = 2e-2 * max(abs(num), float(np.abs(g).max()) * 1e-2)
            assert abs(num - g[idx]) <= tol, (k, idx, num, g[idx])


def test_reduced_precision_restatements_and_teacher_forcing():
    """oracle.onet_forward(emulate=..., force=...) — the checker of the bf16 / tf32 CUDA modes (tests/test_parity_gpu.py):
    rounding happens where the policy says, the tf32 model is truncation, teacher forcing with a run's own taps reproduces that run
    exactly, and forcing another run's conv outputs makes the forward decisions (hence the activations) that run's."""
    import torch
    from oracle import onet_oracle as orc
    st = orc.perturb_bn_affine(orc.init_state(1, seed=3), seed=103)
    x = orc.rayleigh_frames(2, 1, 32, 32, seed=3)
    bf = lambda t: t.bfloat16().float()
    t = torch.tensor([1.0 + 2.0 ** -11, -3.1415927, 1e-30, 65504.0])
    tt = orc._trunc_tf32(t)
    assert torch.equal(tt.view(torch.int32) & 0x1FFF, torch.zeros(4, dtype=torch.int32)) and float(tt[0]) == 1.0 and abs(float(tt[1])) < abs(float(t[1]))
    outs, grads, taps = {}, {}, {}
    for mode in (None, "tf32", "bf16"):
        taps[mode] = {}
        outs[mode], grads[mode], _ = orc.train_step_outputs(st, x, emulate=mode, taps=taps[mode])
    rel = lambda a, b: float((a - b).norm() / b.norm())
    # precision ladder against the FP32 run
    assert rel(outs["tf32"]["Vt"], outs[None]["Vt"]) < 5e-3 < 5 * rel(outs["bf16"]["Vt"], outs[None]["Vt"])
    # bf16 policy: stored conv outputs and activations are bf16 values - except the first conv of a 1-channel network, whose
    # output is never stored (closed-form statistics / recompute: onet_b200/csrc/first_layer.cuh) and therefore never rounded;
    # tf32 policy: storage stays fp32
    for key, v in taps["bf16"]["top"].items():
        if key == "inc.double_conv.0.raw":
            assert not torch.equal(bf(v.detach()), v.detach())
        elif key.endswith(".raw") or key.endswith(".act") or key.endswith(".up.out"):
            assert torch.equal(bf(v.detach()), v.detach()), key
    raw = taps["tf32"]["top"]["down1.maxpool_conv.1.double_conv.0.raw"].detach()
    assert not torch.equal(bf(raw), raw)
    # teacher forcing with the run's own stored outputs: the identical run
    for mode in ("bf16", "tf32"):
        force = {b: {k: v.detach() for k, v in taps[mode][b].items() if k.endswith(".raw") or k.endswith(".up.out")} for b in ("top", "dwn")}
        assert len(force["top"]) == 22
        o2, g2, _ = orc.train_step_outputs(st, x, emulate=mode, force=force)
        assert float(o2["loss"]) == float(outs[mode]["loss"])
        assert max(rel(g2[k], grads[mode][k]) for k in g2) == 0.0
    # forcing the bf16 run's conv outputs into the FP32 oracle: its activations become the bf16 run's (up to the activation
    # rounding the FP32 oracle does not do), and `.own` keeps what the FP32 oracle computed from the forced inputs
    force = {b: {k: v.detach() for k, v in taps["bf16"][b].items() if k.endswith(".raw") or k.endswith(".up.out")} for b in ("top", "dwn")}
    t3 = {}
    o3, _, _ = orc.train_step_outputs(st, x, emulate=None, force=force, taps=t3)
    assert rel(o3["Vt"], outs["bf16"]["Vt"]) < 0.2 * rel(outs[None]["Vt"], outs["bf16"]["Vt"])
    own = t3["top"]["inc.double_conv.3.raw.own"]
    assert 0 < rel(own, force["top"]["inc.double_conv.3.raw"]) < 1e-2
