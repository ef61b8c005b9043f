"""On-device frame synthesis (-m gpu): onet_b200.synth against the golden outputs of the UNMODIFIED
reference generator (same fp32 backgrounds, same target parameters), against the CPU oracle on more frames, and the
random backgrounds against the distributions they claim."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check_frame(got, got_mask, want, want_mask, tag):
    rel = np.abs(got - want).max() / np.abs(want).max()
    # fp32 on the device vs float64 in the reference: values agree to rounding; a pixel may switch sides of
    # `template > background` or `kgauss > threshold` only when the two are equal to fp32 rounding
    bad = np.abs(got - want) > 1e-4 * np.abs(want).max()
    assert bad.sum() <= 2, (tag, int(bad.sum()), rel)
    assert (got_mask != want_mask).sum() <= 2, (tag, int((got_mask != want_mask).sum()))


def test_target_compositing_matches_reference_golden(golden_dir):
    from onet_b200 import synth
    z = np.load(os.path.join(golden_dir, "synth.npz"))
    for f in range(3):
        par = z[f"par{f}"]
        frames = torch.from_numpy(z[f"bg{f}"]).cuda()[None].contiguous()
        out, mask, erc = synth.add_gaussian_targets(frames, *(p[None] for p in par), int(z[f"snr{f}"]))
        bg64 = z[f"bg{f}"].astype(np.float64)
        assert abs(float(erc[0]) - (bg64 ** 2).mean()) < 1e-5
        _check_frame(out[0].cpu().numpy().astype(np.float64), mask[0].cpu().numpy().astype(np.uint8), z[f"out{f}"], z[f"mask{f}"], f)


def test_target_compositing_matches_oracle_batched():
    """8 frames of 400 x 400 with the reference's 20 targets each (overlapping windows, sequential order matters)."""
    from onet_b200 import synth
    from oracle import synth_oracle as so
    rs = np.random.RandomState(5)
    n = 8
    bg = rs.rayleigh(1.0, size=(n, 400, 400)).astype(np.float32)
    cx, cy, w, h, theta = synth.draw_targets(n, (400, 400), 20, rs)
    for snr in (0, 7):
        out, mask, _ = synth.add_gaussian_targets(torch.from_numpy(bg).cuda(), cx, cy, w, h, theta, snr)
        out, mask = out.cpu().numpy().astype(np.float64), mask.cpu().numpy().astype(np.uint8)
        for f in range(n):
            want, want_mask, _ = so.composite_frame(bg[f], cx[f], cy[f], w[f], h[f], theta[f], snr)
            _check_frame(out[f], mask[f], want, want_mask.astype(np.uint8), (snr, f))
    with pytest.raises(ValueError):
        synth.add_gaussian_targets(torch.from_numpy(bg).cuda(), cx, cy, w, h, theta, 13)
    bad = cx.copy()
    bad[0, 0] = 2.0
    with pytest.raises(ValueError):
        synth.add_gaussian_targets(torch.from_numpy(bg).cuda(), bad, cy, w, h, theta, 4)


def test_background_distributions(golden_dir):
    """Rayleigh(1): mean sqrt(pi/2), E[x^2] = 2, CDF 1 - exp(-x^2/2).  K clutter (nu = 5): E[x^2] = 2 E[texture] = 2,
    E[x^4] = 8 (1 + 1/nu), and the same quantiles as the CPU stand-in of onet_b200.data.  Same seed -> same frames; different
    stream ids or seeds -> different frames; the reference frame's moments (golden) sit inside the same bands."""
    from onet_b200 import synth
    x = synth.rayleigh_background(4, 400, 400, seed=11)
    assert torch.equal(x, synth.rayleigh_background(4, 400, 400, seed=11))
    assert not torch.equal(x, synth.rayleigh_background(4, 400, 400, seed=12))
    assert not torch.equal(x, synth.rayleigh_background(4, 400, 400, seed=11, stream_id=1))
    assert torch.equal(x.flatten()[:1001], synth.rayleigh_background(1, 1, 1001, seed=11).flatten())   # geometry independent
    v = x.double().flatten()
    assert abs(float(v.mean()) - np.sqrt(np.pi / 2)) < 5e-3 and abs(float((v ** 2).mean()) - 2.0) < 1e-2
    for q in (0.5, 1.0, 2.0, 3.0):
        assert abs(float((v <= q).double().mean()) - (1 - np.exp(-q * q / 2))) < 3e-3
    assert float(v.min()) > 0 and torch.isfinite(v).all()
    z = np.load(os.path.join(golden_dir, "synth.npz"))
    ref_mean, ref_e2 = z["ray_frame_stats"][:2]          # reference frame WITH targets: slightly above the pure background
    assert 0 <= ref_mean - float(v.mean()) < 0.05 and 0 <= ref_e2 - float((v ** 2).mean()) < 0.2
    k = synth.k_background(4, 400, 400, seed=11, nu=5).double().flatten()
    assert abs(float((k ** 2).mean()) - 2.0) < 2e-2 and abs(float((k ** 4).mean()) - 8 * 1.2) < 0.3
    from onet_b200.data import k_clutter_frames      # CPU stand-in draws the same compound distribution (before normalisation)
    gen = torch.Generator().manual_seed(3)
    u = torch.rand(640000, generator=gen).clamp_min(1e-12)
    e = -torch.log(torch.rand(5, 640000, generator=gen).clamp_min(1e-12))
    cpu = (torch.sqrt(-2 * torch.log(u)) * torch.sqrt(e.sum(0) / 5)).double()
    for p in (0.1, 0.5, 0.9, 0.99):
        assert abs(float(torch.quantile(k[:640000], p)) - float(torch.quantile(cpu, p))) < 2e-2 * (1 + float(torch.quantile(cpu, p)))


def test_dataset_and_checkpoint_formats(tmp_path):
    """prepare_data writes the dictionary the reference dataloader reads (simbg4onet_20230209.py:298-305: keys, shapes,
    dtypes, per-frame [0,1] range); save_checkpoint / load_checkpoint use the reference's {'net', 'epoch'} file (:264-266)."""
    import onet_b200
    from onet_b200 import synth
    fn = str(tmp_path / "rayleigh.pt")
    synth.prepare_data(img_sz=(224, 224), bg_type="rayleigh", file_name=fn, fnums=3, snrs=(0, 5, 10), seed=3)
    data = torch.load(fn, map_location=lambda storage, loc: storage)
    imgs, labels, snrs = data["rayleigh_imgs"], data["rayleigh_labels"], torch.tensor(data["psnr"])
    assert imgs.shape == (9, 1, 224, 224) and imgs.dtype == torch.float32
    assert labels.shape == (9, 224, 224) and labels.dtype == torch.float32 and set(labels.unique().tolist()) <= {0.0, 1.0}
    assert snrs.tolist() == [0, 0, 0, 5, 5, 5, 10, 10, 10] and isinstance(data["desc"], str)
    assert float(imgs.min()) >= 0 and float(imgs.max()) <= 1 and 0 < float(labels.mean()) < 0.2
    # labelled pixels are brighter than the clutter around them
    assert float(imgs[:, 0][labels > 0].mean()) > float(imgs[:, 0][labels == 0].mean())
    net = onet_b200.Onet(1, True, True).cuda()
    ck = str(tmp_path / "ck.pytorch")
    synth.save_checkpoint(net, 17, ck)
    raw = torch.load(ck, map_location=lambda storage, loc: storage)
    assert set(raw) == {"net", "epoch"} and raw["epoch"] == 17
    assert any(k.startswith("topu.inc.double_conv.0.") for k in raw["net"]) and any(k.startswith("dwnu.up4.") for k in raw["net"])
    net2 = onet_b200.Onet(1, False, True).cuda()
    assert synth.load_checkpoint(net2, ck) == 17
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k


def test_correlated_k_field_matches_reference_golden(golden_dir):
    """onet_b200.synth.k_correlated_background on the reference's own replayed white-noise draws against the outputs of the
    UNMODIFIED generate_K_distributed_noise (amplitude and Gamma texture; float64 pipeline, fp32 amplitude), the mnlt kernel
    against the reference's known answers, and the device white noise against N(0,1)."""
    from onet_b200 import synth
    z = np.load(os.path.join(golden_dir, "kclutter.npz"))
    for i in range(3):
        size, v, _ = (int(t) for t in z[f"meta_{i}"])
        w1, w2 = torch.from_numpy(z[f"w1_{i}"])[None], torch.from_numpy(z[f"w2_{i}"])[None]
        amp, tex = synth.k_correlated_background(1, size, v=v, white=(w1, w2), return_texture=True)
        amp, tex = amp[0].double().cpu().numpy(), tex[0].cpu().numpy()
        assert np.abs(tex - z[f"tex_{i}"]).max() <= 1e-8 * np.abs(z[f"tex_{i}"]).max(), i
        assert np.abs(amp - z[f"amp_{i}"]).max() <= 2e-7 * np.abs(z[f"amp_{i}"]).max(), i       # fp32 output
    x = torch.from_numpy(z["mnlt_x"]).cuda()
    for v in (1, 3, 5, 8):
        got = synth.mnlt(x, v).cpu().numpy()
        # the reference forms Phi(x) as 1 - erfc(x / sqrt 2) / 2 and loses digits in the upper tail; the kernel does not
        assert np.abs(got / z[f"mnlt_v{v}"] - 1).max() < 1e-7, v
    # batched: two frames at once == frame by frame
    w = torch.randn(2, 2, 64, 64, dtype=torch.float64)
    both = synth.k_correlated_background(2, 64, white=(w[0], w[1]))
    one = synth.k_correlated_background(1, 64, white=(w[0, :1], w[1, :1]))
    assert torch.allclose(both[:1], one, rtol=1e-6, atol=0)
    g = synth.normal_white(4, 400, 400, seed=5).flatten()
    assert torch.equal(g, synth.normal_white(4, 400, 400, seed=5).flatten())
    assert abs(float(g.mean())) < 5e-3 and abs(float(g.var()) - 1.0) < 1e-2 and abs(float((g ** 4).mean()) - 3.0) < 5e-2
    f, m = synth.get_k_frames(3, snr=6, seed=9)                     # the reference's setup: 400 x 400, 20 targets
    assert f.shape == (3, 400, 400) and m.shape == (3, 400, 400) and bool(m.any()) and torch.isfinite(f).all()
    e2 = float((synth.k_correlated_background(3, 400, seed=9) ** 2).mean())
    assert 0.6 < e2 < 0.85, e2         # the unmodified reference at 400 x 400 (seed 9, 5.8 s on the CPU): 0.727
