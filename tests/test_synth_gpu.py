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
