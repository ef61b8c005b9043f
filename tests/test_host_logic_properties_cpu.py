"""Seeded property tests of host-side logic (no GPU): metrics from confusion counts == the reference-pinned oracle on random
label pairs; the frame-synthesis target table raises exactly when the reference's compositing would; the window geometry it
hands to the kernel is the oracle's."""
import numpy as np
import pytest
import torch


def test_metrics_from_counts_equal_oracle_on_random_labels():
    from onet_b200.evaluate import segmentation_metrics
    from oracle import eval_oracle as ev
    rs = np.random.RandomState(0)
    for trial in range(200):
        shape = (rs.randint(1, 4), rs.randint(1, 20), rs.randint(1, 20))
        p_fg, agree = rs.choice([0.0, 0.02, 0.3, 0.5, 0.9, 1.0]), rs.choice([0.0, 0.3, 0.5, 0.8, 1.0])
        gt = torch.from_numpy((rs.rand(*shape) < p_fg).astype(np.int64))
        pred = torch.where(torch.from_numpy(rs.rand(*shape) < agree), gt, 1 - gt)
        counts = [int(((pred == p) & (gt == g)).sum()) for p in (0, 1) for g in (0, 1)]
        m = segmentation_metrics(counts, reassign=True)
        re = ev.re_assign_label(pred, gt)
        want = ev.evaluate(re, gt)
        assert m["flipped"] == (not torch.equal(re, pred)), trial
        assert np.allclose([m["acc"], m["miou"], m["dr"], m["far"], m["t_iou"]], want, rtol=1e-6, atol=1e-7), (trial, m, want)


def test_target_table_raises_exactly_when_the_reference_compositing_would():
    from onet_b200 import synth
    from oracle import synth_oracle as so
    rs = np.random.RandomState(1)
    H, W = 96, 80
    bg = rs.rayleigh(1.0, size=(H, W))
    n_raise = n_ok = 0
    for trial in range(400):
        cx, cy = rs.uniform(-5, W + 5), rs.uniform(-5, H + 5)
        w, h, theta = rs.normal(10, 2), rs.normal(18, 2), rs.rand() * 180
        try:
            so.composite_frame(bg, [cx], [cy], [w], [h], [theta], 4)
            ref_raises = False
        except ValueError:
            ref_raises = True
        try:
            tab = synth.target_table([[cx]], [[cy]], [[w]], [[h]], [[theta]], H, W)
            ours_raises = False
        except ValueError:
            ours_raises = True
        assert ours_raises == ref_raises, (trial, cx, cy, w, h)
        if not ours_raises:
            n_ok += 1
            kg = so.gaussian_kernel2d((w / 2 - 0.5) / 2, (h / 2 - 0.5) / 2, theta)
            t = tab[0, 0]
            assert (2 * t["hr"] + 1, 2 * t["wr"] + 1) == kg.shape
            assert t["ly"] == int(cy - (kg.shape[0] - 1) / 2) and t["lx"] == int(cx - (kg.shape[1] - 1) / 2)
            assert t["thr"] < 0                                   # the kernel reduces the mask threshold itself
        else:
            n_raise += 1
    assert n_ok > 100 and n_raise > 50


def test_snr_outside_the_reference_table_raises():
    from oracle import synth_oracle as so
    from onet_b200 import synth
    for snr in (13, -3, 2.5):
        with pytest.raises(ValueError):
            synth.SNR_LIST.index(snr)
        with pytest.raises(ValueError):
            so.SNR_LIST.index(snr)
    assert synth.SNR_LIST == so.SNR_LIST


def test_fastdiv_multiply_high_is_exact():
    """onet_b200/csrc/fastdiv.cuh: n / d = umulhi(n, mul) >> shr with mul = ceil(2^(31 + ceil(log2 d)) / d), for every 0 <= n < 2^31.
    The constants are parsed from the header's own formula (restated here) and checked on the divisors the kernels use (tile counts,
    window counts, channel-octet counts, H * W) at the interval ends, around multiples of d, and on random dividends."""
    import os
    import random
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "onet_b200", "csrc", "fastdiv.cuh")).read()
    assert re.search(r"pw = 31u \+ lg", src) and re.search(r"\(1ull << pw\) \+ f\.d - 1ull\) / f\.d", src) and "f.shr = pw - 32u" in src

    def make(d):
        if d == 1:
            return None
        lg = 0
        while (1 << lg) < d:
            lg += 1
        pw = 31 + lg
        mul = ((1 << pw) + d - 1) // d
        assert mul < (1 << 32)
        return mul, pw - 32

    def div(n, d, f):
        return n if f is None else ((n * f[0]) >> 32) >> f[1]
    rng = random.Random(7)
    divisors = set(range(1, 300)) | {2 ** k for k in range(0, 31)} | {2 ** k + 1 for k in range(1, 30)} | {2 ** k - 1 for k in range(2, 31)}
    divisors |= {128 * 128, 256 * 256, 224 * 224, 1022, 1023, 1025, 2048 * 2048, 65535, 65537, 148, 8192 + 37, 2 ** 31 - 1}
    top = 2 ** 31 - 1
    for d in sorted(divisors):
        f = make(d)
        samples = {0, 1, d - 1, d, d + 1, top, top - 1, top - d, (top // d) * d, (top // d) * d - 1}
        samples |= {rng.randrange(0, top + 1) for _ in range(200)}
        samples |= {k * d + e for k in (1, 2, 3, 1000, top // d - 1) for e in (-1, 0, 1)}
        for n in samples:
            if 0 <= n <= top:
                assert div(n, d, f) == n // d, (n, d)
