"""Golden vectors of the frame synthesis from the UNMODIFIED reference generator (container only; needs /root/reference).

    python tests/golden/make_synth_golden.py

Imports Rayleigh_bg_Gaussian_EOT_generator_20230208.py (matplotlib / skimage / albumentations stubbed as in
oracle/ref_import.py) and (a) calls its gaussian_kernel2d / add_gaussian_template_on_clutter_v3 on seeded backgrounds and
target parameters, (b) calls its get_rayleigh_frame under a fixed numpy seed while recording the draws it makes, so that a
test can replay the same background and target parameters.  Stored at 128 x 128 (targets drawn inside) to stay small."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import_reference()
    argv, sys.argv = sys.argv, ["x"]
    import Rayleigh_bg_Gaussian_EOT_generator_20230208 as gen
    sys.argv = argv
    store = {}
    # (a) direct calls: 3 frames of 128 x 128, 6 targets each, snr 0 / 4 / 10
    rs = np.random.RandomState(77)
    for f, snr in enumerate((0, 4, 10)):
        bg = rs.rayleigh(1.0, size=(128, 128))
        T = 6
        cx, cy = rs.normal(64, 18, T), rs.normal(64, 14, T)
        w, h, theta = rs.normal(10, 2, T), rs.normal(18, 2, T), rs.rand(T) * 180
        erc = np.sum(bg ** 2) / bg.size
        out, mask = bg.copy(), np.zeros_like(bg)
        for i in range(T):
            out, mask = gen.add_gaussian_template_on_clutter_v3(cx[i], cy[i], w[i], h[i], theta[i], erc, snr, out, mask, 0)
        store[f"bg{f}"] = bg.astype(np.float32)          # the device works in fp32: the fixture holds the fp32 background
        store[f"par{f}"] = np.stack([cx, cy, w, h, theta])
        store[f"snr{f}"] = np.int64(snr)
        # re-run on the fp32-rounded background so that fixture input and fixture output belong together
        bg32 = store[f"bg{f}"].astype(np.float64)
        erc = np.sum(bg32 ** 2) / bg32.size
        out, mask = bg32.copy(), np.zeros_like(bg32)
        for i in range(T):
            out, mask = gen.add_gaussian_template_on_clutter_v3(cx[i], cy[i], w[i], h[i], theta[i], erc, snr, out, mask, 0)
        store[f"out{f}"] = out
        store[f"mask{f}"] = mask.astype(np.uint8)
    store["kg"] = gen.gaussian_kernel2d(2.25, 4.25, 37.0, bnorm=False)
    # (b) get_rayleigh_frame under a fixed seed: first two moments and mask size (distribution-level anchors)
    np.random.seed(1981)
    frame, fmask = gen.get_rayleigh_frame(snr=6)
    store["ray_frame_stats"] = np.array([frame.mean(), (frame ** 2).mean(), frame.max(), float(fmask.sum())])
    np.savez_compressed(os.path.join(OUT, "synth.npz"), **store)
    print({k: getattr(v, "shape", v) for k, v in store.items()})
    print("KB", os.path.getsize(os.path.join(OUT, "synth.npz")) // 1024)


if __name__ == "__main__":
    main()
