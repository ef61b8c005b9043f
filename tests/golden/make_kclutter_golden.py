"""Golden vectors of the correlated K-clutter field from the UNMODIFIED reference (container only; needs /root/reference).

    python tests/golden/make_kclutter_golden.py

Seeds numpy, calls the reference's generate_K_distributed_noise(height, width, gamma_shape) and stores its two outputs
together with the two white-noise fields it drew (replayed from the same seed: np.random.normal(size=(h, w)) at :483, then
np.random.normal(size=(M, M)) at :287), plus known answers of its mnlt()."""
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
    import K_distributed_SeaClutter_Simulation_20210919 as km
    sys.argv = argv
    store = {}
    for i, (size, v, seed) in enumerate(((64, 5, 3), (96, 5, 4), (48, 3, 5))):
        np.random.seed(seed)
        amp, tex = km.generate_K_distributed_noise(size, size, gamma_shape=v)
        np.random.seed(seed)
        w1 = np.random.normal(loc=0, scale=1, size=(size, size))
        w2 = np.random.normal(loc=0, scale=1, size=(size, size))
        store[f"w1_{i}"], store[f"w2_{i}"], store[f"amp_{i}"], store[f"tex_{i}"] = w1, w2, amp, tex
        store[f"meta_{i}"] = np.array([size, v, seed])
    x = np.array([-6.0, -4.0, -2.5, -1.0, -0.1, 0.0, 0.3, 1.0, 2.0, 3.5, 5.0, 6.5])
    store["mnlt_x"] = x
    for v in (1, 3, 5, 8):
        store[f"mnlt_v{v}"] = km.mnlt(x, v)
    np.savez_compressed(os.path.join(OUT, "kclutter.npz"), **store)
    print({k: getattr(v, "shape", v) for k, v in store.items()}, os.path.getsize(os.path.join(OUT, "kclutter.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
