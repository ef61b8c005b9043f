"""Golden vectors of the evaluation step from the UNMODIFIED reference (container only; needs /root/reference).

    python tests/golden/make_eval_golden.py

Imports `utils_20231218.py` (matplotlib stubbed: it is imported at module top but unused by these functions), runs
`re_assign_label` + `evaluate_nau_segmentation_v2` on seeded label pairs and stores labels + the five metrics."""
import os
import sys
import types

import numpy as np
import torch

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference/source_code")
sys.argv = ["x"]
import utils_20231218 as uti  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    gen = torch.Generator().manual_seed(77)
    store = {}
    cases = []
    for i, (p_fg, flip, agree) in enumerate(((0.1, False, 0.9), (0.1, True, 0.9), (0.5, False, 0.5), (0.02, True, 0.99), (0.0, False, 1.0))):
        gt = (torch.rand(2, 24, 40, generator=gen) < p_fg).long()
        noise = torch.rand(2, 24, 40, generator=gen) > agree
        pred = torch.where(noise, 1 - gt, gt)
        if flip:
            pred = 1 - pred
        re = uti.re_assign_label(pred, gt)
        m = uti.evaluate_nau_segmentation_v2(re, gt)
        store[f"pred{i}"] = pred.numpy().astype(np.int64)
        store[f"gt{i}"] = gt.numpy().astype(np.int64)
        store[f"re{i}"] = re.numpy().astype(np.int64)
        store[f"metrics{i}"] = np.array([float(v) for v in m], dtype=np.float64)
        cases.append(i)
    store["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "eval_kat.npz"), **store)
    print("wrote eval_kat.npz", {k: v.shape for k, v in store.items() if k.startswith("metrics")})


if __name__ == "__main__":
    main()
