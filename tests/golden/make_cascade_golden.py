"""Golden vectors of the evaluation loops from the UNMODIFIED reference (container only; needs /root/reference).

    python tests/golden/make_cascade_golden.py

Imports the reference training script `Train_Onet_on_simclutter_20250407.py` (its top-level imports of matplotlib /
skimage / albumentations are stubbed by oracle/ref_import.py; nothing of it is copied) and calls its own
`test_simclutter` (:97-172) and `test_2nd_stage_simclutter` (:296-390) with verbose=0 on two seeded batches and two
seeded reference Onets whose BatchNorm running statistics come from three training-mode forwards of the reference.
Stored: inputs, labels, the running buffers of both networks, and the tuples the two functions return.

Labels are derived from the reference's own stage-1 prediction (10 % of the pixels flipped; batch 1 inverted) so that
`re_assign_label`'s keep / flip decision is far from a tie and both branches of :328-331 are exercised."""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import onet_oracle as orc  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B, H, W = 2, 48, 48
SEEDS = (21, 22)


def build(ref, seed):
    st = orc.perturb_bn_affine(orc.init_state(1, seed=seed), seed=seed + 100)
    onet = ref.Onet(in_chns=1, binit=True, bshare=True)
    sd = OrderedDict()
    for k, v in st.items():
        sd["topu." + k] = v.clone()
        sd["dwnu." + k] = v.clone()
    onet.load_state_dict(sd)
    return onet


def main():
    ref = import_reference()
    argv, sys.argv = sys.argv, ["x"]
    import Train_Onet_on_simclutter_20250407 as tr
    import utils_20231218 as uti
    sys.argv = argv
    torch.set_num_threads(8)
    onet, onet2 = build(ref, SEEDS[0]), build(ref, SEEDS[1])
    xs = [orc.rayleigh_frames(B, 1, H, W, seed=300 + i) for i in range(2)]
    # running statistics: three training-mode forwards each (stage 2 sees normalised stage-1 maps)
    with torch.no_grad():
        onet.train()
        for i in range(3):
            _, Vt, _, Vd, _ = onet(xs[i % 2])
        onet2.train()
        for i in range(3):
            onet2(uti.tensor_normal_per_frame(Vd))
        onet.eval()
        onet2.eval()
        labels = []
        for i, x in enumerate(xs):
            _, _, _, _, S = onet(x)
            raw = onet.predict_label(S)
            gen = torch.Generator().manual_seed(500 + i)
            noise = torch.rand(raw.shape, generator=gen) < 0.1
            lab = torch.where(noise, 1 - raw, raw)
            labels.append(1 - lab if i == 1 else lab)
    loader = [(x, lab, torch.zeros(B)) for x, lab in zip(xs, labels)]
    cfg = types.SimpleNamespace(device="cpu", out_root="/tmp")
    one = tr.test_simclutter("golden", cfg, onet, loader, verbose=0)
    two = tr.test_2nd_stage_simclutter("golden", cfg, onet, onet2, loader, verbose=0)
    store = {"meta": np.array([B, H, W, SEEDS[0], SEEDS[1]]), "one_stage": np.array([float(v) for v in one]),
             "two_stage": np.array([float(v) for v in two])}
    for i, (x, lab) in enumerate(zip(xs, labels)):
        store[f"x{i}"] = x.numpy()
        store[f"label{i}"] = lab.numpy().astype(np.uint8)
    for tag, net in (("net1", onet), ("net2", onet2)):
        for k, v in net.state_dict().items():
            if k.startswith("topu.") and ("running" in k or "num_batches" in k):
                store[f"{tag}.{k[len('topu.'):]}"] = v.numpy().copy()
    # tensor_normal_per_frame known answers (utils_20231218.py:673-689), including a constant frame
    gen = torch.Generator().manual_seed(9)
    t = torch.randn(3, 2, 17, 23, generator=gen) * 5
    t[1, 0] = 2.5
    store["norm_in"] = t.numpy()
    store["norm_out"] = uti.tensor_normal_per_frame(t).numpy()
    np.savez_compressed(os.path.join(OUT, "cascade.npz"), **store)
    print("one stage", one)
    print("two stage", two)
    print("KB", os.path.getsize(os.path.join(OUT, "cascade.npz")) // 1024)


if __name__ == "__main__":
    main()
