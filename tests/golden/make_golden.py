"""Generate golden vectors from the UNMODIFIED reference (container only; needs /root/reference).

    python tests/golden/make_golden.py

Imports the reference `Onet` (oracle/ref_import.py), loads a seeded state produced by
`oracle.onet_oracle.init_state` into it through `load_state_dict`, runs the reference's own
training-step sequence (Train_Onet_on_simclutter_20250407.py:209-217: forward, slice S,
compute_loss, backward) and an eval-mode forward, and stores inputs + results as small .npz files.
Large tensors are stored as (norm, sum, strided sample) so the fixtures stay small.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import onet_oracle as orc  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE_STRIDE = 9973
CASES = [
    # name, in_chns, batch, H, W, bshare, seed
    ("c1_b2_32x32", 1, 2, 32, 32, True, 11),
    ("c3_b1_48x32", 3, 1, 48, 32, True, 12),
    ("c1_b2_32x32_noshare", 1, 2, 32, 32, False, 13),
    ("c1_b2_40x56_pad", 1, 2, 40, 56, True, 14),   # H,W not divisible by 16 -> F.pad branch (:92-96)
]


def summarize(name, t, store):
    a = t.detach().cpu().numpy().astype(np.float32)
    store[f"{name}.norm"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))
    store[f"{name}.sum"] = np.float64(a.astype(np.float64).sum())
    if a.size <= 8192:
        store[f"{name}.full"] = a
    else:
        store[f"{name}.sample"] = a.reshape(-1)[::SAMPLE_STRIDE].copy()


def load_into_reference(onet, st_top, st_dwn=None):
    sd = OrderedDict()
    for k, v in st_top.items():
        sd["topu." + k] = v.clone()
    for k, v in (st_top if st_dwn is None else st_dwn).items():
        sd["dwnu." + k] = v.clone()
    onet.load_state_dict(sd)


def main():
    ref = import_reference()
    torch.set_num_threads(8)
    for name, cin, b, h, w, bshare, seed in CASES:
        st = orc.perturb_bn_affine(orc.init_state(cin, seed=seed), seed=seed + 100)
        st_d = None if bshare else orc.perturb_bn_affine(orc.init_state(cin, seed=seed + 50), seed=seed + 150)
        x = orc.rayleigh_frames(b, cin, h, w, seed=seed)
        onet = ref.Onet(in_chns=cin, binit=True, bshare=bshare)
        load_into_reference(onet, st, st_d)
        onet.train()
        onet.zero_grad()
        Lt, Vt, Ld, Vd, S = onet(x)
        St = S[:, 0, :, :].unsqueeze(dim=1)
        Sd = S[:, 1, :, :].unsqueeze(dim=1)
        loss = onet.compute_loss(Lt, St, Ld, Sd)
        loss.backward()
        store = {"x": x.numpy(), "loss": np.float64(loss.item()),
                 "Vt": Vt.detach().numpy(), "Vd": Vd.detach().numpy(), "S": S.detach().numpy(),
                 "meta": np.array([cin, b, h, w, int(bshare), seed])}
        summarize("Lt", Lt, store)
        summarize("Ld", Ld, store)
        for k, p in onet.named_parameters():      # shared twin: named_parameters() lists topu.* only
            summarize("grad." + k, p.grad, store)
        for k, v in onet.state_dict().items():
            if "running" in k or "num_batches" in k:
                store["buf." + k] = v.detach().numpy().copy()
        # eval-mode forward with the updated running statistics + labels
        onet.eval()
        with torch.no_grad():
            Lt2, Vt2, Ld2, Vd2, S2 = onet(x)
            lab = onet.predict_label(S2)
        store["eval.Vt"] = Vt2.numpy()
        store["eval.Vd"] = Vd2.numpy()
        store["eval.S"] = S2.numpy()
        store["eval.label"] = lab.numpy().astype(np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        print(name, "loss", loss.item(), "files KB", os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024)

    # known-answer vector for the piecewise softplus (Onet_vanilla_20240606.py:237-251)
    onet = ref.Onet(1, False, True)
    xs = torch.tensor([-80.0, -50.0, -37.0001, -37.0, -36.9999, -30.0, -17.0, -16.0, -5.0, -1.0, 0.0, 0.5, 3.0,
                       10.0, 17.9999, 18.0, 18.0001, 25.0, 33.2, 33.3, 33.4, 50.0, 100.0], dtype=torch.float32)
    xin = xs.clone().requires_grad_(True)
    y = onet.log1pexp(xin * 1.0)      # the reference mutates its argument in place -> pass a temporary
    y.sum().backward()
    np.savez_compressed(os.path.join(OUT, "log1pexp_kat.npz"), x=xs.numpy(), y=y.detach().numpy(),
                        dy=xin.grad.numpy())
    print("log1pexp", y.detach().numpy())


if __name__ == "__main__":
    main()
