"""Training-trajectory parity (-m gpu): OnetTrainer (zero_grad -> forward -> JSD loss -> backward -> fused Adam, every step
through the CUDA kernels) — and the reference's literal loop with torch.optim.Adam on this module's parameters — against the CPU oracle driven by torch.optim.Adam — the reference's loop
(Train_Onet_on_simclutter_20250407.py:181, 209-218) — from identical weights on identical batches, over 12 optimizer steps.

What is compared is the LOSS SEQUENCE: each step's loss depends on all previous updates, so a wrong gradient, a wrong Adam
bias correction or a stale packed weight shows up within a few steps.  Individual weights are not compared: Adam's first
steps are ~lr * sign(g) and elements whose gradient is ~0 flip under fp32 summation-order noise (see test_model_gpu.py)."""
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,graph,tol", [("fp32", False, 2e-4), ("bf16", False, 1e-2), ("bf16", True, 1e-2),
                                            ("fp32", "torch.optim", 2e-4), ("bf16", "torch.optim", 1e-2)])
def test_loss_sequence_matches_oracle_training(mode, graph, tol):
    import onet_b200
    from onet_b200.trainer import OnetTrainer
    from oracle import onet_oracle as orc
    steps, B, H, W, lr = 12, 4, 64, 64, 2e-5
    xs = [orc.rayleigh_frames(B, 1, H, W, seed=900 + i) for i in range(3)]
    st = orc.perturb_bn_affine(orc.init_state(1, seed=77), seed=177)
    # oracle loop on CPU
    torch.set_num_threads(8)
    ost = {k: v.clone() for k, v in st.items()}
    leaves = [k for k, v in ost.items() if v.dtype.is_floating_point and "running" not in k]
    params = [ost[k].requires_grad_(True) for k in leaves]
    opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=1e-8)
    ref = []
    for i in range(steps):
        opt.zero_grad()
        Lt, Vt, Ld, Vd, S = orc.onet_forward(ost, xs[i % 3], training=True)
        loss = orc.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
        loss.backward()
        opt.step()
        ref.append(float(loss))
    # this repo
    net = onet_b200.Onet(1, True, True, mode=mode)
    sd = OrderedDict()
    for k, v in st.items():
        sd["topu." + k] = v.clone()
        sd["dwnu." + k] = v.clone()
    net.load_state_dict(sd)
    net = net.cuda()
    if graph == "torch.optim":
        # the reference's own loop with the one-line import change: torch.optim.Adam on onet.parameters()
        opt2 = torch.optim.Adam(net.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)
        got = []
        for i in range(steps):
            net.train()
            net.zero_grad()
            Lt, Vt, Ld, Vd, S = net(xs[i % 3].cuda())
            St = S[:, 0, :, :].unsqueeze(dim=1)
            Sd = S[:, 1, :, :].unsqueeze(dim=1)
            loss = net.compute_loss(Lt, St, Ld, Sd)
            loss.backward()
            opt2.step()
            got.append(loss.item())
    else:
        tr = OnetTrainer(net, lr=lr, graph=graph)
        got = [float(tr.step(xs[i % 3].cuda())) for i in range(steps)]
    rel = [abs(a - b) / abs(b) for a, b in zip(got, ref)]
    print(f"{mode} graph={graph}: oracle {ref[0]:.5f} -> {ref[-1]:.5f}, here {got[0]:.5f} -> {got[-1]:.5f}, max rel {max(rel):.2e}")
    assert ref[-1] < ref[0]                      # the oracle's loss falls over these steps, so a frozen model would be caught
    assert max(rel) < tol, rel
