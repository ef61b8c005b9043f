"""Size-independent properties at the FULL benchmark size (-m gpu): 128 twin images of 256 x 256 (BASELINE.json configs[1],
batch 64), where an element-by-element oracle is too slow.

  * scaling:   conv(x; 2 W) == 2 conv(x; W) bit for bit (power-of-two scaling commutes with every rounding of the pipeline),
               for the weight-resident 64-channel kernel and the CTA-pair kernel at their real shapes
  * adjoints:  <conv(x; W), g> == <x, dgrad(g; W)> == <W, wgrad(g, x)>  — forward, data-gradient and weight-gradient kernels are
               three views of one bilinear form (dot-product test; bf16 outputs are rounded, hence a statistical tolerance)
  * checksum:  the BatchNorm partial sums produced by the conv epilogue == sums of the tensor it stored, per channel and group
  * the whole step at B = 64: finite loss, reproducible gradient (two runs agree to 1e-6 of the gradient norm), graph replay
    returns the same loss as the eager step, predict_label == (Vd > Vt) == get_label away from ties."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _U():
    import gpu_util as U
    return U


@pytest.mark.parametrize("n,h,w,cin,cout", [(128, 256, 256, 64, 64), (128, 64, 64, 256, 256)])
def test_conv_scaling_adjoint_and_checksum_at_full_size(n, h, w, cin, cout):
    U = _U()
    torch.manual_seed(1)
    bf = torch.bfloat16
    x = torch.randn(n, h, w, cin, device="cuda").to(bf)
    g = torch.randn(n, h, w, cout, device="cuda").to(bf)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)).to(bf).float()       # bf16-representable weights
    wf, wd = U.pack_conv(wt, U.BF16)
    wf2, _ = U.pack_conv(2 * wt, U.BF16)
    y, st = U.conv3x3(x, wf, cout, U.BF16, U.ENGINE_TC, group_images=n // 2, stats=True)
    y2, _ = U.conv3x3(x, wf2, cout, U.BF16, U.ENGINE_TC, group_images=n // 2, stats=False)
    assert torch.equal(y2, (2 * y.float()).to(bf)) and torch.equal(y2.float(), 2 * y.float())
    # checksum of the epilogue statistics against the stored tensor (fp64 reduction in chunks)
    for grp in range(2):
        part = y[grp * (n // 2):(grp + 1) * (n // 2)]
        s = torch.zeros(cout, dtype=torch.float64, device="cuda")
        q = torch.zeros(cout, dtype=torch.float64, device="cuda")
        for i in range(0, part.shape[0], 8):
            c = part[i:i + 8].double().reshape(-1, cout)
            s += c.sum(0)
            q += (c * c).sum(0)
        assert torch.allclose(st[0, grp], s, rtol=1e-6, atol=1e-3 * float(q.sqrt().max()))
        assert torch.allclose(st[1, grp], q, rtol=1e-6, atol=0)
    # adjoint identities
    dx, _ = U.conv3x3(g, wd, cin, U.BF16, U.ENGINE_TC, stats=False)
    dw = U.conv3x3_wgrad(g, x, U.BF16, U.ENGINE_TC)

    def dot(a, b):
        t = torch.zeros((), dtype=torch.float64, device="cuda")
        for i in range(0, a.shape[0], 8):
            t += (a[i:i + 8].double() * b[i:i + 8].double()).sum()
        return float(t)

    lhs = dot(y, g)
    via_dgrad = dot(x, dx)
    via_wgrad = float((wt.double() * dw.double()).sum())
    scale = (float((y.float() ** 2).sum()) * float((g.float() ** 2).sum())) ** 0.5 / (n * h * w * cout) ** 0.5
    # y and dx are stored in bf16: each element carries a rounding error of relative size <= 2^-9, so the inner products differ
    # by a random sum of about 1.1e-3 * scale (one rounded side) / 1.6e-3 * scale (two rounded sides); 6e-3 is ~4-5 sigma.
    # A structural error (wrong tap, transposed weights, missing filter row) moves them by O(1).
    assert abs(lhs - via_dgrad) <= 6e-3 * scale, (lhs, via_dgrad, scale)
    assert abs(lhs - via_wgrad) <= 6e-3 * scale, (lhs, via_wgrad, scale)


def test_training_step_properties_at_benchmark_batch():
    import onet_b200
    from onet_b200.data import k_clutter_frames
    from onet_b200.trainer import OnetTrainer
    torch.manual_seed(1981)
    x = k_clutter_frames(64, 1, 256, 256, seed=7, n_targets=8).cuda()
    net = onet_b200.Onet(1, True, True, mode="bf16").cuda()
    norms, losses = [], []
    for rep in range(2):
        net.train()
        net.zero_grad()
        Lt, Vt, Ld, Vd, S = net(x)
        loss = net.compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
        loss.backward()
        torch.cuda.synchronize()
        gflat = torch.cat([p.grad.flatten() for p in net.parameters()]).double()
        assert torch.isfinite(gflat).all() and torch.isfinite(loss)
        norms.append(gflat)
        losses.append(float(loss))
        if rep == 0:
            lab = net.predict_label(S)
            assert lab.shape == (64, 256, 256) and lab.dtype == torch.int64
            assert torch.equal(lab, (Vd > Vt).squeeze(1).long())                  # argmax of the 2-way softmax, ties -> 0
            lab2, V = net.get_label(Vt, Vd)                                       # reference :204-219: argmax of softmax([Vt, Vd])
            clear = (Vt - Vd).abs().squeeze(1) > 1e-3 * (Vt.abs() + Vd.abs()).squeeze(1)
            assert torch.equal(lab2[clear], lab[clear])
            assert abs(float(S.sum(dim=1).mean()) - 1.0) < 1e-6
    assert losses[0] == losses[1]
    assert float((norms[0] - norms[1]).norm()) <= 1e-6 * float(norms[0].norm())     # only the order of split-K atomics differs
    del norms
    eager = OnetTrainer(onet_b200.Onet(1, True, True, mode="bf16").cuda(), lr=5e-6, graph=False)
    graph = OnetTrainer(onet_b200.Onet(1, True, True, mode="bf16").cuda(), lr=5e-6, graph=True)
    graph.flat.copy_(eager.flat)
    for b_e, b_g in zip(eager.onet.buffers(), graph.onet.buffers()):
        b_g.copy_(b_e)
    onet_b200.invalidate_packed_weights()
    l_e = [float(eager.step(x)) for _ in range(2)]
    l_g = [float(graph.step(x)) for _ in range(2)]
    assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(l_e, l_g)), (l_e, l_g)
    assert l_e[1] != l_e[0]                                                        # the optimizer moved the weights
