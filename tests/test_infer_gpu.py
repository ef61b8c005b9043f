"""Halo-tiled inference on the GPU (-m gpu): the eval-mode onet_b200.Onet evaluated tile by tile (halo 96 px, clamped at
the frame border) must reproduce the whole-frame forward bit for bit - eval-mode BatchNorm makes the network a fixed
convolutional map and every output pixel sees exactly the same operands in the same order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "bf16", "tf32"])
def test_tiled_inference_is_bit_identical_to_whole_frame(mode):
    import onet_b200
    from onet_b200.data import rayleigh_target_frames
    from onet_b200.infer import TiledPredictor
    torch.manual_seed(17)
    net = onet_b200.Onet(1, True, True, mode=mode).cuda()
    with torch.no_grad():                      # non-trivial eval statistics / affine parameters
        for m in net.modules():
            if hasattr(m, "running_mean"):
                m.running_mean.normal_(0.0, 0.3)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.7, 1.3)
                m.bias.normal_(0.0, 0.2)
    x = rayleigh_target_frames(2, 1, 256, 320, seed=5).cuda()
    net.eval()
    with torch.no_grad():
        _, vt, _, vd, S = net(x)
        lab = net.predict_label(S)
    Vt, Vd, label = TiledPredictor.for_onet(net, tile=64, halo=96, max_batch=3).predict(x)
    torch.cuda.synchronize()
    assert torch.equal(Vt, vt) and torch.equal(Vd, vd)
    assert torch.equal(label, lab)
    # a halo below the receptive-field radius is refused instead of silently changing results
    with pytest.raises(ValueError):
        TiledPredictor.for_onet(net, tile=64, halo=32).predict(x)


def test_pipelined_uint8_masks_equal_the_map_path():
    """`predict_labels` (host frames -> copy stream -> compute -> uint8 masks -> copy-out stream) gives the masks of
    `predict` / `Onet.predict_label`, for whole-frame and tiled plans, host and device inputs, repeated calls (buffer reuse)."""
    import onet_b200
    from onet_b200.data import rayleigh_target_frames
    from onet_b200.infer import TiledPredictor
    torch.manual_seed(23)
    net = onet_b200.Onet(1, True, True, mode="bf16").cuda()
    host = rayleigh_target_frames(5, 1, 128, 160, seed=6).pin_memory()
    x = host.cuda()
    for tile in (512, 64):
        pred = TiledPredictor.for_onet(net, tile=tile, halo=96, max_batch=2)
        _, _, label = pred.predict(x)
        for rep in range(2):
            got = pred.predict_labels(host)
            assert got.dtype == torch.uint8 and not got.is_cuda
            assert torch.equal(got.long(), label.cpu())
        assert torch.equal(pred.predict_labels(x).long(), label)
        # two ranks' shares are disjoint and add up to the whole
        a = pred.predict_labels(host, rank=0, world=2).clone()
        b = pred.predict_labels(host, rank=1, world=2).clone()
        assert torch.equal((a + b).long(), label.cpu())
