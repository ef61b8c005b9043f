"""Halo-tiled inference on the GPU (-m gpu): the eval-mode onet_b200.Onet evaluated tile by tile (halo 96 px, clamped at
the frame border) must reproduce the whole-frame forward bit for bit - eval-mode BatchNorm makes the network a fixed
convolutional map and every output pixel sees exactly the same operands in the same order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
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
