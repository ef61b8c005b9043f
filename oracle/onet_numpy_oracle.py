"""Torch-free restatement of the Onet forward + JSD loss — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

oracle/onet_oracle.py evaluates the path with the same ATen CPU primitives (conv2d, conv_transpose2d, max_pool2d) the
reference dispatches to.  This file re-derives those primitives from their definitions in plain numpy (float64) so that the
semantics the CUDA kernels are held to — cross-correlation orientation and zero padding of the 3x3 conv, biased batch variance
and eps placement of BatchNorm, floor-mode 2x2 pooling, the (Cin, Cout, kh, kw) layout and stride-2 scatter of the transposed
conv, the F.pad placement, the channel order of the concat, the quirky piecewise softplus — are pinned independently of torch:
tests/test_oracle_golden.py compares it with the ATen oracle (forward maps, loss) and checks the ATen oracle's autograd
gradient against central finite differences of THIS loss.

Follows /root/reference/source_code/Onet_vanilla_20240606.py: DoubleConv :39-58, Down :61-72, Up :75-101, UNet.forward
:142-153, Onet.forward :174-191, log1pexp :237-251, jensen_shannon_divergence :221-235, compute_loss :253-267."""
import numpy as np

ENCODER = ["down1", "down2", "down3", "down4"]
DECODER = ["up1", "up2", "up3", "up4"]
BN_EPS = 1e-5


def _prefix(block):
    if block == "inc":
        return "inc.double_conv"
    if block.startswith("down"):
        return f"{block}.maxpool_conv.1.double_conv"
    return f"{block}.conv.double_conv"


def conv3x3(x, w):
    """nn.Conv2d(k=3, padding=1, bias=False): out[n,o,y,x] = sum_{c,i,j} w[o,c,i,j] * xpad[n,c,y+i,x+j] (cross-correlation)."""
    n, c, h, wd = x.shape
    xp = np.zeros((n, c, h + 2, wd + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, w.shape[0], h, wd), dtype=np.float64)
    for i in range(3):
        for j in range(3):
            out += np.einsum("nchw,oc->nohw", xp[:, :, i:i + h, j:j + wd], w[:, :, i, j], optimize=True)
    return out


def batchnorm_train(x, gamma, beta):
    """nn.BatchNorm2d in training mode: per-channel mean and BIASED variance over (N, H, W), eps inside the square root."""
    mean = x.mean(axis=(0, 2, 3), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(0, 2, 3), keepdims=True)
    return (x - mean) / np.sqrt(var + BN_EPS) * gamma[None, :, None, None] + beta[None, :, None, None]


def maxpool2x2(x):
    """nn.MaxPool2d(2): floor mode, odd trailing row / column dropped."""
    n, c, h, w = x.shape
    h2, w2 = h // 2, w // 2
    return x[:, :, :2 * h2, :2 * w2].reshape(n, c, h2, 2, w2, 2).max(axis=(3, 5))


def conv_transpose2x2(x, w, b):
    """nn.ConvTranspose2d(Cin, Cout, 2, stride=2): out[n,o,2y+i,2x+j] = b[o] + sum_c x[n,c,y,x] * w[c,o,i,j]."""
    n, c, h, wd = x.shape
    out = np.zeros((n, w.shape[1], 2 * h, 2 * wd), dtype=np.float64)
    for i in range(2):
        for j in range(2):
            out[:, :, i::2, j::2] = np.einsum("nchw,co->nohw", x, w[:, :, i, j], optimize=True)
    return out + b[None, :, None, None]


def double_conv(st, block, x):                                            # :39-58
    p = _prefix(block)
    for conv_i, bn_i in ((0, 1), (3, 4)):
        x = conv3x3(x, st[f"{p}.{conv_i}.weight"])
        x = np.maximum(batchnorm_train(x, st[f"{p}.{bn_i}.weight"], st[f"{p}.{bn_i}.bias"]), 0.0)
    return x


def unet_forward(st, x):                                                  # :142-153, Up :88-101
    x1 = double_conv(st, "inc", x)
    skips, h = [x1], x1
    for name in ENCODER:
        h = double_conv(st, name, maxpool2x2(h))
        skips.append(h)
    y = skips[-1]
    for i, name in enumerate(DECODER):
        skip = skips[3 - i]
        up = conv_transpose2x2(y, st[f"{name}.up.weight"], st[f"{name}.up.bias"])
        dy, dx = skip.shape[2] - up.shape[2], skip.shape[3] - up.shape[3]
        up = np.pad(up, ((0, 0), (0, 0), (dy // 2, dy - dy // 2), (dx // 2, dx - dx // 2)))      # F.pad :95-96
        y = double_conv(st, name, np.concatenate([skip, up], axis=1))                            # cat([x2, x1]) :100
    return x1, y


def log1pexp(x):                                                          # :237-251, see oracle/onet_oracle.py::log1pexp
    lo = x <= -37.0
    t = np.where(lo, np.exp(np.where(lo, x, 0.0)), x)
    mid = (t > -37.0) & (t <= 18.0)
    t = np.where(mid, np.log(1 + np.exp(np.where(mid, t, 0.0))), t)
    hi = (t > 18.0) & (t < 33.3)
    return np.where(hi, t + np.exp(-np.where(hi, t, 0.0)), t)


def onet_forward_loss(st, x, st_dwn=None):
    """Twin forward in training mode + compute_loss; st: {key: float64 array} of ONE UNet (weight-shared twin unless st_dwn)."""
    x = np.asarray(x, dtype=np.float64)
    Lt, Ht = unet_forward(st, x)
    Vt = (Lt * Ht).sum(axis=1, keepdims=True)
    Ld, Hd = unet_forward(st if st_dwn is None else st_dwn, np.clip(1 - x, 0, 1))
    Vd = (Ld * Hd).sum(axis=1, keepdims=True)
    m = np.maximum(Vt, Vd)
    et, ed = np.exp(Vt - m), np.exp(Vd - m)
    St, Sd = et / (et + ed), ed / (et + ed)

    def jsd(L, S, Sp):                                                    # :221-235
        a = L.sum(axis=1)
        return -log1pexp(-a * S[:, 0]).mean() - log1pexp(a * Sp[:, 0]).mean()

    loss = -(jsd(Lt, St, Sd) + jsd(Ld, Sd, St)) / 2                       # :253-267
    return dict(Lt=Lt, Vt=Vt, Ld=Ld, Vd=Vd, St=St, Sd=Sd, loss=float(loss))
