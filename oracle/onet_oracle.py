"""CPU oracle for the Onet hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file, and only as the checker / the timed CPU baseline.  The product
package `onet_b200/` never imports it and has no CPU fallback.

What it is: an independent restatement of the reference's algorithm for the hot path,
`/root/reference/source_code/Onet_vanilla_20240606.py:39-267`, written as plain functions
over an explicit parameter dictionary (no nn.Module, no in-place tricks).  The path's
arithmetic lives in third-party PyTorch (conv / batch-norm / max-pool / transposed-conv
semantics; the reference pins it only as `torch>=1.7.0`, requirements.txt:3), so the
restatement uses the same ATen CPU primitives through `torch.nn.functional` in FP32 and
restates everything the reference itself adds: block wiring, the twin call order with the
shared BatchNorm buffers, the dot-product head, the 2-way softmax, the JSD loss and the
exact (quirky) piecewise softplus.  The head + loss are additionally restated in pure
numpy with closed-form gradients (`head_loss_numpy`) as the checker for the fused CUDA
head kernels.

Pinning: the reference ships no tests, fixtures or golden vectors, so the
oracle is pinned against outputs of the UNMODIFIED reference module run in the build
container: `tests/golden/make_golden.py` imports the reference (oracle/ref_import.py), runs it on
seeded inputs/weights and commits the results under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those vectors (and, when
/root/reference is present, against the live reference).
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, Onet_vanilla_20240606.py:48,52
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default

# (name, in_channels, out_channels) of the nine DoubleConv blocks, Onet_vanilla_20240606.py:111-120
ENCODER = [("down1", 64, 128), ("down2", 128, 256), ("down3", 256, 512), ("down4", 512, 1024)]
DECODER = [("up1", 1024, 512), ("up2", 512, 256), ("up3", 256, 128), ("up4", 128, 64)]


# ----------------------------------------------------------------------------------------
# reduced-precision emulation (checker for the bf16 / tf32 CUDA modes)
# ----------------------------------------------------------------------------------------
# The CUDA path's reduced-precision modes round at known points.  `emulate="bf16"` / `"tf32"` restates the SAME algorithm
# with a rounding at exactly those points, in plain PyTorch on the CPU, so that the CUDA kernels can be held to it tightly
# (what is left is fp32 accumulation order); against the FP32 reference the same modes are only held to the north_star
# tolerances.  bf16 mode (onet_b200/csrc: conv epilogues, bn_relu_value, head_bwd_kernel, bn_bwd_*_kernel):
#   forward : twin input, packed weights, every raw conv output Y, every post-ReLU activation and every up-conv output are
#             stored as bf16; BatchNorm statistics are taken from the stored (rounded) Y; accumulation is fp32.  Exception: the
#             first conv of each U-Net, whose Y (and dY) is never stored - recomputed in fp32 wherever it is needed.
#   backward: dL / dH from the head, every dY (BatchNorm backward) and every data gradient dX (conv / up-conv dgrad) are
#             stored as bf16; weight / bias / BatchNorm parameter gradients stay fp32.
# tf32 mode: storage is fp32 everywhere; only the tensor-core operands (activations, weights, output gradients of the 3x3
#             and transposed convolutions except the first 1->64 / 3->64 layer, which runs on CUDA cores) lose their low 13
#             mantissa bits (tcgen05.mma kind::tf32 reads the upper 19 bits of each fp32 operand: truncation).
def _round_bf16(t):
    return t.bfloat16().float()


def _trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


class _Ste(torch.autograd.Function):
    """y = f_fwd(x) in forward, g -> f_bwd(g) in backward (f = None: identity)."""

    @staticmethod
    def forward(ctx, x, f_fwd, f_bwd):
        ctx.f_bwd = f_bwd
        return x if f_fwd is None else f_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return (g if ctx.f_bwd is None else ctx.f_bwd(g)), None, None


class _Policy:
    """Where a mode rounds.  weight: a conv / up-conv weight operand; conv_in: the activation operand of a convolution
    (its gradient is that convolution's data gradient); conv_out: a raw convolution output (its gradient is the dY the
    convolution's backward kernels read); store: a stored activation; head_in: L / H as the head reads them."""

    def __init__(self, mode):
        assert mode in ("bf16", "tf32")
        self.mode = mode

    def image(self, t):
        return _round_bf16(t) if self.mode == "bf16" else t

    def weight(self, t, first=False):
        if self.mode == "bf16":
            return _Ste.apply(t, _round_bf16, None)
        return t if first else _Ste.apply(t, _trunc_tf32, None)

    def conv_in(self, t, first=False):
        if self.mode == "bf16":
            return _Ste.apply(t, None, _round_bf16)              # data gradient stored as bf16
        return t if first else _Ste.apply(t, _trunc_tf32, None)   # operand truncation; data gradient stays fp32

    def conv_out(self, t, first=False, unstored=False):
        if self.mode == "bf16":
            if unstored:      # first conv of a U-Net: y and dY are recomputed / kept in registers, never rounded
                return t
            return _Ste.apply(t, _round_bf16, _round_bf16)       # Y stored as bf16, dY stored as bf16
        return t if first else _Ste.apply(t, None, _trunc_tf32)   # dY is a tensor-core operand of dgrad and wgrad

    def upconv_out(self, t):
        if self.mode == "bf16":
            return _Ste.apply(t, _round_bf16, None)              # bf16(acc + bias); its gradient is a slice of d(concat)
        return _Ste.apply(t, None, _trunc_tf32)

    def store(self, t):
        return _Ste.apply(t, _round_bf16, None) if self.mode == "bf16" else t

    def head_in(self, t):
        return _Ste.apply(t, None, _round_bf16) if self.mode == "bf16" else t


def _policy(emulate):
    return None if emulate in (None, "fp32") else _Policy(emulate)


def _dc_prefix(block):
    if block == "inc":
        return "inc.double_conv"
    if block.startswith("down"):
        return f"{block}.maxpool_conv.1.double_conv"
    return f"{block}.conv.double_conv"


def unet_layout(in_chns):
    """[(key, shape, kind)] for every parameter/buffer of one UNet in state_dict order
    (Onet_vanilla_20240606.py:104-123).  kind in {conv, bn_w, bn_b, rm, rv, nbt, convT_w, convT_b}."""
    out = []

    def dc(block, cin, cout):
        p = _dc_prefix(block)
        for conv_i, bn_i, ci in ((0, 1, cin), (3, 4, cout)):
            out.append((f"{p}.{conv_i}.weight", (cout, ci, 3, 3), "conv"))
            out.append((f"{p}.{bn_i}.weight", (cout,), "bn_w"))
            out.append((f"{p}.{bn_i}.bias", (cout,), "bn_b"))
            out.append((f"{p}.{bn_i}.running_mean", (cout,), "rm"))
            out.append((f"{p}.{bn_i}.running_var", (cout,), "rv"))
            out.append((f"{p}.{bn_i}.num_batches_tracked", (), "nbt"))

    dc("inc", in_chns, 64)
    for name, cin, cout in ENCODER:
        dc(name, cin, cout)
    for name, cin, cout in DECODER:
        out.append((f"{name}.up.weight", (cin, cin // 2, 2, 2), "convT_w"))
        out.append((f"{name}.up.bias", (cin // 2,), "convT_b"))
        dc(name, cin, cout)
    return out


def init_state(in_chns=1, seed=1981, binit=True):
    """Deterministic UNet state (parameters + BN buffers) keyed like the reference UNet's
    state_dict.  Restates `UNet._initialize_weights` (Onet_vanilla_20240606.py:125-140):
    Kaiming-normal(fan_in, relu) on Conv2d only, BN weight 1 / bias 0; ConvTranspose2d keeps
    PyTorch's default init (kaiming_uniform(a=sqrt(5)) weight, U(-1/sqrt(fan_in), ..) bias, with
    fan_in = weight.size(1) * kh * kw).  The values are NOT bit-identical to what the reference
    constructor draws (different RNG consumption order) — parity tests always copy one state into
    both implementations."""
    g = torch.Generator().manual_seed(seed)
    st = OrderedDict()
    for key, shape, kind in unet_layout(in_chns):
        if kind == "conv":
            fan_in = shape[1] * 9
            if binit:
                st[key] = torch.randn(shape, generator=g, dtype=torch.float32) * math.sqrt(2.0 / fan_in)
            else:  # nn.Conv2d default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                st[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) / math.sqrt(fan_in)
        elif kind == "bn_w":
            st[key] = torch.ones(shape)
        elif kind in ("bn_b", "rm"):
            st[key] = torch.zeros(shape)
        elif kind == "rv":
            st[key] = torch.ones(shape)
        elif kind == "nbt":
            st[key] = torch.zeros((), dtype=torch.int64)
        elif kind == "convT_w":
            fan_in = shape[1] * 4
            bound = 1.0 / math.sqrt(fan_in)
            st[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        elif kind == "convT_b":
            bound = 1.0 / math.sqrt(shape[0] * 4)  # fan_in = weight.size(1) * kh * kw = out_channels * 4
            st[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    return st


def perturb_bn_affine(st, seed=7):
    """Give BN weight/bias non-trivial values so parity tests exercise gamma/beta paths."""
    g = torch.Generator().manual_seed(seed)
    for k in st:
        if k.endswith(".weight") and st[k].dim() == 1:
            st[k] = 1.0 + 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
        elif k.endswith(".bias") and st[k].dim() == 1 and ".up." not in k:
            st[k] = 0.2 * (torch.rand(st[k].shape, generator=g) - 0.5)
    return st


# ----------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------
def _bn(st, prefix, x, training, taps=None):
    """nn.BatchNorm2d semantics: batch statistics (biased variance) in training and an update of
    the running buffers with the UNBIASED variance, momentum 0.1; running statistics in eval."""
    w, b = st[f"{prefix}.weight"], st[f"{prefix}.bias"]
    if training:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        with torch.no_grad():
            rm, rv = st[f"{prefix}.running_mean"], st[f"{prefix}.running_var"]
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * (n / max(n - 1, 1)))
            st[f"{prefix}.num_batches_tracked"] += 1
    else:
        mean, var = st[f"{prefix}.running_mean"], st[f"{prefix}.running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    y = (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]
    return y


def _bn_relu_emulated(st, prefix, y, training, q):
    """BatchNorm + ReLU as the CUDA path evaluates it (bn_finalize_kernel + bn_relu_value, elementwise.cuh): statistics in
    double from the stored Y, scale = gamma * invstd and shift = beta - mean * scale in fp32, act = relu(y * scale + shift),
    stored in the mode's storage type.  Same function of y as `_bn` + relu up to fp32 rounding."""
    w, b = st[f"{prefix}.weight"], st[f"{prefix}.bias"]
    if training:
        n = y.numel() // y.shape[1]
        yd = y.double()
        mean = yd.mean(dim=(0, 2, 3))
        var = yd.var(dim=(0, 2, 3), unbiased=False)
        with torch.no_grad():
            rm, rv = st[f"{prefix}.running_mean"], st[f"{prefix}.running_var"]
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().float())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var.detach() * (n / max(n - 1, 1))).float())
            st[f"{prefix}.num_batches_tracked"] += 1
        invstd = torch.rsqrt(var + BN_EPS).float()
        mean = mean.float()
    else:
        mean = st[f"{prefix}.running_mean"]
        invstd = torch.rsqrt(st[f"{prefix}.running_var"] + BN_EPS)
    sc = w * invstd
    sh = b - mean * sc
    return q.store(torch.relu(y * sc[None, :, None, None] + sh[None, :, None, None]))


def _forced(x, force, key, taps):
    """Teacher forcing: the VALUE of x becomes force[key] (the tensor another implementation stored at this point) while the
    gradient still flows through x's own graph; taps[key + ".own"] keeps what this implementation computed from the forced
    inputs.  With every raw convolution output and up-conv output forced, all data-dependent decisions of the forward pass
    (BatchNorm statistics, ReLU masks, max-pool arg-max) are the other implementation's, so the backward pass is compared as
    the LINEAR map it is - free of the network's ill-conditioning at random init (a 1e-6 relative perturbation of the forward
    moves the free-running gradient by 3e-3, bf16 rounding flips by 0.2)."""
    if force is None or key not in force:
        return x
    if taps is not None:
        taps[key + ".own"] = x.detach()
    return x + (force[key].to(x.dtype) - x).detach()


def double_conv(st, block, x, training, taps=None, q=None, force=None):
    """Onet_vanilla_20240606.py:39-58: (conv3x3 pad 1 no bias -> BN -> ReLU) x 2."""
    p = _dc_prefix(block)
    for conv_i, bn_i in ((0, 1), (3, 4)):
        if q is None:
            x = F.conv2d(x, st[f"{p}.{conv_i}.weight"], None, padding=1)
        else:
            first = block == "inc" and conv_i == 0       # the 1 -> 64 / 3 -> 64 layer runs on CUDA cores in every mode
            unstored = first                              # csrc/first_layer.cuh: this layer's Y / dY is never materialised (closed form)
            x = q.conv_out(F.conv2d(q.conv_in(x, first), q.weight(st[f"{p}.{conv_i}.weight"], first), None, padding=1), first, unstored)
        x = _forced(x, force, f"{p}.{conv_i}.raw", taps)
        if taps is not None:
            taps[f"{p}.{conv_i}.raw"] = x
        x = torch.relu(_bn(st, f"{p}.{bn_i}", x, training)) if q is None else _bn_relu_emulated(st, f"{p}.{bn_i}", x, training, q)
        if taps is not None:
            taps[f"{p}.{bn_i}.act"] = x
    return x


def up_block(st, name, x1, x2, training, taps=None, q=None, force=None):
    """Onet_vanilla_20240606.py:75-101 (bilinear=False): ConvTranspose2d(C, C/2, 2, 2) with bias,
    zero-pad to the skip's size, cat([skip, up]) and DoubleConv."""
    if q is None:
        x1 = F.conv_transpose2d(x1, st[f"{name}.up.weight"], st[f"{name}.up.bias"], stride=2)
    else:
        x1 = q.upconv_out(F.conv_transpose2d(q.conv_in(x1), q.weight(st[f"{name}.up.weight"]), st[f"{name}.up.bias"], stride=2))
    x1 = _forced(x1, force, f"{name}.up.out", taps)
    if taps is not None:
        taps[f"{name}.up.out"] = x1
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    x = torch.cat([x2, x1], dim=1)
    if taps is not None:
        taps[f"{name}.cat"] = x
    return double_conv(st, name, x, training, taps, q, force)


def unet_forward(st, x, training=True, taps=None, q=None, force=None):
    """Onet_vanilla_20240606.py:142-153: returns (x1, y1) = (first-block output, last-block output)."""
    x1 = double_conv(st, "inc", x, training, taps, q, force)
    skips = [x1]
    h = x1
    for name, _, _ in ENCODER:
        h = F.max_pool2d(h, 2)
        h = double_conv(st, name, h, training, taps, q, force)
        skips.append(h)
    y = skips[-1]
    for i, (name, _, _) in enumerate(DECODER):
        y = up_block(st, name, y, skips[3 - i], training, taps, q, force)
    return x1, y


def onet_forward(st_top, x, training=True, st_dwn=None, bias=0.0, taps=None, emulate=None, force=None):
    """Onet.forward, Onet_vanilla_20240606.py:174-191.  `st_dwn=None` is the weight-shared twin
    (`bshare=True`, :163-164): the SAME state (parameters and BN running buffers) is used for
    both branches, top branch first.  `emulate` = "bf16" / "tf32": the same algorithm with the CUDA path's
    reduced-precision roundings (see `_Policy`); None = the reference's FP32 arithmetic.  `force` = {"top": {...}, "dwn": {...}}
    of tensors keyed like `taps` ("<prefix>.<conv>.raw", "<up block>.up.out"): teacher forcing, see `_forced`."""
    st_dwn = st_top if st_dwn is None else st_dwn
    q = _policy(emulate)
    tt = {} if taps is not None else None
    td = {} if taps is not None else None
    Xd = torch.clip(1 - x + bias, 0, 1)
    if q is not None:
        x, Xd = q.image(x), q.image(Xd)
    ft, fd = (None, None) if force is None else (force.get("top"), force.get("dwn"))
    Lt, Ht = unet_forward(st_top, x, training, tt, q, ft)
    if q is not None:        # the head's gradient w.r.t. L (both uses: V and the loss's channel sum) and H is stored rounded
        Lt, Ht = q.head_in(Lt), q.head_in(Ht)
    Vt = (Lt * Ht).sum(dim=1, keepdim=True)
    Ld, Hd = unet_forward(st_dwn, Xd, training, td, q, fd)
    if q is not None:
        Ld, Hd = q.head_in(Ld), q.head_in(Hd)
    Vd = (Ld * Hd).sum(dim=1, keepdim=True)
    S = torch.softmax(torch.cat([Vt, Vd], dim=1), dim=1)
    if taps is not None:
        taps["top"], taps["dwn"] = tt, td
        taps["Ht"], taps["Hd"] = Ht, Hd
    return Lt, Vt, Ld, Vd, S


def log1pexp(x):
    """Exact semantics of Onet.log1pexp (Onet_vanilla_20240606.py:237-251), without mutation.

    The reference mutates x in place in three masked steps and evaluates each mask on the
    ALREADY MUTATED tensor, hence: x <= -37 is first replaced by exp(x) ~ 0+, which the second
    mask (-37, 18] then matches again and maps to log(1 + exp(0+)) = ln 2 (not ~0)."""
    lo = x <= -37.0
    t = torch.where(lo, torch.exp(torch.where(lo, x, torch.zeros_like(x))), x)
    mid = (t > -37.0) & (t <= 18.0)
    t = torch.where(mid, torch.log(1 + torch.exp(torch.where(mid, t, torch.zeros_like(t)))), t)
    hi = (t > 18.0) & (t < 33.3)
    t = torch.where(hi, t + torch.exp(-torch.where(hi, t, torch.zeros_like(t))), t)
    return t


def jensen_shannon_divergence(Li, Si, Sprime):
    """Onet_vanilla_20240606.py:221-235.  The einsum "bpxy,bpxy->bxy" with Si of shape (B,1,H,W)
    broadcasts the size-1 p dimension: LS = Si * sum_p Li_p."""
    a = Li.sum(dim=1)
    LS = a * Si[:, 0]
    LSp = a * Sprime[:, 0]
    return -1 * log1pexp(-1 * LS).mean() - log1pexp(LSp).mean()


def compute_loss(Lt, St, Ld, Sd):
    """Onet.compute_loss, Onet_vanilla_20240606.py:253-267."""
    return -(jensen_shannon_divergence(Lt, St, Sd) + jensen_shannon_divergence(Ld, Sd, St)) / 2


def predict_label(S):
    """Onet.predict_label, Onet_vanilla_20240606.py:193-202: argmax over the 2 channels, ties -> 0."""
    return torch.argmax(S, dim=1)


def train_step_outputs(st, x, st_dwn=None, emulate=None, force=None, taps=None):
    """One reference training-step's forward + loss + backward (Train_Onet_on_simclutter_20250407.py:
    209-217) on a copy of `st` with autograd; returns (outputs dict, grads dict, new state)."""
    st = OrderedDict((k, v.clone()) for k, v in st.items())
    leaves = [k for k, v in st.items() if v.dtype.is_floating_point and "running" not in k]
    for k in leaves:
        st[k].requires_grad_(True)
    sd = None
    if st_dwn is not None:
        sd = OrderedDict((k, v.clone()) for k, v in st_dwn.items())
        for k in leaves:
            sd[k].requires_grad_(True)
    Lt, Vt, Ld, Vd, S = onet_forward(st, x, True, sd, emulate=emulate, force=force, taps=taps)
    loss = compute_loss(Lt, S[:, 0:1], Ld, S[:, 1:2])
    loss.backward()
    grads = OrderedDict((k, st[k].grad.detach().clone()) for k in leaves)
    out = dict(Lt=Lt.detach(), Vt=Vt.detach(), Ld=Ld.detach(), Vd=Vd.detach(), S=S.detach(),
               loss=loss.detach())
    new_state = OrderedDict((k, v.detach()) for k, v in st.items())
    if sd is not None:
        gd = OrderedDict((k, sd[k].grad.detach().clone()) for k in leaves)
        return out, grads, new_state, gd, OrderedDict((k, v.detach()) for k, v in sd.items())
    return out, grads, new_state


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, amsgrad=False) as used at
    Train_Onet_on_simclutter_20250407.py:181-182; `step` is the 1-based step count."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def tensor_normal_per_frame(x):
    """utils_20231218.py:673-689: per-(b,c) min-max scaling with + np.spacing(1) in the denominator."""
    nb, nc, h, w = x.shape
    v = x.reshape(nb, nc, h * w)
    mn = v.min(dim=-1, keepdim=True)[0]
    mx = v.max(dim=-1, keepdim=True)[0]
    return ((v - mn) / (mx - mn + np.spacing(1))).reshape(nb, nc, h, w)


# ----------------------------------------------------------------------------------------
# pure-numpy head + loss with closed-form gradients (checker for the fused CUDA head kernels)
# ----------------------------------------------------------------------------------------
def _sp_np(x):
    """(value, derivative) of the reference's piecewise softplus in float32 numpy, branch by
    branch as Onet_vanilla_20240606.py:245-250 evaluates them."""
    x = x.astype(np.float32)
    val = x.copy()
    der = np.ones_like(x)
    lo = x <= np.float32(-37.0)
    e = np.exp(np.where(lo, x, 0).astype(np.float32))
    # lo: value log(1+exp(exp(x))), derivative sigmoid(exp(x))*exp(x)
    ee = np.exp(e)
    val = np.where(lo, np.log(np.float32(1) + ee), val)
    der = np.where(lo, ee / (np.float32(1) + ee) * e, der)
    mid = (~lo) & (x <= np.float32(18.0))
    em = np.exp(np.where(mid, x, 0).astype(np.float32))
    val = np.where(mid, np.log(np.float32(1) + em), val)
    der = np.where(mid, em / (np.float32(1) + em), der)
    hi = (x > np.float32(18.0)) & (x < np.float32(33.3))
    eh = np.exp(-np.where(hi, x, 0).astype(np.float32))
    val = np.where(hi, x + eh, val)
    der = np.where(hi, np.float32(1) - eh, der)
    return val.astype(np.float32), der.astype(np.float32)


def head_loss_numpy(Lt, Ht, Ld, Hd):
    """numpy restatement of Onet.forward's head (:176-189) + compute_loss (:253-267) with the
    closed-form gradients of SURVEY.md §8a row A10.  Inputs (B,64,H,W) float32 arrays.
    Returns dict(Vt, Vd, St, Sd, loss, dLt, dHt, dLd, dHd)."""
    Lt, Ht, Ld, Hd = (np.asarray(t, dtype=np.float32) for t in (Lt, Ht, Ld, Hd))
    Vt = (Lt * Ht).sum(1, dtype=np.float32)
    Vd = (Ld * Hd).sum(1, dtype=np.float32)
    St = (1.0 / (1.0 + np.exp(-(Vt - Vd).astype(np.float64)))).astype(np.float32)
    Sd = (1.0 / (1.0 + np.exp(-(Vd - Vt).astype(np.float64)))).astype(np.float32)
    a = Lt.sum(1, dtype=np.float32)
    b = Ld.sum(1, dtype=np.float32)
    n = a.size
    v1, d1 = _sp_np(-a * St)
    v2, d2 = _sp_np(a * Sd)
    v3, d3 = _sp_np(-b * Sd)
    v4, d4 = _sp_np(b * St)
    loss = (v1.astype(np.float64) + v2 + v3 + v4).sum() / (2.0 * n)
    c = np.float32(1.0 / (2.0 * n))
    g_a = (-St * d1 + Sd * d2) * c
    g_b = (-Sd * d3 + St * d4) * c
    g_St = (-a * d1 + b * d4) * c
    g_Sd = (a * d2 - b * d3) * c
    g_Vt = St * Sd * (g_St - g_Sd)
    return dict(Vt=Vt, Vd=Vd, St=St, Sd=Sd, loss=np.float32(loss),
                g_a=g_a, g_b=g_b, g_Vt=g_Vt,
                dLt=g_a[:, None] + g_Vt[:, None] * Ht, dHt=g_Vt[:, None] * Lt,
                dLd=g_b[:, None] - g_Vt[:, None] * Hd, dHd=-g_Vt[:, None] * Ld)


# ----------------------------------------------------------------------------------------
# synthetic frames (measurement inputs) — oracle-side generator used by tests
# ----------------------------------------------------------------------------------------
def rayleigh_frames(batch, chans, h, w, seed=1981, targets=6, snr_db=2.0):
    """Rayleigh(sigma=1) clutter + a few Gaussian extended targets, per-frame min-max normalised
    (shape-only stand-in for Rayleigh_bg_Gaussian_EOT_generator_20230208.py:219-249)."""
    rng = np.random.default_rng(seed)
    x = rng.rayleigh(1.0, size=(batch, chans, h, w)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    amp = np.float32(10 ** (snr_db / 20.0) * 3.0)
    for bi in range(batch):
        for _ in range(targets):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            sy, sx = rng.uniform(1.5, 5.0), rng.uniform(1.5, 5.0)
            x[bi] += amp * np.exp(-((yy - cy) ** 2 / (2 * sy ** 2) + (xx - cx) ** 2 / (2 * sx ** 2)))
    t = torch.from_numpy(x)
    return tensor_normal_per_frame(t)
