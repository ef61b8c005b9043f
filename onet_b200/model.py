"""Drop-in `Onet` for the reference's twin U-Net hot path, running on hand-written sm_100a kernels.

Host-side mirror of `/root/reference/source_code/Onet_vanilla_20240606.py` (class surface, argument meaning,
assertions and `state_dict` keys are the reference's; the implementation is not):

    Onet(in_chns=1, binit=False, bshare=True)            reference :157-172
    forward(X) -> (Lt, Vt, Ld, Vd, S)                     reference :174-191
    compute_loss(Lt, St, Ld, Sd) -> 0-dim loss            reference :253-267
    jensen_shannon_divergence / log1pexp                  reference :221-251
    predict_label(S) / get_label(Vt, Vd)                  reference :193-219
    attributes topu, dwnu, softmax, bias; state_dict keys topu.* / dwnu.* as the reference's

Everything numerical goes through the C ABI in `include/onet_b200.h` (`onet_b200/_lib.py`).  Both twin
branches run as ONE batch of 2B images (top branch first) with per-branch BatchNorm statistics.  Activations
are NHWC; the encoder's BN+ReLU pass writes the skip tensor directly into the decoder's concat buffer and the
pooled map in the same pass, and the transposed convolution writes the other half of that buffer, so no `cat`
copy exists.  There is no CPU / eager-PyTorch fallback: without a CUDA device or the built library the module
raises.
"""
import math
import os
import weakref
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from ._lib import BF16, ENGINE_SIMT, ENGINE_TC, F32, call, ptr

# (block, in_channels, out_channels) in forward order, reference :111-120
_ENC = [("down1", 64, 128), ("down2", 128, 256), ("down3", 256, 512), ("down4", 512, 1024)]
_DEC = [("up1", 1024, 512), ("up2", 512, 256), ("up3", 256, 128), ("up4", 128, 64)]

# mode -> (C-ABI dtype, torch storage dtype).  "tf32": fp32 storage everywhere, the tcgen05 kernels read the fp32 operands as
# TF32 (C-ABI: dtype ONET_F32 with engine ONET_ENGINE_TC) - the fast mode that meets the north_star tolerances literally.
_DTYPES = {"fp32": (F32, torch.float32), "bf16": (BF16, torch.bfloat16), "tf32": (F32, torch.float32)}

_PACK_GEN = [0]
_STATS_GEN = [0]        # bumped by every training-mode forward: onet_bn_finalize updates the running statistics through raw pointers

# Backward schedule: the weight-gradient kernels (tensor-core bound, one persistent CTA of <= 171 KB shared memory and
# <= 96 registers x 192 threads per SM) run on a second stream NEXT TO the BatchNorm-backward kernels of the following
# layer (HBM bound, small CTAs), which they do not depend on.  ONET_NO_WGRAD_OVERLAP=1 restores the serial order.
_SIDE_STREAMS = {}

# BatchNorm-backward reduce folded into the producing dgrad epilogue for layers of at least this many channels
# (`_conv_bn_relu_bwd`); ONET_NO_BNRED_FUSION=1 disables it.  Measured (profiles/r1_ab_bnred_fusion.txt): the 8 epilogue warps
# of the 64- and 128-wide tile kernels have no slack for the extra work (the fused dgrad gets 2-2.6x slower), the 256-wide
# CTA-pair kernel absorbs it (+14 % on its 5 launches against -0.43 ms of BatchNorm passes), hence 256.
_BNRED_MIN_C = int(os.environ.get("ONET_BNRED_MIN_C", "256"))


# FP32 verification mode: split-K partial sums of the CUDA-core weight-gradient kernels go through this per-device workspace
# and are added in a fixed order (onet_set_splitk_workspace), so that two runs give bit-identical gradients.
_SPLITK_WS = {}
_SPLITK_FLOATS = 16 << 20


def _register_splitk_workspace(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SPLITK_WS:
        _SPLITK_WS[key] = torch.empty(_SPLITK_FLOATS, dtype=torch.float32, device=dev)
        call("onet_set_splitk_workspace", ptr(_SPLITK_WS[key]), _SPLITK_FLOATS, device=dev)


def _side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _SIDE_STREAMS[key]


def invalidate_packed_weights():
    """Call after parameters were modified through raw pointers (e.g. the fused Adam kernel), which does not bump
    the tensors' autograd version counters; the packed operand copies (and the cached eval-mode BatchNorm affines) are rebuilt on
    the next forward."""
    _PACK_GEN[0] += 1


# ------------------------------------------------------------------------------------------------------
# parameter containers: same attribute tree / state_dict keys as the reference modules, no compute
# ------------------------------------------------------------------------------------------------------
class _Conv3x3(nn.Module):            # stands where nn.Conv2d(k=3, p=1, bias=False) stands in the reference (:47,51)
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))     # nn.Conv2d default


class _BatchNorm(nn.Module):          # nn.BatchNorm2d (:48,52): eps 1e-5, momentum 0.1, affine, running stats
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.momentum = 0.1


class _UpConv(nn.Module):             # nn.ConvTranspose2d(C, C//2, 2, 2) with bias (:86)
    def __init__(self, cin):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cin, cin // 2, 2, 2))
        self.bias = nn.Parameter(torch.empty(cin // 2))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt((cin // 2) * 4)
        nn.init.uniform_(self.bias, -bound, bound)


class DoubleConv(nn.Module):          # reference :39-58, keys double_conv.{0,1,3,4}.*
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        self.double_conv = nn.ModuleDict({"0": _Conv3x3(in_channels, mid), "1": _BatchNorm(mid),
                                          "3": _Conv3x3(mid, out_channels), "4": _BatchNorm(out_channels)})

    def layers(self):
        d = self.double_conv
        return [(d["0"], d["1"]), (d["3"], d["4"])]


class Down(nn.Module):                # reference :61-72, keys maxpool_conv.1.double_conv.*
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.ModuleDict({"1": DoubleConv(in_channels, out_channels)})


class Up(nn.Module):                  # reference :75-101 (bilinear=False), keys up.*, conv.double_conv.*
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.up = _UpConv(in_channels)
        self.conv = DoubleConv(in_channels, out_channels)


class UNet(nn.Module):
    """Parameter tree of the reference UNet (:104-153).  Calling it runs the single-branch forward and
    returns (x1, y1) like the reference; Onet.forward uses the twin-batched engine instead."""

    def __init__(self, n_channels=1, n_classes=1, binit=False, bilinear=False):
        super().__init__()
        if bilinear:
            raise NotImplementedError("Onet always builds UNet(bilinear=False) (reference :162,166)")
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self.inc = DoubleConv(n_channels, 64)
        for name, cin, cout in _ENC:
            setattr(self, name, Down(cin, cout))
        for name, cin, cout in _DEC:
            setattr(self, name, Up(cin, cout))
        if binit:
            self._initialize_weights()

    def _initialize_weights(self, mode="fan_in"):
        # reference :125-140: Kaiming-normal on the 3x3 convs only, BN weight 1 / bias 0, up-convs keep defaults
        for m in self.modules():
            if isinstance(m, _Conv3x3):
                nn.init.kaiming_normal_(m.weight, mode=mode, nonlinearity="relu")
            elif isinstance(m, _BatchNorm):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def conv_layers(self):
        """[(block, conv, bn)] x 18 in forward order."""
        out = [("inc",) + l for l in self.inc.layers()]
        for name, _, _ in _ENC:
            out += [(name,) + l for l in getattr(self, name).maxpool_conv["1"].layers()]
        for name, _, _ in _DEC:
            out += [(name,) + l for l in getattr(self, name).conv.layers()]
        return out

    def up_layers(self):
        return [getattr(self, name).up for name, _, _ in _DEC]

    def forward(self, x):
        eng = _Engine.for_unets([self], x, mode=getattr(self, "_mode", "bf16"), use_tc=getattr(self, "_use_tc", True))
        with torch.cuda.device(x.device):
            rec = eng.forward_unet(x_is_twin=False, save=False, training_stats=self.training)
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        return _nchw_view(rec.cat0, B, H, W, 0, 64), _nchw_view(rec.Hf, B, H, W, 0, 64)


def _nchw_view(buf, B, H, W, coff, C, n0=0):
    """(B,C,H,W)-shaped strided view of channels [coff, coff+C) of an NHWC buffer, images [n0, n0+B)."""
    return buf[n0:n0 + B, :, :, coff:coff + C].permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------
# engine: orchestrates the C-ABI kernels for a (possibly twin) batch
# ------------------------------------------------------------------------------------------------------
class _Rec:
    """Everything one forward leaves behind for the head / backward."""
    pass


class _Segment:
    """A run of images sharing one set of U-Net weights: the whole twin batch for the weight-shared Onet
    (2 BatchNorm groups), or one branch each for bshare=False (1 group)."""

    def __init__(self, unet, n0, n, groups):
        self.unet, self.n0, self.n, self.groups = unet, n0, n, groups
        self.group_images = n // groups


_EVAL_AFFINE = weakref.WeakKeyDictionary()      # BatchNorm module -> {(groups, device): (versions, affine, stream, event)}, see _eval_affine


class _Engine:
    def __init__(self, segments, x, mode, use_tc):
        if not x.is_cuda:
            raise _lib.OnetLibError("onet_b200 has no CPU path: move the module and its input to a CUDA device")
        if mode not in _DTYPES:
            raise ValueError(f"mode must be one of {list(_DTYPES)}")
        self.mode = mode
        self.dt, self.tdt = _DTYPES[mode]
        self.use_tc = bool(use_tc) and mode in ("bf16", "tf32")
        self.segments = segments
        self.x = x
        self.dev = x.device
        self._main = torch.cuda.current_stream(self.dev)
        self.stream = self._main.cuda_stream
        self._side = None

    @staticmethod
    def for_unets(unets, x, mode, use_tc, twin=False):
        B = x.shape[0]
        if twin and len(unets) == 1:
            segs = [_Segment(unets[0], 0, 2 * B, 2)]
        elif twin:
            segs = [_Segment(unets[0], 0, B, 1), _Segment(unets[1], B, B, 1)]
        else:
            segs = [_Segment(unets[0], 0, B, 1)]
        return _Engine(segs, x, mode, use_tc)

    # -------------------------------------------------------------------------------- helpers
    def _engine_for(self, cin, cout):
        return ENGINE_TC if (self.use_tc and cin % 64 == 0 and cout % 64 == 0) else ENGINE_SIMT

    def _empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.tdt, device=self.dev)

    def _pack_key(self, mod, kind):
        w = mod.weight
        return (w.data_ptr(), w._version, _PACK_GEN[0], self.dt, kind)

    def _pack_all(self):
        """Refresh the packed operand copies of every stale conv / up-conv weight in ONE launch (per 24 layers)."""
        import ctypes
        jobs = []
        for seg in self.segments:
            mods = [(conv, "conv") for _, conv, _ in seg.unet.conv_layers()]
            if self.use_tc:
                mods += [(up, "convT") for up in seg.unet.up_layers()]
            for mod, kind in mods:
                key = self._pack_key(mod, kind)
                cache = mod.__dict__.get("_onet_pack")
                if cache is not None and cache[0] == key:
                    continue
                w = mod.weight
                d0, d1 = w.shape[0], w.shape[1]
                taps = 9 if kind == "conv" else 4
                if cache is not None and cache[1].dtype == self.tdt and cache[1].numel() == w.numel() and cache[1].device == w.device:
                    wf, wd = cache[1], cache[2]                      # reuse the buffers, only the contents are stale
                elif kind == "conv":
                    wf, wd = self._empty(d0, 9, d1), self._empty(d1, 9, d0)
                else:
                    wf, wd = self._empty(4 * d1, d0), self._empty(d0, 4 * d1)
                mod.__dict__["_onet_pack"] = (key, wf, wd)
                jobs.append((ptr(w), d0, d1, taps, ptr(wf), ptr(wd)))
        for i in range(0, len(jobs), 24):
            chunk = jobs[i:i + 24]
            n = len(chunk)
            vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
            call("onet_pack_all_weights", n, vp(*[c[0] for c in chunk]), ip(*[c[1] for c in chunk]), ip(*[c[2] for c in chunk]),
                 ip(*[c[3] for c in chunk]), vp(*[c[4] for c in chunk]), vp(*[c[5] for c in chunk]), self.dt, self.stream)

    def _packed(self, mod, kind):
        """Packed operand copies of a conv / up-conv weight, refreshed when the parameter changed."""
        cache = mod.__dict__.get("_onet_pack")
        if cache is not None and cache[0] == self._pack_key(mod, kind):
            return cache[1], cache[2]
        w = mod.weight
        if kind == "conv":
            co, ci = w.shape[0], w.shape[1]
            wf = self._empty(co, 9, ci)
            wd = self._empty(ci, 9, co)
            call("onet_pack_conv_weights", ptr(w), co, ci, ptr(wf), ptr(wd), self.dt, self.stream)
        else:
            ci, co = w.shape[0], w.shape[1]
            wf = self._empty(4 * co, ci)
            wd = self._empty(ci, 4 * co)
            call("onet_pack_convT_weights", ptr(w), ci, co, ptr(wf), ptr(wd), self.dt, self.stream)
        mod.__dict__["_onet_pack"] = (self._pack_key(mod, kind), wf, wd)
        return wf, wd

    # -------------------------------------------------------------------------------- forward
    def forward_unet(self, x_is_twin, save, training_stats, bias=0.0):
        """Runs every segment's U-Net.  training_stats: BatchNorm uses batch statistics (and updates the running
        buffers); save: keep what backward needs."""
        x = self.x
        B, Cin, H, W = x.shape
        if H < 16 or W < 16:
            raise ValueError("Onet needs at least 16 x 16 pixels (four 2x2 poolings)")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.contiguous().float()
        N2 = 2 * B if x_is_twin else B
        rec = _Rec()
        rec.B, rec.H, rec.W, rec.Cin, rec.N2 = B, H, W, Cin, N2
        rec.xin = self._empty(N2, H, W, Cin)
        if x_is_twin:
            call("onet_prep_input", ptr(x), B, Cin, H, W, float(bias), ptr(rec.xin), self.dt, self.stream)
        else:
            rec.xin.copy_(x.permute(0, 2, 3, 1))
        hs = [H >> k for k in range(5)]        # nn.MaxPool2d(2) floors odd sizes; the Up blocks pad back (F.pad, :92-96)
        ws = [W >> k for k in range(5)]
        cs = [64, 128, 256, 512, 1024]
        rec.hs, rec.ws = hs, ws
        # concat buffers: level k holds [skip (C_k) | up-conv output (C_k)]
        rec.cat = [self._empty(N2, hs[k], ws[k], 2 * cs[k]) for k in range(4)]
        rec.cat0 = rec.cat[0]
        rec.pool = [self._empty(N2, hs[k + 1], ws[k + 1], cs[k]) for k in range(4)]
        rec.x5 = self._empty(N2, hs[4], ws[4], 1024)
        rec.mid = {}      # dense activations after the first conv of each block, and decoder block outputs
        rec.saved = {}    # per (segment index, layer index): dict(Y, stats)
        rec.training_stats, rec.save = training_stats, save
        # head fusion (training forward of the twin network): see _conv_bn_relu / head_forward
        rec.y_last, rec.head_aff = None, []
        if x_is_twin and save and os.environ.get("ONET_NO_HEAD_FUSION") is None:
            rec.y_last = self._empty(N2, H, W, 64)
        nseg_groups = sum(seg.groups for seg in self.segments)
        rec.stat_pool = torch.zeros(2 * 5888 * nseg_groups, dtype=torch.float64, device=self.dev) if training_stats else None
        rec.stat_off = 0
        rec.nbt = []
        self._pack_all()

        for si, seg in enumerate(self.segments):
            layers = seg.unet.conv_layers()
            ups = seg.unet.up_layers()
            n0, n = seg.n0, seg.n
            li = 0

            def run(li, src, ld_src, off_src, h, w, dst, ld_dst, off_dst, pool):
                _, conv, bn = layers[li]
                self._conv_bn_relu(rec, si, seg, li, conv, bn, src, ld_src, off_src, n0, n, h, w, dst, ld_dst, off_dst, pool)

            # encoder
            src, ld_src = rec.xin, Cin
            for k in range(5):
                c = cs[k]
                mid = rec.mid.setdefault(("enc", k), self._empty(N2, hs[k], ws[k], c))
                run(li, src, ld_src, 0, hs[k], ws[k], mid, c, 0, None)
                if k < 4:
                    run(li + 1, mid, c, 0, hs[k], ws[k], rec.cat[k], 2 * c, 0, rec.pool[k])
                    src, ld_src = rec.pool[k], c
                else:
                    run(li + 1, mid, c, 0, hs[k], ws[k], rec.x5, c, 0, None)
                li += 2
            # decoder
            below, c_below = rec.x5, 1024
            for j in range(4):
                k = 3 - j                      # level of the skip connection
                c = cs[k]
                up = ups[j]
                self._upconv(seg, up, below, c_below, n0, n, hs[k + 1], ws[k + 1], rec.cat[k], 2 * c, c, hs[k], ws[k])
                mid = rec.mid.setdefault(("dec", k), self._empty(N2, hs[k], ws[k], c))
                run(li, rec.cat[k], 2 * c, 0, hs[k], ws[k], mid, c, 0, None)
                out = rec.mid.setdefault(("dec_out", k), self._empty(N2, hs[k], ws[k], c))
                run(li + 1, mid, c, 0, hs[k], ws[k], out, c, 0, None)
                below, c_below = out, c
                li += 2
        rec.Hf = rec.mid[("dec_out", 0)]
        if rec.nbt:
            torch._foreach_add_([t for t, _ in rec.nbt], [g for _, g in rec.nbt])
        return rec

    def _img_off(self, t, n0):
        return n0 * t.stride(0)

    def _eval_affine(self, bn, G, cout, save, st):
        """(mean, invstd, scale, shift) [4, G, cout] of an eval-mode BatchNorm (running statistics).  Pure inference (nothing saved
        for backward, no graph capture): cached per module until one of its four tensors is modified in place or replaced - a
        forward pass over a 2048 x 2048 frame launched 18 of these 7 us kernels, 1.3 % of its time."""
        cacheable = not save and not torch.cuda.is_current_stream_capturing()
        srcs = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        if cacheable:
            key = (G, self.dev.index)
            ver = tuple(t._version for t in srcs) + tuple(t.data_ptr() for t in srcs) + (_PACK_GEN[0], _STATS_GEN[0])
            cache = _EVAL_AFFINE.setdefault(bn, {})
            hit = cache.get(key)
            if hit is not None and hit[0] == ver:
                if hit[2] != st:                                 # computed on another stream: order this stream behind it, and
                    self._main.wait_event(hit[3])                # keep the allocator from recycling the buffer under this stream
                    hit[1].record_stream(self._main)
                return hit[1]
        aff = torch.empty(4, G, cout, dtype=torch.float32, device=self.dev)
        call("onet_bn_eval_prepare", G, cout, ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
             ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), ptr(aff[2]), ptr(aff[3]), st)
        if save:
            # differentiable eval-mode forward (frozen-BatchNorm fine-tuning, input saliency): the backward kernels read the
            # running statistics where they read the batch statistics in training mode (a handful of C-element host-side ops)
            aff[0] = bn.running_mean
            aff[1] = torch.rsqrt(bn.running_var + 1e-5)
        if cacheable:
            ev = torch.cuda.Event()
            ev.record(self._main)
            cache[key] = (ver, aff, st, ev)
        return aff

    def _conv_bn_relu(self, rec, si, seg, li, conv, bn, src, ld_src, off_src, n0, n, h, w, dst, ld_dst, off_dst, pool):
        cout, cin = conv.weight.shape[0], conv.weight.shape[1]
        wf, _ = self._packed(conv, "conv")
        eng = self._engine_for(cin, cout)
        G = seg.groups
        st = self.stream
        if (li == 0 and cin in (1, 3) and cout == 64 and w % 4 == 0 and pool is None and ld_src == cin and off_src == 0
                and ld_dst == 64 and off_dst == 0 and os.environ.get("ONET_NO_FIRST_RECOMPUTE") is None):
            return self._first_conv_bn_relu(rec, si, seg, li, conv, bn, wf, src, n0, n, h, w, cin, dst)
        fused_eval = (not rec.training_stats) and eng == ENGINE_TC and not rec.save
        # last layer of the twin forward with backward to follow: its BatchNorm + ReLU is applied by the head kernel, and its
        # raw output lives in ONE [2B,H,W,64] buffer across the segments (the head indexes both branches)
        head_fused = getattr(rec, "y_last", None) is not None and li == 17
        if head_fused:
            Y = rec.y_last[n0:n0 + n]
        else:
            Y = None if fused_eval else self._empty(n, h, w, cout)
        if rec.training_stats:
            stats = rec.stat_pool[rec.stat_off:rec.stat_off + 2 * G * cout].view(2, G, cout)
            rec.stat_off += 2 * G * cout
            call("onet_conv3x3_fwd", ptr(src, self._img_off(src, n0) + off_src), ld_src, 0, n, h, w, cin, ptr(wf), cout,
                 ptr(Y), cout, 0, ptr(stats[0]), ptr(stats[1]), seg.group_images, self.dt, eng, st)
            aff = torch.empty(4, G, cout, dtype=torch.float32, device=self.dev)   # mean, invstd, scale, shift
            count = float(seg.group_images * h * w)
            call("onet_bn_finalize", ptr(stats[0]), ptr(stats[1]), G, cout, count,
                 ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                 ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                 float(bn.momentum), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), st)
            rec.nbt.append((bn.num_batches_tracked, G))
            _STATS_GEN[0] += 1
        else:
            aff = self._eval_affine(bn, G, cout, rec.save, st)
            if fused_eval:
                # inference: BatchNorm(eval) + ReLU folded into the conv epilogue, written straight to its destination
                call("onet_conv3x3_bn_relu_infer", ptr(src, self._img_off(src, n0) + off_src), ld_src, 0, n, h, w, cin, ptr(wf),
                     cout, ptr(aff[2]), ptr(aff[3]), seg.group_images, ptr(dst, self._img_off(dst, n0) + off_dst), ld_dst, 0,
                     self.dt, eng, st)
                if pool is not None:
                    call("onet_maxpool2x2", ptr(dst, self._img_off(dst, n0) + off_dst), ld_dst, 0, n, h, w, cout,
                         ptr(pool, self._img_off(pool, n0)), self.dt, st)
                return
            call("onet_conv3x3_fwd", ptr(src, self._img_off(src, n0) + off_src), ld_src, 0, n, h, w, cin, ptr(wf), cout,
                 ptr(Y), cout, 0, None, None, seg.group_images, self.dt, eng, st)
        # with a pooled output kept for backward: the window arg-max (2 bits per channel) is stored, not recomputed there
        parg = torch.empty(n, h // 2, w // 2, cout // 8, dtype=torch.int16, device=self.dev) if (pool is not None and rec.save) else None
        if head_fused:
            rec.head_aff.append((n0, n, aff))
        else:
            call("onet_bn_relu_apply", ptr(Y), n, h, w, cout, ptr(aff[2]), ptr(aff[3]), seg.group_images,
                 ptr(dst, self._img_off(dst, n0) + off_dst), ld_dst, 0,
                 ptr(pool, self._img_off(pool, n0)) if pool is not None else None, ptr(parg), self.dt, st)
        if rec.save:
            rec.saved[(si, li)] = dict(Y=Y, aff=aff, src=(src, ld_src, off_src), h=h, w=w, cin=cin, cout=cout, parg=parg)

    def _first_conv_bn_relu(self, rec, si, seg, li, conv, bn, wf, src, n0, n, h, w, cin, dst):
        """First convolution of a U-Net (reference :111,47-49) WITHOUT materialising its 64-channel raw output: the statistics
        pass and the BatchNorm + ReLU pass both recompute it from the input patch (csrc/first_layer.cuh); backward does the same
        (`onet_first_conv_bwd`).  Same results as the generic conv -> BatchNorm path, 3.2 GB less HBM traffic per step at B = 64."""
        G, st = seg.groups, self.stream
        x = ptr(src, self._img_off(src, n0))
        aff = torch.empty(4, G, 64, dtype=torch.float32, device=self.dev)   # mean, invstd, scale, shift
        # The conv output is linear in the 3 x 3 x in_chns patch (K = 9 in_chns elements), so the patch moments (S[K], G[K][K] per
        # statistics group) give the BatchNorm statistics in closed form and, in backward, the y-dependent part of the weight
        # gradient (csrc/first_layer.cuh); y is then never rounded.  in_chns = 1: every storage type; in_chns = 3: bf16 storage
        # (the moments come from warp-level MMAs, csrc/first_layer_mma.cuh).  ONET_NO_FIRST_GRAM=1: the two-pass form.
        K = 9 * cin
        use_gram = ((cin == 1 or (cin == 3 and self.mode == "bf16" and os.environ.get("ONET_NO_FIRST_MMA") is None))
                    and os.environ.get("ONET_NO_FIRST_GRAM") is None)
        gram = torch.zeros(G, K + K * K, dtype=torch.float64, device=self.dev) if use_gram and (rec.training_stats or rec.save) else None
        if rec.training_stats:
            stats = rec.stat_pool[rec.stat_off:rec.stat_off + 2 * G * 64].view(2, G, 64)
            rec.stat_off += 2 * G * 64
            call("onet_first_conv_stats", x, n, h, w, cin, ptr(wf), ptr(gram), ptr(stats[0]), ptr(stats[1]), seg.group_images,
                 self.dt, st)
            call("onet_bn_finalize", ptr(stats[0]), ptr(stats[1]), G, 64, float(seg.group_images * h * w),
                 ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                 ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
                 float(bn.momentum), ptr(aff[0]), ptr(aff[1]), ptr(aff[2]), ptr(aff[3]), st)
            rec.nbt.append((bn.num_batches_tracked, G))
            _STATS_GEN[0] += 1
        else:
            aff = self._eval_affine(bn, G, 64, rec.save, st)
        if gram is not None and not rec.training_stats:       # eval-mode forward with backward to follow: moments only
            scratch = torch.empty(2, G, 64, dtype=torch.float64, device=self.dev)
            call("onet_first_conv_stats", x, n, h, w, cin, ptr(wf), ptr(gram), ptr(scratch[0]), ptr(scratch[1]), seg.group_images,
                 self.dt, st)
        round_y = 0 if use_gram else 1
        call("onet_first_conv_bn_relu", x, n, h, w, cin, ptr(wf), ptr(aff[2]), ptr(aff[3]), seg.group_images,
             ptr(dst, self._img_off(dst, n0)), round_y, self.dt, st)
        if rec.save:
            rec.saved[(si, li)] = dict(Y=None, aff=aff, src=(src, cin, 0), h=h, w=w, cin=cin, cout=64, first=True, wf=wf, gram=gram)

    def _upconv(self, seg, up, x, cx, n0, n, h, w, cat, ld_cat, off_cat, ho, wo):
        """x [n,h,w,cin] -> channels [off_cat, off_cat+co) of the concat buffer [n,ho,wo,ld_cat]; ho - 2h, wo - 2w in {0,1}:
        the reference's F.pad (:92-96) puts the up-sampled map at offset (0,0) and zero-fills the last row / column."""
        cin, co = up.weight.shape[0], up.weight.shape[1]
        eng = self._engine_for(cin, co)
        if eng == ENGINE_TC:
            wf, _ = self._packed(up, "convT")
            wptr = ptr(wf)
        else:
            wptr = ptr(up.weight)
        dst = ptr(cat, self._img_off(cat, n0) + off_cat)
        if ho != 2 * h or wo != 2 * w:
            call("onet_zero_border", dst, n, ho, wo, ld_cat, 0, co, 2 * h, 2 * w, self.dt, self.stream)
        call("onet_convT2x2_fwd", ptr(x, self._img_off(x, n0)), cx, 0, n, h, w, cin, wptr, ptr(up.bias), co,
             dst, ld_cat, 0, ho, wo, self.dt, eng, self.stream)

    # -------------------------------------------------------------------------------- head
    def head_forward(self, rec):
        B, H, W = rec.B, rec.H, rec.W
        f32 = torch.float32
        rec.Vt = torch.empty(B, 1, H, W, dtype=f32, device=self.dev)
        rec.Vd = torch.empty(B, 1, H, W, dtype=f32, device=self.dev)
        rec.S = torch.empty(B, 2, H, W, dtype=f32, device=self.dev)
        rec.a = torch.empty(B, H, W, dtype=f32, device=self.dev)
        rec.b = torch.empty(B, H, W, dtype=f32, device=self.dev)
        rec.loss_acc = torch.zeros((), dtype=torch.float64, device=self.dev)
        if rec.y_last is not None:
            # the global feature H = relu(bn(y_last)) is formed inside the head kernel (per-branch BatchNorm constants)
            st_, sh_ = {}, {}
            for n0, n, aff in rec.head_aff:           # shared twin: one segment, groups 0 / 1; otherwise one segment per branch
                for gi in range(aff.shape[1]):
                    st_[n0 + gi * (n // aff.shape[1])] = aff[2][gi]
                    sh_[n0 + gi * (n // aff.shape[1])] = aff[3][gi]
            call("onet_head_fwd_bn", ptr(rec.cat0), 128, 0, ptr(rec.y_last), 64, 0, B, H, W, ptr(st_[0]), ptr(sh_[0]), ptr(st_[B]),
                 ptr(sh_[B]), ptr(rec.Vt), ptr(rec.Vd), ptr(rec.S), ptr(rec.a), ptr(rec.b), ptr(rec.loss_acc), self.dt, self.stream)
            return
        call("onet_head_fwd", ptr(rec.cat0), 128, 0, ptr(rec.Hf), 64, 0, B, H, W, ptr(rec.Vt), ptr(rec.Vd), ptr(rec.S),
             ptr(rec.a), ptr(rec.b), ptr(rec.loss_acc), self.dt, self.stream)

    # -------------------------------------------------------------------------------- backward
    def backward(self, rec, gscale, gVt, gVd, gS, gLt, gLd, grad_of, after_block=None):
        """Full backward of head + both U-Nets.  `grad_of(param)` returns the fp32 tensor the parameter's
        gradient must be ACCUMULATED into."""
        B, H, W, N2 = rec.B, rec.H, rec.W, rec.N2
        st = self.stream
        if self.mode in ("fp32", "tf32"):       # (tf32: the first layer's weight gradient runs on the CUDA-core kernels)
            _register_splitk_workspace(self.dev)
        # per-call event timing (bench.py's profile pass) records on the main stream only: keep that pass serial
        overlap = os.environ.get("ONET_NO_WGRAD_OVERLAP") is None and _lib.PROFILE is None
        self._side = _side_stream(self.dev) if overlap else None
        self._deferred, self._inflight, self._hooks_deferred, self._hooks_ready = [], [], [], []
        self._after_block = after_block
        dL = self._empty(N2, H, W, 64)
        head_fused = rec.y_last is not None
        if head_fused:
            # only the per-pixel part of the head backward runs here; dH = dV * L is formed on the fly by the last layer's
            # BatchNorm backward, whose reduce pass also writes dL (onet_bn_relu_bwd_head)
            dH = None
            gv = torch.empty(N2 * H * W, dtype=torch.float32, device=self.dev)
            gab = torch.empty(N2 * H * W, dtype=torch.float32, device=self.dev)
            call("onet_head_bwd_scalars", ptr(rec.Vt), ptr(rec.Vd), ptr(rec.a), ptr(rec.b), ptr(gscale), ptr(gVt), ptr(gVd), ptr(gS),
                 B, H, W, ptr(gv), ptr(gab), st)
        else:
            dH = self._empty(N2, H, W, 64)
            call("onet_head_bwd", ptr(rec.cat0), 128, 0, ptr(rec.Hf), 64, 0, B, H, W, ptr(rec.Vt), ptr(rec.Vd), ptr(rec.a),
                 ptr(rec.b), ptr(gscale), ptr(gVt), ptr(gVd), ptr(gS), ptr(dL), ptr(dH), self.dt, st)

        def add_external(lo, hi):   # external gradient w.r.t. the returned local features (generic autograd path)
            for g_ext, m0 in ((gLt, 0), (gLd, B)):
                if g_ext is not None and lo <= m0 and m0 + B <= hi:
                    dL[m0:m0 + B].add_(g_ext.permute(0, 2, 3, 1).to(self.tdt))
        if not head_fused:
            add_external(0, N2)
        cs = [64, 128, 256, 512, 1024]
        hs, ws = rec.hs, rec.ws
        rec.sum_pool = torch.zeros(2 * 5888 * sum(seg.groups for seg in self.segments), dtype=torch.float64, device=self.dev)
        rec.sum_off = 0
        for si, seg in enumerate(self.segments):
            layers = seg.unet.conv_layers()
            ups = seg.unet.up_layers()
            n0, n = seg.n0, seg.n

            def bwd(li, g1, ld1, off1, g2=None, ld2=0, off2=0, gp=None, need_dgrad=True, colsum=None, head=None):
                return self._conv_bn_relu_bwd(rec, si, seg, li, layers[li][1], layers[li][2], n0, n, g1, ld1, off1, g2,
                                              ld2, off2, gp, need_dgrad, grad_of, colsum, head)

            # decoder, top (level 0) to bottom (level 3): layer indices 10+2j, 11+2j for j = 0..3 (k = 3-j)
            g_out = None if head_fused else dH[n0:n0 + n]            # gradient w.r.t. the block output at level k
            dcat = [None] * 4
            for k in range(4):
                j = 3 - k
                li = 10 + 2 * j
                c = cs[k]
                if k == 0 and head_fused:
                    px0 = n0 * H * W
                    d_mid = bwd(li + 1, rec.cat0[n0:n0 + n], 128, 0, head=(gv[px0:], gab[px0:], dL[n0:n0 + n]))
                    add_external(n0, n0 + n)
                else:
                    d_mid = bwd(li + 1, g_out, c, 0)                   # -> grad wrt mid activation [n,h,w,c]
                # grad wrt concat buffer [n,h,w,2c]; its column sums (up half = the up-conv's bias gradient) come
                # out of the same kernel's epilogue
                colsum = torch.zeros(2, 2 * c, dtype=torch.float64, device=self.dev)
                dcat[k] = bwd(li, d_mid, c, 0, colsum=colsum)
                # transposed conv: go = up half of dcat
                up = ups[j]
                below = rec.x5 if k == 3 else rec.mid[("dec_out", k + 1)]
                g_out = self._upconv_bwd(seg, up, below, 2 * c, n0, n, hs[k + 1], ws[k + 1], dcat[k], 2 * c, c, grad_of,
                                         colsum[0, c:], hs[k], ws[k])
                self._block_done(seg.unet, _DEC[j][0])
            # encoder, bottom (level 4) to top
            d_mid = bwd(9, g_out, 1024, 0)
            d_pool = bwd(8, d_mid, 1024, 0)                             # grad wrt pool[3]
            self._block_done(seg.unet, "down4")
            for k in (3, 2, 1, 0):
                c = cs[k]
                li = 2 * k
                if k == 0:
                    d_mid = bwd(li + 1, dcat[k], 2 * c, 0, g2=dL[n0:n0 + n], ld2=64, off2=0, gp=d_pool)
                    bwd(li, d_mid, c, 0, need_dgrad=False)
                else:
                    d_mid = bwd(li + 1, dcat[k], 2 * c, 0, gp=d_pool)
                    d_pool = bwd(li, d_mid, c, 0)
                self._block_done(seg.unet, "inc" if k == 0 else _ENC[k - 1][0])
        self._flush_side()
        self._join_side()

    # ---- two-stream schedule of backward: weight gradients next to the following BatchNorm backward
    def _wgrad(self, keep, name, *args):
        """Issue a weight-gradient call (its last argument, the stream, is appended here).  With the side stream it is
        deferred until the next `_flush_side`, i.e. until just before the next HBM-bound kernel goes to the main stream."""
        if self._side is None:
            call(name, *args, self.stream)
        else:
            self._deferred.append((keep, name, args))

    def _flush_side(self):
        """Everything issued on the main stream so far is a dependency of the deferred weight gradients: fork."""
        if self._side is None or not (self._deferred or self._hooks_deferred):
            return
        if self._deferred:
            self._side.wait_stream(self._main)
            for keep, name, args in self._deferred:
                call(name, *args, self._side.cuda_stream)
                self._inflight.append(keep)          # operands stay alive until the main stream has joined
            self._deferred = []
        self._hooks_ready += self._hooks_deferred
        self._hooks_deferred = []

    def _join_side(self):
        """Main stream waits for the side stream: the next tensor-core kernel runs alone, operand buffers may be reused,
        and the finished blocks' gradient buckets can go to the all-reduce."""
        if self._side is None:
            return
        if self._inflight:
            self._main.wait_stream(self._side)
            self._inflight = []
        for unet, block in self._hooks_ready:
            self._after_block(unet, block)
        self._hooks_ready = []

    def _block_done(self, unet, block):
        if self._after_block is None:
            return
        if self._side is None:
            self._after_block(unet, block)
        else:
            self._hooks_deferred.append((unet, block))

    def _conv_bn_relu_bwd(self, rec, si, seg, li, conv, bn, n0, n, g1, ld1, off1, g2, ld2, off2, gp, need_dgrad, grad_of,
                          colsum=None, head=None):
        """g1/g2/gp are SEGMENT-LOCAL tensors (first image = image n0 of the batch)."""
        sv = rec.saved[(si, li)]
        Y, aff, h, w, cin, cout = sv["Y"], sv["aff"], sv["h"], sv["w"], sv["cin"], sv["cout"]
        G = seg.groups
        st = self.stream
        if sv.get("first"):
            # first convolution: BatchNorm backward + weight gradient with y and dY recomputed / kept in registers
            assert g2 is None and gp is None and ld1 == 64 and off1 == 0 and not need_dgrad
            sums = rec.sum_pool[rec.sum_off:rec.sum_off + 2 * G * 64]
            rec.sum_off += 2 * G * 64
            count = float(seg.group_images * h * w) if rec.training_stats else float("inf")
            src = sv["src"][0]
            # the last call of a U-Net's backward: queued behind the deferred weight gradients on the same (side) stream - it
            # shares the deterministic split-K workspace of the FP32 mode with them, which must be used from ONE stream
            gram = sv.get("gram")
            acc_a = torch.zeros(G, 64, 9 * cin, dtype=torch.float32, device=self.dev) if gram is not None else None
            self._wgrad((g1, src, sums, acc_a), "onet_first_conv_bwd", ptr(src, self._img_off(src, n0)), n, h, w, cin, ptr(sv["wf"]),
                        ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]), seg.group_images, ptr(g1), ptr(gram), ptr(acc_a),
                        ptr(sums), count,
                        ptr(grad_of(conv.weight)), ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), ptr(grad_of(bn.weight)),
                        ptr(grad_of(bn.bias)), self.dt)
            return None
        if sv.get("prered") is None:
            sums = rec.sum_pool[rec.sum_off:rec.sum_off + 2 * G * cout]
            rec.sum_off += 2 * G * cout
        dY = self._empty(n, h, w, cout)
        # eval-mode statistics are constants: the mean / projection terms of the BatchNorm backward vanish (1 / count = 0)
        count = float(seg.group_images * h * w) if rec.training_stats else float("inf")
        self._flush_side()          # the previous layers' weight gradients run next to this BatchNorm backward
        if head is not None:
            # last layer, fused with the head backward: g1 is the local feature L, the gradient is dV * L (head = (dV, d(a), dL))
            assert g2 is None and gp is None
            gv, gab, dL = head
            call("onet_bn_relu_bwd_head", ptr(Y), n, h, w, cout, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]),
                 seg.group_images, ptr(g1, off1), ld1, 0, ptr(gv), ptr(gab), ptr(dL), ptr(sums), count, ptr(dY),
                 ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), self.dt, st)
        elif sv.get("prered") is not None:
            # the dgrad launch that produced g1 already reduced this layer's sums in its epilogue: apply pass only
            assert g2 is None and gp is None and ld1 == cout and off1 == 0
            call("onet_bn_relu_bwd_apply", ptr(Y), n, h, w, cout, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]),
                 seg.group_images, ptr(g1), ld1, 0, ptr(sv["prered"]), count, ptr(dY),
                 ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), self.dt, st)
        else:
            call("onet_bn_relu_bwd", ptr(Y), n, h, w, cout, ptr(aff[2]), ptr(aff[3]), ptr(aff[0]), ptr(aff[1]),
                 seg.group_images, ptr(g1, off1), ld1, 0, ptr(g2, off2) if g2 is not None else None, ld2, 0,
                 ptr(gp) if gp is not None else None, ptr(sv.get("parg")) if gp is not None else None, ptr(sums), count, ptr(dY),
                 ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), ptr(grad_of(bn.weight)), ptr(grad_of(bn.bias)), self.dt, st)
        self._join_side()
        eng = self._engine_for(cin, cout)
        src, ld_src, off_src = sv["src"]
        self._wgrad((dY, src), "onet_conv3x3_wgrad", ptr(dY), cout, 0, ptr(src, self._img_off(src, n0) + off_src), ld_src, 0,
                    n, h, w, cin, cout, ptr(grad_of(conv.weight)), self.dt, eng)
        if not need_dgrad:
            return None
        _, wd = self._packed(conv, "conv")
        dX = self._empty(n, h, w, cin)
        prev = rec.saved.get((si, li - 1)) if (li & 1) else None
        if (prev is not None and eng == ENGINE_TC and self.mode == "bf16" and colsum is None and prev["cout"] == cin and prev["h"] == h
                and prev.get("Y") is not None and cin >= _BNRED_MIN_C and os.environ.get("ONET_NO_BNRED_FUSION") is None):
            # second conv of a DoubleConv: its data gradient is the gradient w.r.t. the first conv's activation - reduce the
            # first conv's BatchNorm-backward sums in this launch's epilogue (saves one pass over (Y, g) in HBM)
            paff = prev["aff"]
            G = seg.groups
            psums = rec.sum_pool[rec.sum_off:rec.sum_off + 2 * G * cin]
            rec.sum_off += 2 * G * cin
            call("onet_conv3x3_dgrad_bnred", ptr(dY), cout, 0, n, h, w, cout, ptr(wd), cin, ptr(dX), ptr(prev["Y"]),
                 ptr(paff[2]), ptr(paff[3]), ptr(paff[0]), ptr(paff[1]), ptr(psums), seg.group_images, self.dt, eng, st)
            prev["prered"] = psums
            return dX
        call("onet_conv3x3_fwd", ptr(dY), cout, 0, n, h, w, cout, ptr(wd), cin, ptr(dX), cin, 0,
             ptr(colsum[0]) if colsum is not None else None,
             ptr(colsum[1]) if (colsum is not None and eng != ENGINE_TC) else None,     # tcgen05 epilogue: sums only
             n if colsum is not None else seg.group_images, self.dt, eng, st)
        return dX

    def _upconv_bwd(self, seg, up, x, ld_go_total, n0, n, h, w, dcat, ld_cat, off_cat, grad_of, bias_sums, ho, wo):
        cin, co = up.weight.shape[0], up.weight.shape[1]
        eng = self._engine_for(cin, co)
        st = self.stream
        padded = ho != 2 * h or wo != 2 * w
        if not padded:       # bias gradient = column sums of d(concat)'s up half, already reduced by the dgrad epilogue
            call("onet_add_colsums", ptr(bias_sums), co, ptr(grad_of(up.bias)), st)
        # with an F.pad border the gradient of the border pixels is dropped: reduce over the valid window only
        self._wgrad((x, dcat), "onet_convT2x2_wgrad", ptr(x, self._img_off(x, n0)), cin, 0, ptr(dcat, off_cat), ld_cat, 0, n, h, w,
                    cin, co, ptr(grad_of(up.weight)), ptr(grad_of(up.bias)) if padded else None, ho, wo, self.dt, eng)
        dX = self._empty(n, h, w, cin)
        if eng == ENGINE_TC:
            _, wd = self._packed(up, "convT")
            wptr = ptr(wd)
        else:
            wptr = ptr(up.weight)
        call("onet_convT2x2_dgrad", ptr(dcat, off_cat), ld_cat, 0, n, h, w, cin, wptr, co, ptr(dX), cin, 0, ho, wo, self.dt, eng, st)
        return dX


# ------------------------------------------------------------------------------------------------------
# autograd glue
# ------------------------------------------------------------------------------------------------------
class _OnetFn(torch.autograd.Function):
    """forward(X) of the twin network as one autograd node.  Outputs (Lt, Vt, Ld, Vd, S, anchor); `anchor` is a
    0-dim tensor whose incoming gradient carries d(loss) of the fused JSD loss (see Onet.compute_loss)."""

    @staticmethod
    def forward(ctx, onet, X, leaf):
        # `leaf` is a throw-away 0-dim tensor that requires grad: it makes the outputs differentiable without putting
        # the parameters' AccumulateGrad nodes (which remember the stream they were created on, and would tie a CUDA-graph
        # capture to uncaptured streams) into the autograd graph.  Parameter gradients are accumulated by the kernels.
        eng, rec = onet._run_forward(X, save=True)
        ctx.onet, ctx.eng, ctx.rec = onet, eng, rec
        ctx.set_materialize_grads(False)
        B, H, W = rec.B, rec.H, rec.W
        Lt = _nchw_view(rec.cat0, B, H, W, 0, 64, 0)
        Ld = _nchw_view(rec.cat0, B, H, W, 0, 64, B)
        anchor = torch.zeros((), dtype=torch.float32, device=X.device)
        onet._fwd_rec = rec
        return Lt, rec.Vt, Ld, rec.Vd, rec.S, anchor

    @staticmethod
    def backward(ctx, gLt, gVt, gLd, gVd, gS, g_anchor):
        onet, eng, rec = ctx.onet, ctx.eng, ctx.rec
        if rec is None:
            raise RuntimeError("onet_b200: backward called twice on the same forward")
        f32 = lambda t: None if t is None else t.contiguous().float()
        with torch.no_grad(), torch.cuda.device(eng.dev):
            grad_of = onet._grad_targets()
            eng.backward(rec, f32(g_anchor), f32(gVt), f32(gVd), f32(gS), gLt, gLd, grad_of,
                         after_block=getattr(onet, "_after_block", None))
        ctx.rec = None
        return None, None, None


class _FusedJsdFn(torch.autograd.Function):
    """compute_loss on the tensors of the last forward: the value was produced by the head kernel already; the
    gradient is handed to _OnetFn.backward through the anchor."""

    @staticmethod
    def forward(ctx, anchor, loss_value):
        return loss_value.clone()

    @staticmethod
    def backward(ctx, g):
        return g, None


class Onet(nn.Module):
    def __init__(self, in_chns=1, binit=False, bshare=True, mode="bf16", use_tc=True):
        """`mode`: "bf16" (bf16 storage/operands, fp32 accumulate - the throughput mode), "tf32" (fp32 storage, tensor-core
        operands read as TF32, fp32 accumulate - the fast mode inside the north_star tolerances) or "fp32" (FP32
        verification mode on CUDA cores).  `use_tc=False` forces the CUDA-core kernels in the bf16 / tf32 modes too."""
        super().__init__()
        self.topu = UNet(n_channels=in_chns, n_classes=1, bilinear=False, binit=binit)
        if bshare:
            self.dwnu = self.topu
        else:
            self.dwnu = UNet(n_channels=in_chns, n_classes=1, bilinear=False, binit=binit)
        self.softmax = nn.Softmax2d()
        self.bias = 0
        self.mode = mode
        self.use_tc = use_tc
        self._last = None
        self._fwd_rec = None
        self._arena = None
        self._after_block = None

    # ------------------------------------------------------------------ parameters / gradients
    def _unets(self):
        return [self.topu] if self.dwnu is self.topu else [self.topu, self.dwnu]

    def flatten_parameters(self):
        """Re-home every parameter (and its gradient) in ONE contiguous fp32 arena so the optimizer step is a
        single kernel and the data-parallel gradient all-reduce a few large buckets.  Idempotent; call after
        `.to(device)`.  Returns (flat_params, flat_grads)."""
        params = list(self.parameters())
        dev = params[0].device
        ar = self._arena
        if ar is not None and ar["flat"].device == dev and all(
                p.data_ptr() == ar["flat"].data_ptr() + o * 4 for p, o in zip(params, ar["offsets"])):
            return ar["flat"], ar["grads"]
        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (p.numel() + 63) // 64 * 64          # keep every tensor 256-byte aligned
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        grads = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offsets):
                flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat[o:o + p.numel()].view(p.shape)
                p.grad = grads[o:o + p.numel()].view(p.shape)
        self._arena = dict(flat=flat, grads=grads, offsets=offsets, total=total)
        return flat, grads

    def _grad_targets(self):
        """Returns grad_of(param) -> tensor to accumulate into; attaches zeroed .grad tensors where missing."""
        ar = self._arena
        if ar is not None:
            params = list(self.parameters())
            if all(p.grad is None for p in params):       # zero_grad(set_to_none=True): re-attach arena views
                ar["grads"].zero_()
                for p, o in zip(params, ar["offsets"]):
                    p.grad = ar["grads"][o:o + p.numel()].view(p.shape)

        def grad_of(p):
            if p.grad is None:
                p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
            elif not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            return p.grad
        return grad_of

    # ------------------------------------------------------------------ forward / loss
    def _run_forward(self, X, save):
        eng = _Engine.for_unets(self._unets(), X, self.mode, self.use_tc, twin=True)
        with torch.cuda.device(X.device):
            rec = eng.forward_unet(x_is_twin=True, save=save, training_stats=self.training, bias=float(self.bias))
            eng.head_forward(rec)
        return eng, rec

    def forward(self, X):
        assert X.dim() == 4
        if torch.is_grad_enabled() and (self.training or any(p.requires_grad for p in self.parameters())):
            # eval mode with autograd on (the reference's outputs are differentiable there too): BatchNorm uses and keeps its
            # running statistics, the backward treats them as constants
            leaf = torch.zeros((), dtype=torch.float32, device=X.device, requires_grad=True)
            Lt, Vt, Ld, Vd, S, anchor = _OnetFn.apply(self, X, leaf)
            self._last = dict(Lt=Lt, Ld=Ld, S=S, anchor=anchor, rec=self._fwd_rec)
            self._fwd_rec = None
            return Lt, Vt, Ld, Vd, S
        eng, rec = self._run_forward(X, save=False)
        B, H, W = rec.B, rec.H, rec.W
        Lt = _nchw_view(rec.cat0, B, H, W, 0, 64, 0)
        Ld = _nchw_view(rec.cat0, B, H, W, 0, 64, B)
        self._last = dict(Lt=Lt, Ld=Ld, S=rec.S, anchor=None, rec=rec)
        return Lt, rec.Vt, Ld, rec.Vd, rec.S

    def _is_last(self, Lt, St, Ld, Sd):
        """True when the arguments are exactly (Lt, S[:,0:1], Ld, S[:,1:2]) of the most recent forward."""
        l = self._last
        if l is None or l["rec"] is None or Lt is not l["Lt"] or Ld is not l["Ld"]:
            return False
        S = l["S"]
        B, _, H, W = S.shape
        hw = H * W
        for t, off in ((St, 0), (Sd, hw)):
            if (tuple(t.shape) != (B, 1, H, W) or t.dtype != S.dtype or t.data_ptr() != S.data_ptr() + 4 * off
                    or t.stride(0) != 2 * hw or t.stride(2) != W or t.stride(3) != 1):
                return False
        return True

    def compute_loss(self, Lt, St, Ld, Sd):
        """-(JSD(Lt,St,Sd) + JSD(Ld,Sd,St)) / 2, reference :253-267.  When the arguments are the tensors of the
        last forward (the way both reference training loops call it) the value and the closed-form gradient come
        from the fused head kernels; otherwise the generic path below is used."""
        if self._is_last(Lt, St, Ld, Sd):
            l = self._last
            rec = l["rec"]
            n = rec.B * rec.H * rec.W
            value = (rec.loss_acc / (2.0 * n)).float()
            if l["anchor"] is not None and l["anchor"].requires_grad:
                return _FusedJsdFn.apply(l["anchor"], value)
            return value
        jsd_top = self.jensen_shannon_divergence(Lt, St, Sd)
        jsd_dwn = self.jensen_shannon_divergence(Ld, Sd, St)
        return -(jsd_top + jsd_dwn) / 2

    def jensen_shannon_divergence(self, Li, Si, Sprime):
        """Generic (unfused) form, reference :221-235, for callers that pass their own tensors."""
        assert (Li.dim() == 4 and Si.dim() == 4 and Sprime.dim() == 4)
        a = Li.float().sum(dim=1)
        jsd = -1 * self.log1pexp(-1 * (a * Si[:, 0])).mean() - self.log1pexp(a * Sprime[:, 0]).mean()
        assert (torch.isnan(jsd) == False)
        return jsd

    def log1pexp(self, x):
        """The reference's piecewise softplus with its exact semantics (:237-251; ln 2 plateau below -37),
        written without in-place mutation."""
        zero = torch.zeros_like(x)
        lo = x <= -37.0
        t = torch.where(lo, torch.exp(torch.where(lo, x, zero)), x)
        mid = (t > -37.0) & (t <= 18.0)
        t = torch.where(mid, torch.log(1 + torch.exp(torch.where(mid, t, zero))), t)
        hi = (t > 18.0) & (t < 33.3)
        return torch.where(hi, t + torch.exp(-torch.where(hi, t, zero)), t)

    def predict_label(self, S):
        """argmax over the two class maps, reference :193-202 (ties -> 0)."""
        with torch.no_grad():
            assert (S.dim() == 4)
            l = self._last
            if l is not None and l["rec"] is not None and S is l["S"]:
                rec = l["rec"]
                Y = torch.empty(rec.B, rec.H, rec.W, dtype=torch.long, device=S.device)
                call("onet_predict_label", ptr(rec.Vt), ptr(rec.Vd), Y.numel(), ptr(Y),
                     torch.cuda.current_stream(S.device).cuda_stream, device=S.device)
                return Y
            return torch.argmax(S, dim=1)

    def get_label(self, Vt, Vd):
        """reference :204-219"""
        with torch.no_grad():
            V = self.softmax(torch.concat([Vt, Vd], dim=1))
            return torch.argmax(V, dim=1), V
