"""Helpers for the -m gpu parity tests: thin wrappers that call the C ABI on torch CUDA tensors."""
import torch

from onet_b200 import _lib
from onet_b200._lib import BF16, ENGINE_SIMT, ENGINE_TC, F32, call, ptr

TDT = {F32: torch.float32, BF16: torch.bfloat16}


def stream():
    return torch.cuda.current_stream().cuda_stream


def to_nhwc(x, dtype):
    """(N,C,H,W) fp32 -> contiguous [N,H,W,C] of dtype"""
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def from_nhwc(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def pack_conv(w, dt):
    co, ci = w.shape[0], w.shape[1]
    wf = torch.empty(co, 9, ci, dtype=TDT[dt], device=w.device)
    wd = torch.empty(ci, 9, co, dtype=TDT[dt], device=w.device)
    call("onet_pack_conv_weights", ptr(w), co, ci, ptr(wf), ptr(wd), dt, stream())
    return wf, wd


def pack_convT(w, dt):
    ci, co = w.shape[0], w.shape[1]
    wf = torch.empty(4 * co, ci, dtype=TDT[dt], device=w.device)
    wd = torch.empty(ci, 4 * co, dtype=TDT[dt], device=w.device)
    call("onet_pack_convT_weights", ptr(w), ci, co, ptr(wf), ptr(wd), dt, stream())
    return wf, wd


def conv3x3(x_nhwc, wpacked, cout, dt, engine, group_images=0, stats=False, ld_in=None, off_in=0, cin=None):
    n, h, w, c_total = x_nhwc.shape
    cin = cin or c_total
    out = torch.empty(n, h, w, cout, dtype=TDT[dt], device=x_nhwc.device)
    st = torch.zeros(2, 2, cout, dtype=torch.float64, device=x_nhwc.device) if stats else None
    call("onet_conv3x3_fwd", ptr(x_nhwc, off_in), ld_in or c_total, 0, n, h, w, cin, ptr(wpacked), cout, ptr(out), cout, 0,
         ptr(st[0]) if stats else None, ptr(st[1]) if stats else None, group_images, dt, engine, stream())
    return out, st


def conv3x3_wgrad(g_nhwc, x_nhwc, dt, engine):
    n, h, w, cout = g_nhwc.shape
    cin = x_nhwc.shape[3]
    dw = torch.zeros(cout, cin, 3, 3, dtype=torch.float32, device=g_nhwc.device)
    call("onet_conv3x3_wgrad", ptr(g_nhwc), cout, 0, ptr(x_nhwc), cin, 0, n, h, w, cin, cout, ptr(dw), dt, engine, stream())
    return dw
