"""ctypes binding of the C-ABI library `libonet_b200.so` (include/onet_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is NO fallback: if the
shared object is missing or a call fails, an exception is raised."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libonet_b200.so")
CSRC = os.path.join(_HERE, "csrc")

F32, BF16 = 0, 1
ENGINE_SIMT, ENGINE_TC = 0, 1

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC"]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(os.path.dirname(_HERE), "include", "onet_b200.h")]


def build(force=False, verbose=False):
    """Compile csrc/capi.cu -> libonet_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and os.path.isfile(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in _sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "capi.cu"), "-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


class OnetLibError(RuntimeError):
    pass


_p, _i, _i64, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> argtypes, exactly the prototypes of include/onet_b200.h
SIGNATURES = {
    "onet_version": [],
    "onet_device_info": [_p, _p, _p],
    "onet_set_splitk_workspace": [_p, _i64],
    "onet_prep_input": [_p, _i, _i, _i, _i, _f, _p, _i, _p],
    "onet_pack_conv_weights": [_p, _i, _i, _p, _p, _i, _p],
    "onet_pack_convT_weights": [_p, _i, _i, _p, _p, _i, _p],
    "onet_pack_all_weights": [_i, _p, _p, _p, _p, _p, _p, _i, _p],
    "onet_conv3x3_fwd": [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _i64, _i, _p, _p, _i, _i, _i, _p],
    "onet_first_conv_stats": [_p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p],
    "onet_first_conv_bn_relu": [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _p],
    "onet_first_conv_bwd": [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _d, _p, _p, _p, _p, _p, _i, _p],
    "onet_conv3x3_bn_relu_infer": [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p, _i64, _i, _i, _i, _p],
    "onet_maxpool2x2": [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p],
    "onet_conv3x3_wgrad": [_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p],
    "onet_bn_finalize": [_p, _p, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p],
    "onet_bn_eval_prepare": [_i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "onet_bn_relu_apply": [_p, _i, _i, _i, _i, _p, _p, _i, _p, _i64, _i, _p, _p, _i, _p],
    "onet_bn_relu_bwd": [_p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _i64, _i, _p, _i64, _i, _p, _p, _p, _d, _p,
                         _p, _p, _p, _p, _i, _p],
    "onet_conv3x3_dgrad_bnred": [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "onet_bn_relu_bwd_apply": [_p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _i64, _i, _p, _d, _p, _p, _p, _p, _p, _i, _p],
    "onet_convT2x2_fwd": [_p, _i64, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i64, _i, _i, _i, _i, _i, _p],
    "onet_convT2x2_dgrad": [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _i64, _i, _i, _i, _i, _i, _p],
    "onet_convT2x2_wgrad": [_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p],
    "onet_zero_border": [_p, _i, _i, _i, _i64, _i, _i, _i, _i, _i, _p],
    "onet_add_colsums": [_p, _i, _p, _p],
    "onet_head_fwd": [_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p],
    "onet_head_bwd": [_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p],
    "onet_head_fwd_bn": [_p, _i64, _i, _p, _i64, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p],
    "onet_head_bwd_scalars": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p],
    "onet_bn_relu_bwd_head": [_p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _i64, _i, _p, _p, _p, _p, _d, _p, _p, _p, _p, _p, _i, _p],
    "onet_predict_label": [_p, _p, _i64, _p, _p],
    "onet_predict_label_u8": [_p, _p, _i64, _p, _p],
    "onet_eval_confusion": [_p, _p, _p, _i64, _p, _p],
    "onet_normalize_per_frame": [_p, _i, _i64, _p, _p, _p],
    "onet_synth_rayleigh": [_p, _i64, _f, _i64, _i, _p],
    "onet_synth_kclutter": [_p, _i64, _i, _i64, _i, _p],
    "onet_synth_add_targets": [_p, _p, _i, _i, _i, _p, _i, _f, _p, _p],
    "onet_synth_normal": [_p, _i64, _i64, _i, _p],
    "onet_kfield_mnlt": [_p, _i64, _i, _p, _p],
    "onet_kfield_coeff_sums": [_p, _p, _i, _i64, _p, _p],
    "onet_kfield_acf_root": [_p, _p, _i, _i64, _p, _p],
    "onet_kfield_amplitude": [_p, _p, _i64, _p, _p],
    "onet_adam_step": [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i, _f, _p],
    "onet_adam_step_dev": [_p, _p, _p, _p, _i64, _p, _p, _f, _p],
}

_lib = None
LAUNCHES = 0   # number of kernel-launching C-ABI calls made by this process (bench.py reports it)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise OnetLibError(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU / PyTorch fallback for the Onet hot path")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        _lib.onet_last_error.restype = ctypes.c_char_p
        _lib.onet_last_error.argtypes = []
        _lib.onet_last_kernel.restype = ctypes.c_char_p
        _lib.onet_last_kernel.argtypes = []
        _lib.onet_launch_count.restype = ctypes.c_int64
        _lib.onet_launch_count.argtypes = []
    return _lib


PROFILE = None   # set to a list to record (name, int args, start event, end event) of every call (bench.py)


def call(name, *args, device=None):
    """Invoke a C-ABI entry point; raise with the library's message on a non-zero status.  The library launches on the
    CURRENT device: call sites outside `Onet._run_forward` / backward (which set it themselves) pass `device=`."""
    global LAUNCHES
    if device is not None:
        import torch
        with torch.cuda.device(device):
            return call(name, *args)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise OnetLibError(f"{name} failed: {lib().onet_last_error().decode()}")
    LAUNCHES += 1
    if PROFILE is not None:
        e1.record()
        PROFILE.append((name, args, e0, e1, lib().onet_last_kernel().decode()))


def launch_count():
    """Kernels launched through the library by this process."""
    return int(lib().onet_launch_count())


def ptr(t, elem_offset=0):
    """Device pointer of a torch tensor (+ element offset), or None."""
    if t is None:
        return None
    return t.data_ptr() + elem_offset * t.element_size()
