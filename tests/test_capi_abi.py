"""The C-ABI boundary without a GPU: libonet_b200.so builds for sm_100a, loads, exports every symbol that
include/onet_b200.h declares, and the ctypes signature table of onet_b200/_lib.py matches the header's prototypes
(argument count and pointer / integer / floating kinds).  No compute call is made here."""
import ctypes
import os
import re

import pytest

from onet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "onet_b200.h")


def _prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(onet_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argl = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, argl)
    return protos


def _kind(carg):
    if "*" in carg:
        return "p"
    t = carg.rsplit(" ", 1)[0].replace("const ", "").strip()
    return {"int": "i", "int64_t": "l", "float": "f", "double": "d"}[t]


_CT = {ctypes.c_void_p: "p", ctypes.c_int: "i", ctypes.c_int64: "l", ctypes.c_float: "f", ctypes.c_double: "d"}


@pytest.fixture(scope="module")
def lib_path():
    return _lib.build(force=False)


def test_header_declares_the_hot_path_entry_points():
    names = set(_prototypes())
    for must in ("onet_conv3x3_fwd", "onet_conv3x3_wgrad", "onet_bn_finalize", "onet_bn_relu_apply", "onet_bn_relu_bwd",
                 "onet_convT2x2_fwd", "onet_convT2x2_dgrad", "onet_convT2x2_wgrad", "onet_head_fwd", "onet_head_bwd",
                 "onet_predict_label", "onet_adam_step", "onet_prep_input", "onet_version", "onet_last_error"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _prototypes():
        assert hasattr(lib, name), f"{name} declared in include/onet_b200.h but not exported by {lib_path}"
    lib.onet_version.restype = ctypes.c_int
    assert lib.onet_version() >= 100


def test_ctypes_signatures_match_header():
    protos = _prototypes()
    for name, (ret, args) in protos.items():
        if name in ("onet_last_error", "onet_launch_count", "onet_last_kernel"):
            continue
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in onet_b200/_lib.py"
        want = [_kind(a) for a in args]
        got = [_CT[t] for t in _lib.SIGNATURES[name]]
        assert want == got, f"{name}: header {want} vs ctypes {got}"
    assert set(_lib.SIGNATURES) <= set(protos), set(_lib.SIGNATURES) - set(protos)


def test_library_is_sm100a_with_tcgen05_and_tma(lib_path):
    """The shipped binary must contain sm_100a SASS with tensor-core (UTC*MMA), TMEM (LDTM) and TMA (UTMALDG) code."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, f"{mnem} missing from the SASS of {lib_path}"


def test_no_cpu_fallback_in_product_package():
    """The product package never imports the oracle and raises on CPU tensors instead of falling back."""
    import torch
    import onet_b200
    pkg = os.path.join(ROOT, "onet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), f"{fn} mentions the oracle"
    net = onet_b200.Onet(1, True, True)
    with pytest.raises(_lib.OnetLibError):
        net(torch.rand(1, 1, 16, 16))


def test_driver_entry_points_and_tools_compile():
    """__graft_entry__.py (build / smoke), bench.py and every tool script must at least byte-compile and import."""
    import glob
    import importlib
    import py_compile
    for fn in ["__graft_entry__.py", "bench.py"] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))):
        py_compile.compile(os.path.join(ROOT, fn), doraise=True)
    ge = importlib.import_module("__graft_entry__")
    assert callable(ge.build) and callable(ge.smoke)


def test_header_is_plain_c_and_links_from_a_c_program(lib_path, tmp_path):
    """include/onet_b200.h must be consumable by a C compiler (C99, no C++-isms) and the library must link from plain C:
    a tiny program takes the address of every declared entry point and calls onet_version() (no GPU needed)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    names = sorted(_prototypes())
    src = tmp_path / "use_onet.c"
    src.write_text('#include <stdio.h>\n#include "onet_b200.h"\nint main(void) {\n    const void* fns[] = {\n'
                   + "".join(f"        (const void*)&{n},\n" for n in names)
                   + '    };\n    printf("%d %d\\n", onet_version(), (int)(sizeof(fns) / sizeof(fns[0])));\n    return 0;\n}\n')
    exe = tmp_path / "use_onet"
    libdir = os.path.dirname(lib_path)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libonet_b200.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ver, n = out.stdout.split()
    assert int(ver) >= 100 and int(n) == len(names)
