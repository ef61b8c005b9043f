"""Import the UNMODIFIED reference Onet module from /root/reference (container only).

TEST INFRASTRUCTURE — not product code.  Only `tests/golden/make_golden.py` and the
`-m "not gpu"` pinning tests may call this, and only where /root/reference exists
(it does not exist on the GPU box).

The reference module `source_code/Onet_vanilla_20240606.py` imports matplotlib, skimage and
(through `dataloader/simbg4onet_20230209.py:18-20`) albumentations at module top although the
model classes never use them (Onet_vanilla_20240606.py:14,20,26).  Those three packages are not
installed in this image, so empty stub modules are registered before the import.  Nothing of the
reference is copied: the classes are used exactly as shipped.
"""
import os
import sys
import types

REF_ROOT = "/root/reference/source_code"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Onet_vanilla_20240606.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference module object (classes DoubleConv/Down/Up/UNet/Onet)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    try:
        import matplotlib  # noqa: F401
    except Exception:
        dummy = lambda *a, **k: None
        mpl = _stub("matplotlib", use=dummy)
        plt = _stub("matplotlib.pyplot")
        pat = _stub("matplotlib.patches", Ellipse=object, Rectangle=object)
        mpl.pyplot = plt
        mpl.patches = pat
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = _stub("skimage")
        tr = _stub("skimage.transform", resize=lambda *a, **k: None)
        sk.transform = tr
    try:
        import albumentations  # noqa: F401
    except Exception:
        names = ["Compose", "Defocus", "CLAHE", "Equalize", "PixelDropout", "GaussianBlur",
                 "RandomBrightnessContrast", "CoarseDropout", "HorizontalFlip"]
        _stub("albumentations", **{n: (lambda *a, **k: None) for n in names})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    argv = sys.argv
    sys.argv = ["x"]
    try:
        import Onet_vanilla_20240606 as ref  # noqa: E402
    finally:
        sys.argv = argv
    return ref
