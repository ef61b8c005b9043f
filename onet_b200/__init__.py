"""onet_b200 — B200-native (sm_100a) implementation of the Onet twin U-Net training / inference hot path.

Public surface (mirrors /root/reference/source_code/Onet_vanilla_20240606.py):
    Onet, UNet, DoubleConv, Down, Up
plus the build helper `build()` and the raw C-ABI binding in `onet_b200._lib`.
"""
from ._lib import build, LIB_PATH, OnetLibError  # noqa: F401
from .model import Onet, UNet, DoubleConv, Down, Up, invalidate_packed_weights  # noqa: F401

__all__ = ["Onet", "UNet", "DoubleConv", "Down", "Up", "build", "invalidate_packed_weights"]
