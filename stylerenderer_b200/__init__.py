"""stylerenderer_b200 -- B200-native (sm_100a) hot path of WestlyPark/StyleRenderer.

Drop-in surface (same names / signatures as the reference):
    stylerenderer_b200.op.upfirdn2d / fused_leaky_relu / FusedLeakyReLU / rasterize   (reference op/__init__.py:1-3)
    stylerenderer_b200.layers.ModulatedConv2d, Blur, Upsample, EqualLinear, ...        (reference layers.py)
    stylerenderer_b200.model.StyledConv / StyledMapConv / ToRGB / Generator / ...      (reference model.py)
Every operator is a hand-written CUDA kernel in csrc/, exported through the C ABI declared in
include/stylerenderer_b200.h and reached via ctypes (_lib.py).  No CPU fallback exists.
"""
from . import _lib  # noqa: F401

__all__ = ["op", "layers", "model"]
