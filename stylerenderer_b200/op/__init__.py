"""Same public names as the reference's op package (reference op/__init__.py:1-3)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu, fused_bias_act
from .upfirdn2d import upfirdn2d, upfirdn2d_raw
from .rasterize import rasterize, rasterize_forward, rasterize_backward, rasterize_pyramid, rasterize_pyramid_maps

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "fused_bias_act", "upfirdn2d", "upfirdn2d_raw", "rasterize",
           "rasterize_forward", "rasterize_backward", "rasterize_pyramid", "rasterize_pyramid_maps"]
