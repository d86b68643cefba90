"""ctypes binding of libstylerenderer_b200.so -- the only way Python reaches the sm_100a kernels.

The library is a plain C ABI (include/stylerenderer_b200.h): raw device pointers, sizes and a
cudaStream_t.  This module only marshals `tensor.data_ptr()` / the current torch stream into those
calls; it allocates nothing and owns no CUDA state.  There is NO fallback: if the shared object is
missing or a call fails, an exception is raised (the product must never silently run on CPU or on a
library path).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstylerenderer_b200.so")

_lib = None

_P, _I, _L, _F, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

_SIGNATURES = {
    "sr_abi_version": (_I, []),
    "sr_last_error": (ctypes.c_char_p, []),
    "sr_launch_count": (_L, []),
    "sr_fused_bias_act_f32": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _L, _L, _L, _P]),
    "sr_fused_lrelu_backward_f32": (_I, [_P, _P, _P, _P, _F, _F, _L, _L, _L, _P]),
    "sr_upfirdn2d_f32": (_I, [_P, _P, _P, _L, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "sr_rasterize_workspace_bytes": (_L, [_L, _L, _L, _L, _L, _I]),
    "sr_rasterize_forward_f32": (_I, [_L, _L, _L, _L, _L, _I, _I, _I, _P, _P, _P, _P, _P, _F, _P, _L, _P, _P]),
    "sr_rasterize_forward_f64": (_I, [_L, _L, _L, _L, _L, _I, _I, _I, _P, _P, _P, _P, _P, _D, _P, _L, _P, _P]),
    "sr_rasterize_dcoeff_f32": (_I, [_L, _L, _L, _L, _I, _P, _P, _P, _F, _P]),
    "sr_rasterize_dcoeff_f64": (_I, [_L, _L, _L, _L, _I, _P, _P, _P, _D, _P]),
    "sr_rasterize_backward_f32": (_I, [_L, _L, _L, _L, _L, _I, _P, _P, _P, _P, _P, _P, _P, _F, _P]),
    "sr_rasterize_backward_f64": (_I, [_L, _L, _L, _L, _L, _I, _P, _P, _P, _P, _P, _P, _P, _D, _P]),
    "sr_rasterize_pyramid_workspace_bytes": (_L, [_L, _L, _L, _I, _P]),
    "sr_rasterize_pyramid_forward_f32": (_I, [_L, _L, _L, _I, _P, _I, _I, _I, _P, _P, _P, _F, _P, _L, _P]),
    "sr_rasterize_pyramid_backward_f32": (_I, [_L, _L, _I, _P, _L, _I, _P, _P, _P, _P, _F, _P]),
    "sr_conv_igemm_tf32": (_I, [_P, _P]),
    "sr_conv_igemm_multi_tf32": (_I, [_P, _I, _P]),
    "sr_conv_wgrad_tf32": (_I, [_P, _P]),
    "sr_modulate_tf32": (_I, [_P, _P, _P, _L, _L, _L, _P]),
    "sr_conv_weight_prep_tf32": (_I, [_P, _P, _F, _L, _L, _I, _I, _I, _P]),
    "sr_blur_nhwc_styled_f32": (_I, [_P, _P, _P, _L, _L, _L, _L, _I, _I, _P, _L, _P, _P, _F, _F, _P]),
    "sr_styled_bwd_prologue_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _L, _L, _L, _F, _F, _P]),
    "sr_styled_bwd_prologue2_f32": (_I, [_P] * 13 + [_L, _P, _P, _P, _L, _L, _L, _F, _F, _P]),
    "sr_blur_nhwc_scaledot_f32": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P]),
    "sr_blur_nhwc_styled2_f32": (_I, [_P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P, _L, _P, _P, _F, _F, _P]),
    "sr_scale_dot_nhwc_f32": (_I, [_P, _P, _P, _P, _P, _L, _L, _L, _I, _P]),
    "sr_style_scales_forward_f32": (_I, [_P, _I, _P, _L, _L, _L, _F, _F, _F, _P]),
    "sr_style_scales_backward_f32": (_I, [_P, _I, _P, _P, _L, _L, _L, _F, _F, _P]),
    "sr_weight_sq_f32": (_I, [_P, _P, _F, _L, _L, _I, _P]),
    "sr_weight_sq_backward_f32": (_I, [_P, _P, _P, _F, _L, _L, _I, _P]),
    "sr_weight_grad_layout_f32": (_I, [_P, _P, _F, _L, _L, _I, _P]),
    "sr_conv_weight_prep_dual_tf32": (_I, [_P, _P, _P, _P, _F, _L, _L, _I, _I, _P]),
    "sr_conv_igemm_multi_bf16": (_I, [_P, _I, _P]),
    "sr_conv_wgrad_bf16": (_I, [_P, _P]),
    "sr_modulate_bf16": (_I, [_P, _P, _P, _L, _L, _L, _P]),
    "sr_conv_weight_prep_dual_bf16": (_I, [_P, _P, _P, _P, _F, _L, _L, _I, _I, _P]),
    "sr_blur_nhwc_styled3_bf16": (_I, [_P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P, _L, _P, _P, _F, _F, _P, _L, _P]),
    "sr_blur_nhwc_scaledot_bf16": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P]),
    "sr_styled_bwd_prologue3_bf16": (_I, [_P] * 13 + [_L, _P, _P, _P, _L, _L, _L, _F, _F, _P, _L, _P, _P]),
    "sr_mesh_vertex_normals_f32": (_I, [_P, _P, _P, _L, _L, _L, _I, _F, _P]),
    "sr_mesh_pose_apply_f32": (_I, [_P, _P, _P, _L, _L, _L, _P]),
    "sr_mesh_normal_pyramid_f32": (_I, [_L, _L, _L, _P, _L, _P, _P, _P, _P, _I, _P, _P, _F, _P]),
    "sr_rasterize_pyramid_maps_f32": (_I, [_L, _L, _L, _I, _P, _I, _I, _I, _P, _P, _P, _F, _P, _L, _I, _P]),
    "sr_blur_nhwc_styled3_f32": (_I, [_P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P, _L, _P, _P, _F, _F, _P, _L, _P]),
    "sr_styled_bwd_prologue3_f32": (_I, [_P] * 13 + [_L, _P, _P, _P, _L, _L, _L, _F, _F, _P, _L, _P, _P]),
    "sr_stylemap_resblock_forward_f32": (_I, [_P] * 9 + [_L, _I, _I, _L, _L, _F, _F, _P]),
    "sr_stylemap_resblock_backward_f32": (_I, [_P] * 10 + [_L, _I, _I, _L, _L, _F, _F, _P]),
    "sr_conv_weight_prep_multi_tf32": (_I, [_P, _I, _P]),
    "sr_conv_weight_prep_multi_bf16": (_I, [_P, _I, _P]),
    "sr_weight_sq_backward_multi_f32": (_I, [_P, _I, _P]),
    "sr_residual_combine_tf32": (_I, [_P, _P, _P, _P, _F, _L, _P]),
    "sr_residual_combine_bf16": (_I, [_P, _P, _P, _P, _F, _L, _P]),
    "sr_small_conv_f32": (_I, [_P, _P, _P, _L, _I, _I, _I, _L, _L, _P]),
    "sr_small_conv_wgrad_f32": (_I, [_P, _P, _P, _L, _I, _I, _I, _L, _L, _P]),
    "sr_stem_conv_forward_f32": (_I, [_P] * 5 + [_L, _I, _L, _L, _L, _I, _F, _F, _P]),
    "sr_stem_conv_backward_f32": (_I, [_P] * 7 + [_L, _I, _L, _L, _L, _I, _F, _F, _P]),
}

EXPORTS = tuple(_SIGNATURES)
CONV_EXPORTS = ("sr_blur_nhwc_scaledot_f32", "sr_blur_nhwc_styled_f32", "sr_blur_nhwc_styled2_f32", "sr_styled_bwd_prologue_f32", "sr_styled_bwd_prologue2_f32", "sr_scale_dot_nhwc_f32",
                "sr_conv_igemm_tf32", "sr_conv_igemm_multi_tf32", "sr_conv_wgrad_tf32", "sr_modulate_tf32", "sr_conv_weight_prep_tf32",
                "sr_style_scales_forward_f32", "sr_style_scales_backward_f32", "sr_weight_sq_f32", "sr_weight_sq_backward_f32",
                "sr_weight_grad_layout_f32", "sr_conv_weight_prep_dual_tf32", "sr_blur_nhwc_styled3_f32",
                "sr_styled_bwd_prologue3_f32", "sr_conv_igemm_multi_bf16", "sr_conv_wgrad_bf16", "sr_modulate_bf16",
                "sr_conv_weight_prep_dual_bf16", "sr_blur_nhwc_styled3_bf16", "sr_blur_nhwc_scaledot_bf16",
                "sr_styled_bwd_prologue3_bf16", "sr_conv_weight_prep_multi_tf32", "sr_conv_weight_prep_multi_bf16",
                "sr_weight_sq_backward_multi_f32")


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the C-ABI library.  Raises NativeLibraryError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(stylerenderer_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)         # AttributeError here = header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def check(rc, what):
    if rc != 0:
        msg = lib().sr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: stylerenderer_b200 runs on CUDA tensors only (got device {t.device}); "
                           "there is deliberately no CPU fallback")


def launch_count():
    return int(lib().sr_launch_count())
