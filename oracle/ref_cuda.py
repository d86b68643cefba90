"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's GPU formulation, for timing beside ours.

SURVEY.md section 2a sets the bar "beat the reference's own generic kernels recompiled for sm_100a, plus cuDNN's grouped
conv".  The reference's Python layer cannot travel to the GPU box, its compiled `op/` extensions can (oracle/_ref, built by
oracle/build_ref.py from the unmodified sources).  This module wraps those compiled CUDA kernels in first-order autograd
Functions with the reference's gradient rules (op/upfirdn2d.py:19-142: the gradient of upfirdn2d is upfirdn2d with flipped
taps, up <-> down and the pads of :111-114; op/fused_act.py:20-71: dx = fused_bias_act(g, empty, out, act=3, grad=1), db =
dx summed over all but the channel axis) and plugs them into oracle/torch_ref.py, whose ModulatedConv2d is the reference's
per-sample-weight grouped-conv formulation (layers.py:293-323) -- on CUDA tensors that is cuDNN's grouped conv.

Only bench.py's `gpu_reference` leg and tests may import this."""
import contextlib

import torch
from torch.autograd import Function

from . import build_ref
from . import torch_ref as T

_MODS = {}


def _ext(name):
    if name not in _MODS:
        _MODS[name] = build_ref.load(name)
    return _MODS[name]


def available():
    try:
        _ext("ref_upfirdn2d"), _ext("ref_fused")
        return True
    except (FileNotFoundError, ImportError, OSError):
        return False


class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, x, taps, up, down, pad):
        n, c, h, w = x.shape
        kh, kw = taps.shape
        p0, p1 = pad
        out = _ext("ref_upfirdn2d").upfirdn2d(x.reshape(-1, h, w, 1), taps, up, up, down, down, p0, p1, p0, p1)
        oh, ow = out.shape[1], out.shape[2]
        ctx.save_for_backward(torch.flip(taps, [0, 1]))
        ctx.cfg = (n, c, h, w, oh, ow, up, down,
                   (kw - p0 - 1, w * up - ow * down + p0 - up + 1, kh - p0 - 1, h * up - oh * down + p0 - up + 1))
        return out.view(n, c, oh, ow)

    @staticmethod
    def backward(ctx, g):
        flipped, = ctx.saved_tensors
        n, c, h, w, oh, ow, up, down, gp = ctx.cfg
        gi = _ext("ref_upfirdn2d").upfirdn2d(g.reshape(-1, oh, ow, 1), flipped, down, down, up, up, *gp)
        return gi.view(n, c, h, w), None, None, None, None


class _FusedLeakyReLU(Function):
    @staticmethod
    def forward(ctx, x, bias, slope, scale):
        empty = x.new_empty(0)
        out = _ext("ref_fused").fused_bias_act(x, bias, empty, 3, 0, slope, scale)
        ctx.save_for_backward(out)
        ctx.cfg = (slope, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        out, = ctx.saved_tensors
        slope, scale = ctx.cfg
        empty = g.new_empty(0)
        gx = _ext("ref_fused").fused_bias_act(g.contiguous(), empty, out, 3, 1, slope, scale)
        dims = [0] + list(range(2, gx.dim()))
        return gx, gx.sum(dims), None, None


def upfirdn2d(x, taps, up=1, down=1, pad=(0, 0)):
    return _UpFirDn2d.apply(x, taps, up, down, pad)


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    return _FusedLeakyReLU.apply(x, bias, negative_slope, scale)


@contextlib.contextmanager
def reference_cuda_ops():
    """Inside: oracle.torch_ref modules run upfirdn2d / fused_leaky_relu on the reference's own CUDA kernels."""
    old = (T.upfirdn2d, T.fused_leaky_relu)
    T.upfirdn2d, T.fused_leaky_relu = upfirdn2d, fused_leaky_relu
    try:
        yield
    finally:
        T.upfirdn2d, T.fused_leaky_relu = old
