"""fused_leaky_relu / FusedLeakyReLU behind the reference's signatures (reference op/fused_act.py).

Autograd structure mirrors the reference (Function -> Backward-Function -> forward op again) so the
R1 / path-length double-backward keeps working (reference op/fused_act.py:20-71).  Differences, all
inside the boundary: the bias gradient is reduced inside the backward kernel (no second pass), and
channels-last tensors are processed in place (no .contiguous() copy).
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib


def _geometry(x):
    """(tensor laid out densely, step_b, size_b) such that channel(i) = (i / step_b) % size_b."""
    if x.dim() < 2:
        raise RuntimeError("fused_bias_act: input needs a channel dimension (dim 1)")
    c = x.shape[1]
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        return x, 1, c                               # NHWC storage: channel is the fastest axis
    x = x.contiguous()
    step = 1
    for d in x.shape[2:]:
        step *= d
    return x, step, c


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    """Mirror of the reference pybind entry `fused.fused_bias_act` (reference op/fused_bias_act.cpp:5-30):
    empty tensors mean "absent"; returns a new tensor shaped like `input`."""
    _lib.require_cuda(input, "fused_bias_act")
    if input.dtype != torch.float32:
        raise RuntimeError("fused_bias_act: float32 only (like the reference kernel)")
    x, step_b, size_b = _geometry(input)
    b = bias.contiguous() if bias is not None and bias.numel() else None
    r = None
    if refer is not None and refer.numel():
        r = refer
        if r.stride() != x.stride() or r.shape != x.shape:
            r = r.expand_as(x).contiguous(memory_format=torch.channels_last if step_b == 1 and x.dim() == 4
                                          else torch.contiguous_format)
    if b is not None and b.numel() != size_b:
        raise RuntimeError(f"fused_bias_act: bias has {b.numel()} elements, expected {size_b}")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().sr_fused_bias_act_f32(_lib.ptr(y), _lib.ptr(x), _lib.ptr(b), _lib.ptr(r), int(act), int(grad),
                                              float(alpha), float(scale), x.numel(), step_b, size_b,
                                              _lib.stream_of(x))
    _lib.check(rc, "sr_fused_bias_act_f32")
    return y


def _lrelu_backward(grad_output, out, negative_slope, scale, want_bias):
    x, step_b, size_b = _geometry(out)
    g = grad_output
    if g.stride() != x.stride():
        g = g.contiguous(memory_format=torch.channels_last if step_b == 1 and x.dim() == 4
                         else torch.contiguous_format)
    dx = torch.empty_like(x)
    db = torch.empty(size_b, dtype=torch.float32, device=x.device) if want_bias else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().sr_fused_lrelu_backward_f32(_lib.ptr(dx), _lib.ptr(db), _lib.ptr(g), _lib.ptr(x),
                                                    float(negative_slope), float(scale), x.numel(), step_b, size_b,
                                                    _lib.stream_of(x))
    _lib.check(rc, "sr_fused_lrelu_backward_f32")
    return dx, db


class FusedLeakyReLUFunctionBackward(Function):     # reference op/fused_act.py:20-49
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input, grad_bias = _lrelu_backward(grad_output, out, negative_slope, scale, True)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = fused_bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):             # reference op/fused_act.py:52-71
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):                    # reference op/fused_act.py:74-83
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """reference op/fused_act.py:86-97: CPU tensors take plain torch ops (like the reference's dispatcher), CUDA tensors
    the sm_100a kernel (no fallback there).  The reference's CPU branch hard-codes slope 0.2 whatever `negative_slope`
    says (op/fused_act.py:91); that is reproduced, because it is what every CPU caller of the reference gets."""
    if input.device.type == "cpu":
        shape = [1, bias.shape[0]] + [1] * (input.dim() - 2)
        return torch.nn.functional.leaky_relu(input + bias.reshape(shape), 0.2) * scale
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
