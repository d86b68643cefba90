"""upfirdn2d behind the reference's signature (reference op/upfirdn2d.py).

Function -> Backward-Function -> forward op again, exactly like the reference (op/upfirdn2d.py:19-142):
the gradient of upfirdn2d is upfirdn2d with flipped taps, up and down swapped and the pads of
op/upfirdn2d.py:111-114, and its own gradient is the forward operator, so double-backward works.
"""
import torch
from torch.autograd import Function

from .. import _lib


def upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Mirror of the reference pybind entry `upfirdn2d.upfirdn2d` (reference op/upfirdn2d.cpp:24-83):
    input [major, in_h, in_w, minor] -> new tensor [major, out_h, out_w, minor]."""
    _lib.require_cuda(input, "upfirdn2d")
    if input.dtype != torch.float32 or kernel.dtype != torch.float32:
        raise RuntimeError("upfirdn2d: float32 only (like the reference kernel)")
    if input.dim() != 4 or kernel.dim() != 2:
        raise RuntimeError("upfirdn2d: expected input [major,h,w,minor] and a 2-D kernel")
    x = input.contiguous()
    k = kernel.contiguous()
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1
    if out_h < 1 or out_w < 1:
        raise RuntimeError("upfirdn2d: FIR larger than the padded input")
    out = torch.empty(major, out_h, out_w, minor, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().sr_upfirdn2d_f32(_lib.ptr(out), _lib.ptr(x), _lib.ptr(k), major, in_h, in_w, minor, kh, kw,
                                         up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                         _lib.stream_of(x))
    _lib.check(rc, "sr_upfirdn2d_f32")
    return out


def _is_channels_last(t):
    return t.dim() == 4 and t.shape[1] > 1 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _planes_or_nhwc(t, h, w):
    """[N,C,h,w] -> ([major,h,w,minor] view for the kernel, channels_last?).  A channels_last tensor IS an
    [N,h,w,C] array in memory, which the kernel takes with major = N, minor = C (no layout copy)."""
    if _is_channels_last(t) and t.shape[1] % 4 == 0:
        return t.permute(0, 2, 3, 1), True
    return t.reshape(-1, h, w, 1), False


def _restore(out4, n, c, cl):
    """kernel output [major,oh,ow,minor] -> logical [N,C,oh,ow] (channels_last strides if the input had them)."""
    if cl:
        return out4.permute(0, 3, 1, 2)
    return out4.view(n, c, out4.shape[1], out4.shape[2])


class UpFirDn2dBackward(Function):                  # reference op/upfirdn2d.py:19-85
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_output, cl = _planes_or_nhwc(grad_output, out_size[0], out_size[1])
        grad_input = upfirdn2d_raw(grad_output, grad_kernel, down_x, down_y, up_x, up_y,
                                   g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        grad_input = _restore(grad_input, in_size[0], in_size[1], cl)
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_input, cl = _planes_or_nhwc(gradgrad_input, ctx.in_size[2], ctx.in_size[3])
        gradgrad_out = upfirdn2d_raw(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad)
        gradgrad_out = _restore(gradgrad_out, ctx.in_size[0], ctx.in_size[1], cl)
        return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):                          # reference op/upfirdn2d.py:88-142
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        input, cl = _planes_or_nhwc(input, in_h, in_w)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        ctx.g_pad = (kernel_w - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        out = upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return _restore(out, batch, channel, cl)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                             ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d_cpu(input, kernel, up=1, down=1, pad=(0, 0)):
    """The operator on CPU tensors in plain torch ops (differentiable to any order through autograd) -- the role of the
    reference's `upfirdn2d_native` (reference op/upfirdn2d.py:159-200), which its dispatcher takes for CPU tensors
    (op/upfirdn2d.py:146-150; callers: utils_face.py:515-517, BASELINE.json configs[0]).  Steps per [n, c] plane:
    zero-stuff by `up`, pad (negative pads crop), true convolution with `kernel`, keep every `down`-th sample."""
    import torch.nn.functional as F
    n, c, h, w = input.shape
    kh, kw = kernel.shape
    planes = input.reshape(n * c, 1, h, w)
    if up > 1:                                                   # sample (i, j) -> (i*up, j*up), zeros in between
        stuffed = planes.new_zeros(n * c, 1, h * up, w * up)
        stuffed[:, :, ::up, ::up] = planes
        planes = stuffed
    planes = F.pad(planes, [pad[0], pad[1], pad[0], pad[1]])     # F.pad crops for negative amounts
    taps = torch.flip(kernel, [0, 1]).to(planes.dtype).reshape(1, 1, kh, kw)   # conv2d correlates: flip = convolution
    out = F.conv2d(planes, taps)[:, :, ::down, ::down]
    return out.reshape(n, c, out.shape[2], out.shape[3])


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """reference op/upfirdn2d.py:145-157: CPU tensors take the plain-torch formulation (like the reference's dispatcher),
    CUDA tensors the sm_100a kernel -- which fails loudly when the library is missing, it never falls back."""
    if input.device.type == "cpu":
        return upfirdn2d_cpu(input, kernel, up, down, pad)
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
