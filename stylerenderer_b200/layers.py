"""StyleGAN2 building blocks behind the reference's class names, constructor signatures and state_dict keys
(reference layers.py).  Host code stays PyTorch; every operator the reference implements in op/ is a
hand-written sm_100a kernel reached through stylerenderer_b200.op.

ModulatedConv2d is re-formulated B200-first ("Hard parts"): instead of materialising one
modulated+demodulated weight tensor per sample and running a grouped conv with B groups (reference
layers.py:296-322; 302 MB of weights per 512x512 layer at B=32), the per-sample style scales the
*activations* and the demodulation scales the *outputs*:

    y[b,o] = d[b,o] * conv( x[b,i] * s[b,i],  scale * W[o,i] ),   d = rsqrt( (s^2) @ (scale^2 * sum_k W^2)^T + eps )

so all samples share one weight matrix and the contraction becomes ONE dense implicit GEMM
(M = B*H*W, N = Cout, K = k*k*Cin).  `conv_backend` selects who runs that GEMM:
    "tcgen05" -- the hand-written sm_100a kernel (csrc/modconv.cu), where available for the shape;
    "cudnn"   -- stock F.conv2d / F.conv_transpose2d (library baseline; never counted in gpu_launches).
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

_CONFIG = {"conv_backend": "cudnn", "double_backward": False}
HAVE_TCGEN05 = True


def set_conv_backend(name):
    assert name in ("cudnn", "tcgen05")
    _CONFIG["conv_backend"] = name


def get_conv_backend():
    return _CONFIG["conv_backend"]


class double_backward:
    """Context manager for regulariser iterations (R1 / path length, reference train.py:110-134): inside it the
    convolutions run as twice-differentiable autograd Functions (fused.mod_conv_dd / fused.plain_conv_dd on the
    tensor-core kernels, or the composed torch ops with the "cudnn" backend) instead of the fused first-order blocks."""

    def __enter__(self):
        self._old = _CONFIG["double_backward"]
        _CONFIG["double_backward"] = True

    def __exit__(self, *a):
        _CONFIG["double_backward"] = self._old


def double_backward_requested():
    return _CONFIG["double_backward"]


def make_kernel(k):                                   # reference layers.py:7-12
    k = torch.tensor(k, dtype=torch.float32)
    if k.dim() == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class PixelNorm(nn.Module):                           # reference layers.py:100-105
    def __init__(self, eps=1e-8):
        super().__init__()
        self.eps = abs(eps)

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input * input, -1, keepdim=True) + self.eps)


class Upsample(nn.Module):                            # reference layers.py:170-181
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):                          # reference layers.py:182-193
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):                                # reference layers.py:194-203
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):                         # reference layers.py:204-221
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        if _CONFIG["conv_backend"] == "tcgen05" and input.is_cuda and self.weight.shape[0] <= 8 and self.weight.shape[1] <= 8:
            # convolutions with a handful of channels (the style-map nets of GeneratorWithMap): two small kernels that are
            # each other's derivatives, differentiable to any order (fused.SmallConvFn)
            from . import fused
            if fused.small_conv_supported(self, input):
                return fused.small_conv(self, input)
        w = self.weight * self.scale
        if (w.shape[2] == 1 and w.shape[3] == 1 and self.stride == 1 and self.padding == 0 and w.shape[1] <= 8 and input.is_cuda
                and input.dim() == 4):
            # pointwise conv with a handful of input channels (the Discriminator's 3 -> C stem, reference model.py:303) =
            # one [N*H*W, cin] x [cin, cout] product, output in channels_last.  cuDNN picks an "indexed, without shared
            # memory" kernel for this shape and its double backward (R1 iterations): 6.3 ms per call at 256^2, batch 16
            # (torch.profiler, profiles/r2_train_step_profile.md)
            out = torch.matmul(input.permute(0, 2, 3, 1), w[:, :, 0, 0].t())
            if self.bias is not None:
                out = out + self.bias
            return out.permute(0, 3, 1, 2)
        return F.conv2d(input, w, bias=self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        return "%s(%d, %d, %d, stride=%d, padding=%d)" % (self.__class__.__name__, self.weight.shape[1],
                                                          self.weight.shape[0], self.weight.shape[2], self.stride,
                                                          self.padding)


class EqualLinear(nn.Module):                         # reference layers.py:222-251
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation == "fused_lrelu":
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        out = F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)
        if self.activation == "relu":
            out = F.relu(out)
        elif self.activation == "lrelu":
            out = F.leaky_relu(out, negative_slope=0.2)
        elif self.activation == "selu":
            out = F.selu(out)
        elif self.activation == "tanh":
            out = torch.tanh(out)
        return out

    def __repr__(self):
        return "%s(%d, %d)" % (self.__class__.__name__, self.weight.shape[1], self.weight.shape[0])


class ScaledLeakyReLU(nn.Module):                     # reference layers.py:252-258
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, negative_slope=self.negative_slope) * math.sqrt(2)


class ModulatedConv2d(nn.Module):                     # reference layers.py:259-323
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        fan_in = in_channel * kernel_size ** 2
        self.scale = 1 / math.sqrt(fan_in)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return "%s(%d, %d, %d, upsample=%s, downsample=%s)" % (self.__class__.__name__, self.in_channel,
                                                               self.out_channel, self.kernel_size, self.upsample,
                                                               self.downsample)

    def style_scales(self, style):
        """(s [B,Cin], d [B,Cout] or None): per-sample input modulation and output demodulation."""
        s = self.modulation(style)
        d = None
        if self.demodulate:
            wsq = (self.weight[0] * self.scale).pow(2).sum([2, 3])               # [Cout, Cin]
            d = torch.rsqrt(F.linear(s * s, wsq) + self.eps)                       # reference layers.py:297
        return s, d

    def contract(self, x_mod):
        """The shared-weight contraction of the already modulated input (no demodulation)."""
        w = self.weight[0] * self.scale
        if self.upsample:                                                          # reference layers.py:301-310
            out = F.conv_transpose2d(x_mod, w.transpose(0, 1), padding=0, stride=2)
            return self.blur(out)
        if self.downsample:                                                        # reference layers.py:311-317
            return F.conv2d(self.blur(x_mod), w, padding=0, stride=2)
        return F.conv2d(x_mod, w, padding=self.padding)                            # reference layers.py:318-322

    def forward(self, input, style):
        batch, in_channel = input.shape[:2]
        if _CONFIG["conv_backend"] == "tcgen05":
            from . import fused                          # hand-written tensor-core contraction where the shape allows
            if fused.supported(self, input):
                if torch.is_grad_enabled() and _CONFIG["double_backward"]:
                    return fused.mod_conv_dd(self, input, style)     # twice differentiable (R1 / path-length iterations)
                return fused.mod_conv(self, input, style)
        s, d = self.style_scales(style)
        if self.kernel_size == 1 and not self.upsample and not self.downsample:
            # ToRGB: fold the style into B x [Cout, Cin] weights (a few KB) and read the activation once
            wb = (self.weight[0, :, :, 0, 0] * self.scale).unsqueeze(0) * s.unsqueeze(1)
            if d is not None:
                wb = wb * d.unsqueeze(2)
            h, w = input.shape[2:]
            if input.is_contiguous(memory_format=torch.channels_last) and not input.is_contiguous():
                hw, chunk = h * w, 1024
                if hw >= 4 * chunk and hw % chunk == 0:
                    # pixels in chunks of 1024 with the per-sample weights broadcast over the chunks: autograd's weight
                    # gradient then is a batch of [cin, 1024] x [1024, cout] products summed over the chunks (split K)
                    # instead of ONE [cin, h*w] x [h*w, 3] product per sample, which cuBLAS serves at 1.45 ms for
                    # [128, 65536] x [65536, 3] (torch.profiler, regulariser iterations of the train step)
                    x4 = input.permute(0, 2, 3, 1).reshape(batch, hw // chunk, chunk, in_channel)
                    out = torch.matmul(x4, wb.transpose(1, 2).unsqueeze(1))
                else:
                    out = torch.matmul(input.permute(0, 2, 3, 1).reshape(batch, hw, in_channel), wb.transpose(1, 2))
                return out.view(batch, h, w, self.out_channel).permute(0, 3, 1, 2)
            return torch.bmm(wb, input.reshape(batch, in_channel, h * w)).view(batch, self.out_channel, h, w)
        out = self.contract(input * s.view(batch, in_channel, 1, 1))
        if d is not None:
            out = out * d.view(batch, self.out_channel, 1, 1)
        return out


class NoiseInjection(nn.Module):                      # reference layers.py:324-332
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):                       # reference layers.py:333-340
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class ConvLayer(nn.Sequential):                       # reference layers.py:341-378
    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate="lrelu"):
        layers = []
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride = 2
            self.padding = 0
        else:
            stride = 1
            self.padding = kernel_size // 2
        if activate is False or activate is None:        # reference ResBlock passes activate=False (quirk #2):
            activate = "none"                            # upstream semantics = no activation
        if "sp" in activate.lower():
            raise NotImplementedError("SpectralNorm conv layers are outside the hot path (SURVEY.md section 2 row 4)")
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride, bias=bias))
        if activate == "lrelu":
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)

    def forward(self, input):
        # tensor-core path (conv_backend "tcgen05"): [Blur ->] EqualConv2d [+ FusedLeakyReLU] with the bias / activation in
        # the conv epilogue, or -- on R1 / path regulariser iterations under double_backward() -- the twice-differentiable
        # ConvTC + fused_leaky_relu; anything else (3-channel stems, odd shapes) takes the composed cuDNN path below
        if _CONFIG["conv_backend"] == "tcgen05":
            dd = torch.is_grad_enabled() and _CONFIG["double_backward"]
            mods = list(self)
            blur = mods[0] if isinstance(mods[0], Blur) else None
            rest = mods[1:] if blur is not None else mods
            conv = rest[0]
            act = rest[1] if len(rest) > 1 else None
            if isinstance(conv, EqualConv2d) and (isinstance(act, FusedLeakyReLU) or (act is None and conv.bias is None)):
                from . import fused
                if blur is None and isinstance(act, FusedLeakyReLU) and fused.stem_conv_supported(conv, act, input):
                    return fused.stem_conv(conv, act, input)          # 3 -> C pointwise stem: one bandwidth pass
                if blur is not None and not dd and fused.env_blur_conv():
                    kind = fused.blur_conv_supported(conv, blur, input)
                    if kind is not None:                              # FIR writes the GEMM operand directly (fused.BlurConvTC)
                        return fused.blur_conv(conv, act, blur, input, kind)
                x = blur(input) if blur is not None else input
                kind = fused.plain_conv_supported(conv, x)
                if kind is not None:
                    return fused.plain_conv_dd(conv, act, x, kind) if dd else fused.plain_conv(conv, act, x, kind)
                input, start = x, (1 if blur is not None else 0)
                for m in mods[start:]:
                    input = m(input)
                return input
        return super().forward(input)


class ResBlock(nn.Module):                            # reference layers.py:379-391
    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1], downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=downsample)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=downsample, activate=False, bias=False)

    def forward(self, input):
        if _CONFIG["conv_backend"] == "tcgen05" and input.is_cuda and input.dim() == 4 and input.shape[1] == 3:
            # the style-map networks of GeneratorWithMap (reference model.py:194-216): the whole block in one kernel
            from . import fused
            if fused.stylemap_resblock_supported(self, input):
                return fused.stylemap_resblock(self, input)
        out, skip = self.conv2(self.conv1(input)), self.skip(input)
        if _CONFIG["conv_backend"] == "tcgen05" and out.is_cuda and out.shape[1] % 128 == 0:
            from . import fused
            if fused.residual_combine_supported(out, skip):       # one pass: residual sum + the next conv's GEMM operand
                return fused.residual_combine(out, skip, 1 / math.sqrt(2))
        return (out + skip) / math.sqrt(2)
