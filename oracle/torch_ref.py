"""TEST INFRASTRUCTURE ONLY -- native-PyTorch (CPU) restatement of the reference module stack.

This is the "reference's native-PyTorch CPU fallback" BASELINE.json names, restated so that it can
travel to the GPU box (the reference tree cannot): every op is a stock torch op on CPU tensors --
`upfirdn2d_native` (op/upfirdn2d.py:159-200), the `F.leaky_relu` branch of `fused_leaky_relu`
(op/fused_act.py:87-94) and the grouped-conv `ModulatedConv2d` (layers.py:293-323).  It is the
parity oracle for stylerenderer_b200.layers / .model and the `--impl reference` arm of bench.py.

Parameter / buffer names equal the reference's so `load_state_dict` of a reference checkpoint
works (checkpoint row); tests/test_oracle_pinning.py loads a state_dict of the
real reference modules (imported from /root/reference in the authoring container) and compares
outputs, and tests/golden/make_golden.py stores such outputs as fixtures.

Reference quirks kept on purpose: the duplicated ToRGB list (#4), the CPU
activation slope fixed at 0.2 (#3), `ConvLayer(activate=False)` meaning "no activation" (#2).
"""
import math

import torch
from torch import nn
from torch.nn import functional as F


# --------------------------------------------------------------------------- ops
def upfirdn2d(x, taps, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d.py:159-200 for NCHW input, same pad on both axes (:146-151)."""
    n, c, h, w = x.shape
    kh, kw = taps.shape
    p0, p1 = pad
    t = x.reshape(n * c, 1, h, 1, w, 1)
    t = F.pad(t, [0, up - 1, 0, 0, 0, up - 1])                      # zero-stuffing (:168-170)
    t = t.reshape(n * c, 1, h * up, w * up)
    t = F.pad(t, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])   # (:172-174)
    t = t[:, :, max(-p0, 0): t.shape[2] - max(-p1, 0), max(-p0, 0): t.shape[3] - max(-p1, 0)]
    t = F.conv2d(t, torch.flip(taps, [0, 1]).view(1, 1, kh, kw))      # true convolution (:186-188)
    t = t[:, :, ::down, ::down]                                       # (:195)
    return t.reshape(n, c, t.shape[2], t.shape[3])


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:86-94, CPU branch: the slope is hard-wired to 0.2 there (quirk #3)."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    return F.leaky_relu(x + bias.view(*shape), negative_slope=0.2) * scale


def fir_taps(k):
    """layers.py:7-12 `make_kernel`."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.dim() == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


# --------------------------------------------------------------------------- layers.py
class FusedLeakyReLU(nn.Module):      # op/fused_act.py:74-83
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope, self.scale = negative_slope, scale

    def forward(self, x):
        return fused_leaky_relu(x, self.bias, self.negative_slope, self.scale)


class PixelNorm(nn.Module):           # layers.py:100-105
    def forward(self, x):
        return x * torch.rsqrt(torch.mean(x * x, -1, keepdim=True) + 1e-8)


class Upsample(nn.Module):            # layers.py:170-181
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        taps = fir_taps(kernel) * factor ** 2
        self.register_buffer("kernel", taps)
        p = taps.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, x):
        return upfirdn2d(x, self.kernel, up=self.factor, pad=self.pad)


class Blur(nn.Module):                # layers.py:194-203
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        taps = fir_taps(kernel)
        if upsample_factor > 1:
            taps = taps * upsample_factor ** 2
        self.register_buffer("kernel", taps)
        self.pad = pad

    def forward(self, x):
        return upfirdn2d(x, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):         # layers.py:204-221
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(cout, cin, k, k))
        self.scale = 1 / math.sqrt(cin * k * k)
        self.stride, self.padding = stride, padding
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

    def forward(self, x):
        return F.conv2d(x, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)


class EqualLinear(nn.Module):         # layers.py:222-251 (activations used by the models only)
    def __init__(self, din, dout, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(dout, din).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(dout).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(din)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, x):
        if self.activation == "fused_lrelu":
            return fused_leaky_relu(F.linear(x, self.weight * self.scale), self.bias * self.lr_mul)
        assert self.activation is None
        return F.linear(x, self.weight * self.scale, bias=self.bias * self.lr_mul)


class ScaledLeakyReLU(nn.Module):     # layers.py:252-258
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return F.leaky_relu(x, negative_slope=self.negative_slope) * math.sqrt(2)


class ModulatedConv2d(nn.Module):     # layers.py:259-323 (weight-space formulation, grouped conv)
    def __init__(self, cin, cout, k, style_dim, demodulate=True, upsample=False, downsample=False,
                 blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel = k, cin, cout
        self.upsample, self.downsample = upsample, downsample
        if upsample:
            p = (len(blur_kernel) - 2) - (k - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        if downsample:
            p = (len(blur_kernel) - 2) + (k - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(cin * k * k)
        self.padding = k // 2
        self.weight = nn.Parameter(torch.randn(1, cout, cin, k, k))
        self.modulation = EqualLinear(style_dim, cin, bias_init=1)
        self.demodulate = demodulate

    def forward(self, x, style):
        b, cin, h, w = x.shape
        k, cout = self.kernel_size, self.out_channel
        s = self.modulation(style).view(b, 1, cin, 1, 1)
        wt = self.scale * self.weight * s                              # :296
        if self.demodulate:
            d = torch.rsqrt(wt.pow(2).sum([2, 3, 4]) + self.eps)       # :297-299
            wt = wt * d.view(b, cout, 1, 1, 1)
        if self.upsample:                                              # :301-310
            wt = wt.transpose(1, 2).reshape(b * cin, cout, k, k)
            y = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt, padding=0, stride=2, groups=b)
            return self.blur(y.view(b, cout, y.shape[2], y.shape[3]))
        wt = wt.view(b * cout, cin, k, k)
        if self.downsample:                                            # :311-317
            x = self.blur(x)
            y = F.conv2d(x.reshape(1, b * cin, x.shape[2], x.shape[3]), wt, padding=0, stride=2, groups=b)
        else:                                                          # :318-322
            y = F.conv2d(x.reshape(1, b * cin, h, w), wt, padding=self.padding, groups=b)
        return y.view(b, cout, y.shape[2], y.shape[3])


class NoiseInjection(nn.Module):      # layers.py:324-332
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            b, _, h, w = image.shape
            noise = image.new_empty(b, 1, h, w).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):       # layers.py:333-340
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, latent):
        return self.input.repeat(latent.shape[0], 1, 1, 1)


class ConvLayer(nn.Sequential):       # layers.py:341-378 (+ quirk #2: activate=False -> no activation)
    def __init__(self, cin, cout, k, downsample=False, blur_kernel=(1, 3, 3, 1), bias=True, activate="lrelu"):
        layers = []
        if downsample:
            p = (len(blur_kernel) - 2) + (k - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, k // 2
        layers.append(EqualConv2d(cin, cout, k, padding=self.padding, stride=stride, bias=bias))
        if activate == "lrelu":
            layers.append(FusedLeakyReLU(cout) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)


class ResBlock(nn.Module):            # layers.py:379-391
    def __init__(self, cin, cout, blur_kernel=(1, 3, 3, 1), downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(cin, cin, 3)
        self.conv2 = ConvLayer(cin, cout, 3, downsample=downsample)
        self.skip = ConvLayer(cin, cout, 1, downsample=downsample, activate=False, bias=False)

    def forward(self, x):
        return (self.conv2(self.conv1(x)) + self.skip(x)) / math.sqrt(2)


# --------------------------------------------------------------------------- model.py
class StyledConv(nn.Module):          # model.py:11-32
    def __init__(self, cin, cout, k, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1), demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(cin, cout, k, style_dim, upsample=upsample, blur_kernel=blur_kernel,
                                    demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(cout)

    def forward(self, x, style, noise=None):
        return self.activate(self.noise(self.conv(x, style), noise=noise))


class StyledMapConv(StyledConv):      # model.py:33-55
    def forward(self, x, style, stylemap, noise=None):
        y = self.conv(x, style)
        y = y * stylemap[:, :1] + stylemap[:, 1:2]                     # :50
        return self.activate(self.noise(y, noise=noise))


class ToRGB(nn.Module):               # model.py:56-69
    def __init__(self, cin, style_dim, upsample=True, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(cin, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, x, style, skip=None):
        y = self.conv(x, style) + self.bias
        return y if skip is None else y + self.upsample(skip)


CHANNELS = lambda m: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m,  # noqa: E731
                      512: 32 * m, 1024: 16 * m}                       # model.py:96-105


class Generator(nn.Module):           # model.py:71-187
    conv_cls = StyledConv

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01):
        super().__init__()
        self._common(size, style_dim, n_mlp, channel_multiplier, lr_mlp)
        self._blocks(style_dim, blur_kernel)

    def _common(self, size, style_dim, n_mlp, channel_multiplier, lr_mlp):   # model.py:88-124
        self.size, self.style_dim = size, style_dim
        self.style = nn.Sequential(PixelNorm(), *[
            EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp)])
        self.channels = CHANNELS(channel_multiplier)
        self.input = ConstantInput(self.channels[4])
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs, self.upsamples, self.to_rgbs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.noises = nn.Module()
        for i in range(self.num_layers):
            r = (i + 5) // 2
            self.noises.register_buffer("noise_%d" % i, torch.randn(1, 1, 2 ** r, 2 ** r))
        for i in range(3, self.log_size + 1):                            # first copy of the ToRGB list
            self.to_rgbs.append(ToRGB(self.channels[2 ** i], style_dim))
        self.n_latent = self.log_size * 2 - 2

    def _blocks(self, style_dim, blur_kernel):                           # model.py:75-87
        cin = self.channels[4]
        self.conv1 = self.conv_cls(cin, cin, 3, style_dim, blur_kernel=blur_kernel)
        for i in range(3, self.log_size + 1):
            cout = self.channels[2 ** i]
            self.convs.append(self.conv_cls(cin, cout, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(self.conv_cls(cout, cout, 3, style_dim, blur_kernel=blur_kernel))
            self._extra_block(i, style_dim)
            self.to_rgbs.append(ToRGB(cout, style_dim))                  # second, unused copy (quirk #4)
            cin = cout

    def _extra_block(self, i, style_dim):
        pass

    def _latents(self, styles, inject_index, truncation, truncation_latent, input_is_latent):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1 and truncation_latent is not None:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:                                              # model.py:156-162
            return styles[0].unsqueeze(1).repeat(1, self.n_latent, 1) if styles[0].dim() < 3 else styles[0]
        assert inject_index is not None, "oracle: style mixing needs an explicit inject_index"
        return torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                          styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)

    def _noise(self, noise, randomize_noise):
        if noise is not None:
            return noise
        if randomize_noise:
            return [None] * self.num_layers
        return [getattr(self.noises, "noise_%d" % i) for i in range(self.num_layers)]

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True):
        latent = self._latents(styles, inject_index, truncation, truncation_latent, input_is_latent)
        noise = self._noise(noise, randomize_noise)
        out = self.conv1(self.input(latent), latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        for j in range(self.log_size - 2):                               # model.py:173-181
            i = 1 + 2 * j
            out = self.convs[2 * j](out, latent[:, i], noise=noise[i])
            out = self.convs[2 * j + 1](out, latent[:, i + 1], noise=noise[i + 1])
            skip = self.to_rgbs[j](out, latent[:, i + 2], skip)
        return skip, (latent if return_latents else None)


class GeneratorWithMap(Generator):    # model.py:188-295, n_stylemap == 3 (the default and only config used)
    conv_cls = StyledMapConv

    def __init__(self, size, style_dim, n_mlp, n_stylemap=3, channel_multiplier=2, blur_kernel=(1, 3, 3, 1),
                 lr_mlp=0.01, rasterize=None):
        assert n_stylemap == 3
        nn.Module.__init__(self)
        self._common(size, style_dim, n_mlp, channel_multiplier, lr_mlp)
        self.norm_to_style = nn.ModuleList()
        self.norm1 = ResBlock(3, 2, downsample=False)
        self._blocks(style_dim, blur_kernel)
        self._rasterize = rasterize       # callable (v, tex, tri, h, w) -> [b,h,w,c]; injected by the caller

    def _extra_block(self, i, style_dim):
        self.norm_to_style.append(ResBlock(3, 4, downsample=False))

    def forward(self, styles, mesh, return_normals=False, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        latent = self._latents(styles, inject_index, truncation, truncation_latent, input_is_latent)
        noise = self._noise(noise, randomize_noise)
        out = self.input(latent)
        normals = [self._rasterize(mesh[0], mesh[1], mesh[2], out.shape[2], out.shape[3]).permute(0, 3, 1, 2)]
        out = self.conv1(out, latent[:, 0], self.norm1(normals[-1]), noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        for j in range(self.log_size - 2):                               # model.py:266-285
            i = 1 + 2 * j
            r = 2 * out.shape[2]
            normals.append(self._rasterize(mesh[0], mesh[1], mesh[2], r, r).permute(0, 3, 1, 2))
            maps = self.norm_to_style[j](normals[-1])
            out = self.convs[2 * j](out, latent[:, i], maps[:, :2], noise=noise[i])
            out = self.convs[2 * j + 1](out, latent[:, i + 1], maps[:, 2:], noise=noise[i + 1])
            skip = self.to_rgbs[j](out, latent[:, i + 2], skip)
        return skip, (latent if return_latents else None), (normals if return_normals else None)


class Discriminator(nn.Module):       # model.py:296-336
    def __init__(self, size, channel_multiplier=2, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        ch = CHANNELS(channel_multiplier)
        convs = [ConvLayer(3, ch[size], 1)]
        cin = ch[size]
        for i in range(int(math.log(size, 2)), 2, -1):
            convs.append(ResBlock(cin, ch[2 ** (i - 1)], blur_kernel))
            cin = ch[2 ** (i - 1)]
        self.convs = nn.Sequential(*convs)
        self.stddev_group, self.stddev_feat = 4, 1
        self.final_conv = ConvLayer(cin + 1, ch[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(ch[4] * 16, ch[4], activation="fused_lrelu"),
                                          EqualLinear(ch[4], 1))

    def forward(self, x):
        out = self.convs(x)
        b, c, h, w = out.shape
        g = min(b, self.stddev_group)
        sd = out.view(g, -1, self.stddev_feat, c // self.stddev_feat, h, w)
        sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdim=True).squeeze(2)
        out = torch.cat([out, sd.repeat(g, 1, h, w)], 1)
        return self.final_linear(self.final_conv(out).view(b, -1))
