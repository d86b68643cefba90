"""StyledConv / StyledMapConv / ToRGB / Generator / GeneratorWithMap / Discriminator behind the reference's
class names, constructor signatures, forward signatures and state_dict keys (reference model.py).

The state_dict key set is the persistence contract (reference train.py:412-420, generate.py:63-67): a
checkpoint written by the reference loads here unchanged, including the duplicated ToRGB list the
reference builds (model.py:86 + :122).
"""
import math

import numpy as np
import torch
from torch import nn

from . import fused, layers
from .layers import (ConstantInput, ConvLayer, EqualLinear, ModulatedConv2d, NoiseInjection, PixelNorm, ResBlock,
                     Upsample)
from .mesh import NormalMaps
from .op import FusedLeakyReLU, rasterize, rasterize_pyramid, rasterize_pyramid_maps
from .op.rasterize import MAX_LEVELS


class StyledConv(nn.Module):                          # reference model.py:11-32
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.bias = None
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        if (layers.get_conv_backend() == "tcgen05" and fused.supported(self.conv, input)
                and not (torch.is_grad_enabled() and layers.double_backward_requested())
                and not fused._noise_needs_grad(noise)):
            return fused.styled_conv(self.conv, self.noise, self.activate, input, style, noise)
        out = self.conv(input, style)
        out = self.noise(out, noise=noise)
        return self.activate(out)


class StyledMapConv(nn.Module):                       # reference model.py:33-55
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.bias = None
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, stylemap, noise=None):
        out = self.conv(input, style)
        if out.is_cuda and layers.double_backward_requested():
            # regulariser iterations (composed, twice differentiable): the map affine and the noise injection as ONE
            # full-size pass -- out * map0 + (map1 + weight * noise), the per-pixel term is a [B,1,H,W] tensor
            if noise is None:
                noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
            out = torch.addcmul(stylemap[:, 1:2] + self.noise.weight * noise, out, stylemap[:, :1])
            return self.activate(out)
        out = out * stylemap[:, :1] + stylemap[:, 1:2]                    # reference model.py:50
        out = self.noise(out, noise=noise)
        return self.activate(out)


class ToRGB(nn.Module):                               # reference model.py:56-69
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class Generator(nn.Module):                           # reference model.py:71-187
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self._initialize_styled(size, n_mlp, style_dim, channel_multiplier, lr_mlp)
        in_channel = self.channels[4]
        self.conv1 = StyledConv(in_channel, in_channel, 3, style_dim, blur_kernel=blur_kernel)
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))           # duplicate list entries, kept for key parity
            in_channel = out_channel

    def _initialize_styled(self, size, n_mlp, style_dim, channel_multiplier, lr_mlp):   # reference model.py:88-124
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        m = channel_multiplier
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m,
                         1024: 16 * m}
        self.input = ConstantInput(self.channels[4])
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer("noise_%d" % layer_idx, torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            self.to_rgbs.append(ToRGB(self.channels[2 ** i], style_dim))
        self.n_latent = self.log_size * 2 - 2

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def _prepare(self, styles, inject_index, truncation, truncation_latent, input_is_latent, noise, randomize_noise):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, "noise_%d" % i) for i in range(self.num_layers)]
        if truncation < 1 and truncation_latent is not None:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].dim() < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = np.random.choice(self.n_latent - 2) + 1     # reference model.py:167-168
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        return latent, noise

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True):
        latent, noise = self._prepare(styles, inject_index, truncation, truncation_latent, input_is_latent, noise,
                                      randomize_noise)
        if type(self) is Generator and latent.is_cuda and fused.chain_supported(self, self.input.input, noise):
            # chained tensor-core blocks: modulated operands handed from epilogue to GEMM, ToRGB fused (fused.py)
            skip = fused.generator_chain_forward(self, latent, noise)
            return (skip, latent) if return_latents else (skip, None)
        out = self.input(latent)
        out = self.conv1(out, latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                        self.to_rgbs):
            out = conv1(out, latent[:, i], noise=noise1)
            out = conv2(out, latent[:, i + 1], noise=noise2)
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        return (skip, latent) if return_latents else (skip, None)


class GeneratorWithMap(Generator):                    # reference model.py:188-295
    def __init__(self, size, style_dim, n_mlp, n_stylemap=3, channel_multiplier=2, blur_kernel=[1, 3, 3, 1],
                 lr_mlp=0.01):
        nn.Module.__init__(self)
        self._initialize_styled(size, n_mlp, style_dim, channel_multiplier, lr_mlp)
        self.norm_to_style = nn.ModuleList()
        in_channel = self.channels[4]
        if n_stylemap != 3:
            self.norm1 = nn.Sequential(ConvLayer(3, n_stylemap, 3), ResBlock(n_stylemap, 2, downsample=False))
        else:
            self.norm1 = ResBlock(n_stylemap, 2, downsample=False)
        self.conv1 = StyledMapConv(in_channel, in_channel, 3, style_dim, blur_kernel=blur_kernel)
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledMapConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                            blur_kernel=blur_kernel))
            self.convs.append(StyledMapConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            if n_stylemap != 3:
                self.norm_to_style.append(ConvLayer(3, n_stylemap, 3))
            self.norm_to_style.append(ResBlock(n_stylemap, 4, downsample=False))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel

    def _normal_maps(self, mesh):
        """The rasterised normal map at every resolution 4, 8, ..., size as [B,3,r,r] views (reference model.py:260-270
        calls rasterize once per resolution; here all of them come from one pyramid launch set, same values)."""
        sizes = [2 ** i for i in range(2, self.log_size + 1)]
        # already rendered by the fused front-end (mesh.normal_pyramid).  DistributedDataParallel rebuilds its inputs as
        # plain lists, so a list of [b,3,r,r] tensors is recognised by structure, not only by type
        if isinstance(mesh, NormalMaps) or (isinstance(mesh, (list, tuple)) and len(mesh) == len(sizes)
                                            and all(torch.is_tensor(m) and m.dim() == 4 for m in mesh)):
            assert [m.shape[-1] for m in mesh] == sizes, "NormalMaps must hold one map per resolution 4 .. size"
            return list(mesh)
        pyramid_ok = mesh[0].dtype == torch.float32 and len(sizes) <= MAX_LEVELS
        no_grad = not (torch.is_grad_enabled() and (mesh[0].requires_grad or mesh[1].requires_grad))
        if pyramid_ok and no_grad and mesh[0].is_cuda and mesh[1].dim() == mesh[0].dim():
            # forward-only: [b,3,r,r] planes straight from the resolve pass, no index / coefficient buffers
            return rasterize_pyramid_maps(mesh[0], mesh[1], mesh[2], sizes, planar=True)
        if pyramid_ok:
            maps = rasterize_pyramid(mesh[0], mesh[1], mesh[2], sizes)
        else:
            maps = [rasterize(mesh[0], mesh[1], mesh[2], r, r) for r in sizes]
        return [m.permute(0, 3, 1, 2) for m in maps]

    def forward(self, styles, mesh, return_normals=False, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        latent, noise = self._prepare(styles, inject_index, truncation, truncation_latent, input_is_latent, noise,
                                      randomize_noise)
        if latent.is_cuda and fused.chain_supported(self, self.input.input, noise):
            # chained tensor-core StyledMapConv blocks (fused.py); the rasterised normal maps and the style-map nets are
            # evaluated per resolution exactly as below and handed to the blocks' epilogues
            norm_maps, cache = [], {}

            normals = self._normal_maps(mesh)

            def maps_fn(k, h, w):
                j = (k + 1) // 2                                         # resolution index: block 0 -> 0, blocks 1,2 -> 1, ...
                if j not in cache:
                    nm = normals[j]
                    assert nm.shape[2] == h and nm.shape[3] == w
                    norm_maps.append(nm)
                    if j == 0:
                        cache[j] = self.norm1(nm)
                    elif len(self.convs) == len(self.norm_to_style):
                        i = 2 * j - 1
                        cache[j] = self.norm_to_style[i](self.norm_to_style[i - 1](nm))
                    else:
                        cache[j] = self.norm_to_style[j - 1](nm)
                m = cache[j] = cache[j].contiguous()                    # NCHW planes (the nets run on channels_last views)
                if j == 0:
                    return m
                return m[:, :2] if k % 2 == 1 else m[:, 2:]
            skip = fused.generator_chain_forward(self, latent, noise, maps_fn)
            return skip, (latent if return_latents else None), (norm_maps if return_normals else None)
        out = self.input(latent)
        normals = self._normal_maps(mesh)
        norm_maps = [normals[0]]
        maps = self.norm1(norm_maps[-1])
        out = self.conv1(out, latent[:, 0], maps, noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                        self.to_rgbs):
            norm_maps.append(normals[len(norm_maps)])
            if len(self.convs) == len(self.norm_to_style):              # reference model.py:271-275
                maps = self.norm_to_style[i](self.norm_to_style[i - 1](norm_maps[-1]))
            else:
                maps = self.norm_to_style[i // 2](norm_maps[-1])
            out = conv1(out, latent[:, i], maps[:, :2], noise=noise1)
            out = conv2(out, latent[:, i + 1], maps[:, 2:], noise=noise2)
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        return skip, (latent if return_latents else None), (norm_maps if return_normals else None)


class Discriminator(nn.Module):                       # reference model.py:296-336
    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        m = channel_multiplier
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m,
                    1024: 16 * m}
        convs = [ConvLayer(3, channels[size], 1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))

    def forward(self, input):
        out = self.convs(input)
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group)
        stddev = out.reshape(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdim=True).squeeze(2)
        stddev = stddev.repeat(group, 1, height, width)
        out = torch.cat([out, stddev], 1)
        out = self.final_conv(out)
        return self.final_linear(out.reshape(batch, -1))

