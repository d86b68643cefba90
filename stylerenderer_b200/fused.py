"""StyledConv as one differentiable block on the tcgen05 kernels (forward: modulate -> implicit-GEMM conv with the
demodulate/noise/bias/leaky-ReLU epilogue; backward: activation backward -> dgrad GEMM -> wgrad GEMM).

Replaces, for one StyledConv (reference model.py:26-32 + layers.py:293-323 + op/fused_act.py), the chain
  weight*style -> demod -> grouped conv [-> blur] -> + noise -> + bias, lrelu, *sqrt2
and its autograd graph.  Internal layout is NHWC; logical shapes stay NCHW (channels_last strides) so callers and
the state_dict are untouched.  The fused blocks (StyledConvTC, ModConvTC, StyledLayerTC, PlainConvTC) give first-order
gradients; iterations that differentiate through a backward pass (R1 / path-length regularisers, entered through
layers.double_backward()) use the mutually recursive ConvTC / ConvDgradTC / ConvWgradTC Functions below, which are
differentiable to any order on the same tensor-core kernels.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from . import style
from . import tc_conv as tc
from .op.upfirdn2d import upfirdn2d_raw


def _mode_operand(ctx, t):
    """A GEMM operand saved by the forward, as the backward's operand mode wants it.  When the forward ran in the
    fp32-faithful mode (tc_conv "tf32x3": saved operands are unrounded) and the backward runs in the shipped tf32 mode --
    the parity tests do that to check the tf32 backward kernels on the reference's own activations / leaky-ReLU masks --
    the operand gets the tf32 rounding its producer would have applied."""
    if t is not None and getattr(ctx, "exact_fwd", False) and not tc._exact():
        return tc.split_tf32(t)[0]
    return t


def _proxy(op_nhwc):
    """bf16 operand mode: the GEMM operand a block hands to the next one is a bfloat16 tensor, but the gradient that comes
    back for it (the next block's dgrad output) is fp32 -- autograd would cast a gradient to the dtype of the tensor it
    belongs to.  So the autograd edge is carried by this zero-storage fp32 stand-in of the operand's logical [N,C,H,W]
    shape, and the bfloat16 data travels beside it as a non-differentiable output."""
    b, h, w, c = op_nhwc.shape
    return op_nhwc.new_zeros(1, dtype=torch.float32).expand(b, c, h, w)


def to_nhwc(x):
    """logical [N,C,H,W] -> contiguous [N,H,W,C] (a free view when x is channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()


def from_nhwc(x):
    return x.permute(0, 3, 1, 2)


def supported(mod, x):
    """Shapes the fused tensor-core blocks take.  The up-sampling blocks hard-code the reference's default FIR geometry
    (4x4 taps, pad (1,1), reference layers.py:270-275 with blur_kernel of length 4): any other blur_kernel takes the
    composed path."""
    if mod.upsample and not (tuple(mod.blur.kernel.shape) == (4, 4) and tuple(mod.blur.pad) == (1, 1)):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and mod.kernel_size == 3 and not mod.downsample and mod.demodulate
            and tc.supported(mod.in_channel, mod.out_channel) and tc.wgrad_supported(mod.in_channel, mod.out_channel))


def _noise_needs_grad(noise):
    """The fused blocks return no gradient for the noise input (per-layer noise optimisation, StyleGAN2 projection):
    such calls take the composed path, which propagates it."""
    if noise is None:
        return False
    if isinstance(noise, (list, tuple)):
        return any(_noise_needs_grad(n) for n in noise)
    return torch.is_grad_enabled() and noise.requires_grad


class StyledConvTC(Function):
    @staticmethod
    def forward(ctx, x, weight, s, d, noise, noise_weight, act_bias, scale, upsample, blur_taps, alpha, gain):
        b, cin, h, w = x.shape
        cout = weight.shape[1]
        x_nhwc = to_nhwc(x)
        xs = tc.modulate(x_nhwc, s)                                    # tf32(x * s[b,c]): A operand of fwd and wgrad
        wk = tc.weight_prep(weight[0], scale, 0)
        s = s.contiguous()
        d = d.contiguous()
        t = None
        if not upsample:
            y = torch.empty(b, h, w, cout, dtype=torch.float32, device=x.device)
            tc.conv3x3(xs, wk, out=y, epilogue=1, rowscale=d, bias=act_bias, alpha=alpha, gain=gain,
                       noise=noise.reshape(-1, h, w).contiguous(), noise_weight=noise_weight)
        else:
            t = tc.conv_transpose3x3_s2(xs, wk, rowscale=d)            # [b, 2h+1, 2w+1, cout] = d * conv_T(xs)
            # NHWC FIR (pad 1,1) with the noise + bias + leaky-ReLU tail fused -> [b, 2h, 2w, cout]
            y = tc.blur_styled(t, blur_taps, (1, 1), noise, noise_weight, act_bias, alpha, gain)
        ctx.save_for_backward(x_nhwc, xs, y, t, weight, s, d, noise, noise_weight, act_bias, blur_taps)
        ctx.cfg = (scale, upsample, alpha, gain)
        ctx.exact_fwd = tc._exact()
        return from_nhwc(y)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x_nhwc, xs, y, t, weight, s, d, noise, noise_weight, act_bias, blur_taps = ctx.saved_tensors
        xs = _mode_operand(ctx, xs)
        scale, upsample, alpha, gain = ctx.cfg
        b, h, w, cin = x_nhwc.shape
        cout = y.shape[3]
        oh, ow = y.shape[1], y.shape[2]
        gy = to_nhwc(gy)
        if not upsample:
            # one pass: activation backward, bias / noise-weight gradients, e = sum g_pre * (d * acc), ga = tf32(g_pre * d)
            ga, g_bias, g_noise_w, e = tc.bwd_prologue(gy, y, noise, noise_weight, act_bias, d, alpha, gain, True)
            dxs = tc.conv3x3(ga, tc.weight_prep(weight[0], scale, 1))
            dwk = tc.wgrad3x3(ga, xs) if ctx.needs_input_grad[1] else None
        else:
            g_pre, g_bias, g_noise_w, _ = tc.bwd_prologue(gy, y, noise, noise_weight, act_bias, None, alpha, gain, False)
            ga, e = tc.blur_scaledot(g_pre, torch.flip(blur_taps, [0, 1]), (2, 2), d, t)   # FIR^T, * d, tf32, sum gt * t
            dxs = tc.conv3x3_s2_gather(ga, tc.weight_prep(weight[0], scale, 2), (h, w))
            dwk = tc.wgrad_transpose3x3_s2(ga, xs) if ctx.needs_input_grad[1] else None
        g_d = e / d                                                     # dL/dd = sum g * acc = e / d
        g_x, g_s = tc.scale_dot(dxs, x_nhwc, s, False)                 # dx = dxs * s, ds = sum_p dxs * x
        g_x = from_nhwc(g_x)
        g_w = (dwk.view(cout, 3, 3, cin).permute(0, 3, 1, 2) * scale).unsqueeze(0) if dwk is not None else None
        return g_x, g_w, g_s, g_d, None, g_noise_w, g_bias, None, None, None, None, None


class ModConvTC(Function):
    """The bare ModulatedConv2d contraction y = d[b,o] * conv(x * s[b,i], scale * W) (plain 3x3 or the stride-2
    transposed 3x3 + FIR of reference layers.py:301-310) on the tensor-core kernels, for callers that apply their own
    tail (StyledMapConv, user code)."""

    @staticmethod
    def forward(ctx, x, weight, s, d, scale, upsample, blur_taps):
        x_nhwc = to_nhwc(x)
        s, d = s.contiguous(), d.contiguous()
        xs = tc.modulate(x_nhwc, s)
        wk = tc.weight_prep(weight[0], scale, 0)
        if not upsample:
            saved = y = tc.conv3x3(xs, wk, rowscale=d)
        else:
            saved = tc.conv_transpose3x3_s2(xs, wk, rowscale=d)
            y = upfirdn2d_raw(saved, blur_taps, 1, 1, 1, 1, 1, 1, 1, 1)
        ctx.save_for_backward(x_nhwc, xs, saved, weight, s, d, blur_taps)
        ctx.cfg = (scale, upsample)
        ctx.exact_fwd = tc._exact()
        return from_nhwc(y)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x_nhwc, xs, saved, weight, s, d, blur_taps = ctx.saved_tensors
        xs = _mode_operand(ctx, xs)
        scale, upsample = ctx.cfg
        b, h, w, cin = x_nhwc.shape
        cout = saved.shape[3]
        gy = to_nhwc(gy)
        if not upsample:
            ga, e = tc.scale_dot(gy, saved, d, True)                   # tf32(gy * d), e = sum gy * (d * acc)
            dxs = tc.conv3x3(ga, tc.weight_prep(weight[0], scale, 1))
            dwk = tc.wgrad3x3(ga, xs) if ctx.needs_input_grad[1] else None
        else:
            ga, e = tc.blur_scaledot(gy, torch.flip(blur_taps, [0, 1]), (2, 2), d, saved)
            dxs = tc.conv3x3_s2_gather(ga, tc.weight_prep(weight[0], scale, 2), (h, w))
            dwk = tc.wgrad_transpose3x3_s2(ga, xs) if ctx.needs_input_grad[1] else None
        g_x, g_s = tc.scale_dot(dxs, x_nhwc, s, False)
        g_w = (dwk.view(cout, 3, 3, cin).permute(0, 3, 1, 2) * scale).unsqueeze(0) if dwk is not None else None
        return from_nhwc(g_x), g_w, g_s, e / d, None, None, None


# ---------------------------------------------------------------------------------------------------------------------
# Twice (arbitrarily often) differentiable convolutions on the tensor-core kernels, for the regulariser iterations
# (R1 / path length, reference train.py:110-134) that differentiate through a backward pass.  The three bilinear maps
#   conv(x, w),  dgrad(g, w),  wgrad(g, x)
# are autograd Functions whose backward passes are built from each other, exactly like torch's own convolution
# double-backward; every one of them is one launch of the implicit-GEMM / wgrad kernels on tf32-rounded operands.
# kind 'plain': 3x3, stride 1, pad 1;  kind 'up': stride-2 transposed 3x3 (reference layers.py:301-309), output (2H+1)^2;
# kinds 'down' / 'down1': 3x3 / 1x1, stride 2, pad 0 (the Discriminator's down-sampling ConvLayers after their Blur).
def _conv_fwd(x, w, kind):
    xr, wk = tc.modulate(x), tc.weight_prep(w, 1.0, 0)
    if kind == "plain":
        return tc.conv3x3(xr, wk)
    if kind == "up":
        return tc.conv_transpose3x3_s2(xr, wk)
    b, h, wd, _ = x.shape                                       # 'down' (3x3 s2 p0) / 'down1' (1x1 s2 p0)
    k = w.shape[2]
    y = torch.empty(b, (h - k) // 2 + 1, (wd - k) // 2 + 1, w.shape[0], dtype=torch.float32, device=x.device)
    return tc.conv_igemm(xr, wk, TAPS_S2 if k == 3 else [(0, 0, 0)], y, in_stride=2)


def _conv_dgrad(g, w, kind, hw):
    gr = tc.modulate(g)
    if kind == "plain":
        return tc.conv3x3(gr, tc.weight_prep(w, 1.0, 1))
    if kind == "up":
        return tc.conv3x3_s2_gather(gr, tc.weight_prep(w, 1.0, 2), hw)
    wt = tc.weight_prep(w, 1.0, 2)
    if kind == "down":
        assert hw == (2 * g.shape[1] + 1, 2 * g.shape[2] + 1)
        return tc.conv_transpose3x3_s2(gr, wt)
    dx = torch.zeros(g.shape[0], hw[0], hw[1], w.shape[1], dtype=torch.float32, device=g.device)
    return tc.conv_igemm_multi(gr, wt, [([(0, 0, 0)], (g.shape[1], g.shape[2]), (0, 0))], dx, out_stride=2)


def _conv_wgrad(g, x, kind, k=3):
    gr, xr = tc.modulate(g), tc.modulate(x)
    if kind == "plain":
        dwk = tc.wgrad3x3(gr, xr)
    elif kind == "up":
        dwk = tc.wgrad_transpose3x3_s2(gr, xr)
    elif kind == "down":
        dwk = tc.wgrad(gr, xr, [(0, 0, ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)], (g.shape[1], g.shape[2]),
                       x_stride=2)
    else:
        dwk = tc.wgrad(gr, xr, [(0, 0, 0, 0, 0)], (g.shape[1], g.shape[2]), x_stride=2, taps_total=1)
        k = 1
    return style.weight_grad_layout(dwk, 1.0, g.shape[3], x.shape[3], k)[0]


class ConvTC(Function):
    """y = conv(x, w): x [B,H,W,Cin] NHWC, w [Cout,Cin,3,3] -> [B,H',W',Cout]."""

    @staticmethod
    def forward(ctx, x, w, kind):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        ctx.kind = kind
        return _conv_fwd(x, w, kind)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = ConvDgradTC.apply(gy, w, ctx.kind, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[0] else None
        gw = ConvWgradTC.apply(gy, x, ctx.kind) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class ConvDgradTC(Function):
    """gx = dgrad(g, w): the gradient of conv(x, w) w.r.t. x for the output gradient g."""

    @staticmethod
    def forward(ctx, g, w, kind, hw):
        g, w = g.contiguous(), w.contiguous()
        ctx.save_for_backward(g, w)
        ctx.kind = kind
        return _conv_dgrad(g, w, kind, hw)

    @staticmethod
    def backward(ctx, ggx):
        g, w = ctx.saved_tensors
        ggx = ggx.contiguous()
        d_g = ConvTC.apply(ggx, w, ctx.kind) if ctx.needs_input_grad[0] else None
        d_w = ConvWgradTC.apply(g, ggx, ctx.kind) if ctx.needs_input_grad[1] else None
        return d_g, d_w, None, None


class ConvWgradTC(Function):
    """gw = wgrad(g, x) [Cout,Cin,3,3]: the gradient of conv(x, w) w.r.t. w for the output gradient g."""

    @staticmethod
    def forward(ctx, g, x, kind):
        g, x = g.contiguous(), x.contiguous()
        ctx.save_for_backward(g, x)
        ctx.kind = kind
        return _conv_wgrad(g, x, kind)

    @staticmethod
    def backward(ctx, ggw):
        g, x = ctx.saved_tensors
        ggw = ggw.contiguous()
        d_g = ConvTC.apply(x, ggw, ctx.kind) if ctx.needs_input_grad[0] else None
        d_x = ConvDgradTC.apply(g, ggw, ctx.kind, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[1] else None
        return d_g, d_x, None


_DD_KIND = {"s1": "plain", "s2": "down", "p2": "down1"}


def plain_conv_dd(conv, act, x, kind):
    """ConvLayer body for iterations that need double backward (R1, reference train.py:110-114): the contraction is the
    twice-differentiable ConvTC, bias / activation are the (twice differentiable) fused_leaky_relu op."""
    from .op import fused_leaky_relu
    y = ConvTC.apply(x.permute(0, 2, 3, 1).contiguous(), conv.weight * conv.scale, _DD_KIND[kind]).permute(0, 3, 1, 2)
    if act is None:
        return y
    bias = act.bias if conv.bias is None else act.bias + conv.bias
    return fused_leaky_relu(y, bias, act.negative_slope, act.scale)


class ScaleBC(Function):
    """y[b,h,w,c] = a[b,h,w,c] * s[b,c] on channels-last tensors (sr_scale_dot_nhwc_f32).  With DotP (r[b,c] = sum over pixels
    of a * o) and ScaleDot (both in one pass) it is closed under differentiation -- the modulation / demodulation multiplies
    of ModulatedConv2d (reference layers.py:296-299, here in activation form) for the regulariser iterations that
    differentiate twice: torch's `x * s.view(b,c,1,1)` costs a multiply, and per differentiation two more multiplies with a
    full-size temporary plus a reduction pass."""

    @staticmethod
    def forward(ctx, a, s):
        a, s = a.contiguous(), s.contiguous()
        ctx.save_for_backward(a, s)
        return tc.scale_dot(a, None, s, False)[0]

    @staticmethod
    def backward(ctx, g):
        a, s = ctx.saved_tensors
        need_a, need_s = ctx.needs_input_grad
        if need_a and need_s:
            return ScaleDot.apply(g, a, s)
        if need_a:
            return ScaleBC.apply(g, s), None
        return None, (DotP.apply(g, a) if need_s else None)


class DotP(Function):
    """r[b,c] = sum_{h,w} a[b,h,w,c] * o[b,h,w,c] (one read of both, no full-size temporary)."""

    @staticmethod
    def forward(ctx, a, o):
        a, o = a.contiguous(), o.contiguous()
        ctx.save_for_backward(a, o)
        return tc.scale_dot(a, o, None, False, want_out=False)[1]

    @staticmethod
    def backward(ctx, g):
        a, o = ctx.saved_tensors
        return (ScaleBC.apply(o, g) if ctx.needs_input_grad[0] else None,
                ScaleBC.apply(a, g) if ctx.needs_input_grad[1] else None)


class ScaleDot(Function):
    """(g * s[b,c], sum_{h,w} g * a) in ONE pass over g and a: the two gradients of ScaleBC."""

    @staticmethod
    def forward(ctx, g, a, s):
        g, a, s = g.contiguous(), a.contiguous(), s.contiguous()
        ctx.save_for_backward(g, a, s)
        ctx.set_materialize_grads(False)
        return tc.scale_dot(g, a, s, False)

    @staticmethod
    def backward(ctx, go, gd):
        g, a, s = ctx.saved_tensors
        gg = ga = gs = None
        if go is not None:                               # out = g * s
            gg, gs = ScaleBC.apply(go, s), DotP.apply(go, g)
        if gd is not None:                               # dot = sum g * a
            t = ScaleBC.apply(a, gd)
            gg = t if gg is None else gg + t
            ga = ScaleBC.apply(g, gd)
        return gg, ga, gs


def mod_conv_dd(mod, x, style):
    """ModulatedConv2d.forward for iterations that need double backward: modulation / demodulation through the ScaleBC /
    DotP pair (closed under differentiation), the contraction through ConvTC on the tensor cores."""
    s, d = mod.style_scales(style)
    xs = ScaleBC.apply(to_nhwc(x), s)
    y = ConvTC.apply(xs, mod.weight[0] * mod.scale, "up" if mod.upsample else "plain")
    if mod.upsample:
        y = to_nhwc(mod.blur(from_nhwc(y)))
    if d is not None:
        y = ScaleBC.apply(y, d)
    return from_nhwc(y)


_ONES = {}


def _ones(b, c, device):
    key = (b, c, str(device))
    if key not in _ONES:
        _ONES[key] = torch.ones(b, c, dtype=torch.float32, device=device)
    return _ONES[key]


TAPS_S2 = [(ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]


def _operand_of(x, x_nhwc):
    """The GEMM operand of an activation: the copy its producer already wrote (ResidualCombine tags its output with it), or
    one operand pass (sr_modulate_*)."""
    tag = getattr(x, "_sr_operand", None)
    if tag is not None and tag[0] == tc.get_precision() and tag[1].shape == x_nhwc.shape and tag[1].device == x_nhwc.device:
        return tag[1]
    return tc.modulate(x_nhwc)


class ResidualCombine(Function):
    """(a + b) * scale on channels-last activations in one pass, with the GEMM-operand copy of the result for the next
    convolution (sr_residual_combine_*): the tail of the Discriminator's ResBlock (reference layers.py:386-391)."""

    @staticmethod
    def forward(ctx, a, b, scale):
        a_n, b_n = to_nhwc(a), to_nhwc(b)
        out = torch.empty_like(a_n)
        op = tc.operand_like(a_n)
        fn = _lib.lib().sr_residual_combine_bf16 if tc._bf16() else _lib.lib().sr_residual_combine_tf32
        with torch.cuda.device(a.device):
            rc = fn(_lib.ptr(out), _lib.ptr(op), _lib.ptr(a_n), _lib.ptr(b_n), float(scale), a_n.numel(), _lib.stream_of(a_n))
        _lib.check(rc, "sr_residual_combine")
        ctx.scale = float(scale)
        y = from_nhwc(out)
        ctx.mark_non_differentiable(op)
        return y, op

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _gop):
        gs = g * ctx.scale
        return gs, gs, None


def residual_combine_supported(a, b):
    import os
    from .layers import double_backward_requested
    return (os.environ.get("SR_RES_COMBINE", "1") != "0" and a.is_cuda and a.dtype == torch.float32 and a.shape == b.shape
            and a.dim() == 4 and a.shape[1] % 4 == 0 and not tc._exact()
            and not (torch.is_grad_enabled() and double_backward_requested()))


def residual_combine(a, b, scale):
    y, op = ResidualCombine.apply(a, b, scale)
    y._sr_operand = (tc.get_precision(), op)          # read by _operand_of in the next block's first convolution
    return y


class PlainConvTC(Function):
    """EqualConv2d (+ FusedLeakyReLU) of the Discriminator / ConvLayer stack (reference layers.py:204-221, 341-378) on the
    tensor-core kernels: kind 's1' = 3x3 stride 1 pad 1, 's2' = 3x3 stride 2 pad 0 (after the Blur), 'p2' = 1x1 stride 2.
    Same implicit-GEMM kernels as the modulated convolution (no style: the operand is only rounded to tf32), bias +
    leaky-ReLU in the conv epilogue, activation backward + bias gradient + tf32 operand in one prologue pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, kind, alpha, gain):
        x_nhwc = to_nhwc(x)
        return from_nhwc(_plain_conv_forward(ctx, _operand_of(x, x_nhwc), weight, bias, scale, kind, alpha, gain))

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        dx, g_w, g_bias = _plain_conv_backward(ctx, gy)
        return from_nhwc(dx), g_w, g_bias, None, None, None, None


def _plain_conv_forward(ctx, xr, weight, bias, scale, kind, alpha, gain):
    """Shared forward of PlainConvTC / BlurConvTC from the GEMM operand xr [B,H,W,Cin] (tf32-rounded fp32 or bf16)."""
    b, h, w, cin = xr.shape
    cout, _, k, _ = weight.shape
    wk_f, wk_t, _ = tc.weight_prep_dual(weight, scale, kind == "s1", want_wsq=False)
    act = bias is not None
    kw = dict(epilogue=1, bias=bias, alpha=alpha, gain=gain) if act else dict(epilogue=0)
    if kind == "s1":
        y = tc.conv3x3(xr, wk_f, **kw)
    else:
        oh, ow = (h - k) // 2 + 1, (w - k) // 2 + 1
        y = torch.empty(b, oh, ow, cout, dtype=torch.float32, device=xr.device)
        taps = TAPS_S2 if k == 3 else [(0, 0, 0)]
        tc.conv_igemm(xr, wk_f, taps, y, in_stride=2, **kw)
    ctx.save_for_backward(xr, y if act else None, wk_t, bias)
    ctx.cfg = (scale, kind, alpha, gain, k, cin, cout, (h, w))
    ctx.exact_fwd = tc._exact()
    return y


def _plain_conv_backward(ctx, gy):
    """-> (dx [B,H,W,Cin] fp32 w.r.t. the conv's input, g_weight or None, g_bias or None)."""
    xr, y, wk_t, bias = ctx.saved_tensors
    xr, wk_t = _mode_operand(ctx, xr), _mode_operand(ctx, wk_t)
    scale, kind, alpha, gain, k, cin, cout, (h, w) = ctx.cfg
    gy = to_nhwc(gy)
    b, oh, ow, _ = gy.shape
    g_bias = None
    if y is not None:       # activation backward + bias gradient + tf32 operand in one pass (d = 1)
        ga, g_bias, _, _ = tc.bwd_prologue(gy, y, None, None, bias, _ones(b, cout, gy.device), alpha, gain, False)
    else:
        ga = tc.modulate(gy)
    need_w = ctx.needs_input_grad[1]          # False in the generator's step of the GAN loop (reference train.py:292-293)
    dwk = None
    if kind == "s1":
        dx = tc.conv3x3(ga, wk_t)
        if need_w:
            dwk = tc.wgrad3x3(ga, xr)
    elif kind == "s2":
        dx = tc.conv_transpose3x3_s2(ga, wk_t)
        if need_w:
            dwk = tc.wgrad(ga, xr, [(0, 0, ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)], (oh, ow),
                           x_stride=2)
    else:
        dx = torch.zeros(b, h, w, cin, dtype=torch.float32, device=gy.device)
        tc.conv_igemm_multi(ga, wk_t, [([(0, 0, 0)], (oh, ow), (0, 0))], dx, out_stride=2)
        if need_w:
            dwk = tc.wgrad(ga, xr, [(0, 0, 0, 0, 0)], (oh, ow), x_stride=2, taps_total=1)
    g_w = style.weight_grad_layout(dwk, scale, cout, cin, k)[0] if need_w else None
    return dx, g_w, g_bias


class BlurConvTC(Function):
    """Blur -> EqualConv2d (+ FusedLeakyReLU): the down-sampling ConvLayers of the Discriminator's ResBlocks (reference
    layers.py:346-351, 379-391).  The 4x4 FIR writes the GEMM OPERAND (tf32-rounded / bf16) directly -- no fp32 blurred tensor,
    no separate operand pass (6 instead of 14 bytes per element around the blur) -- and the backward runs the FIR's adjoint
    (flipped taps, pads 3 - p) on the dgrad output inside the same Function."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, kind, alpha, gain, taps, pad):
        x_nhwc = to_nhwc(x)
        b, _, _, cin = x_nhwc.shape
        xr, _ = tc.blur_scaledot(x_nhwc, taps, pad, _ones(b, cin, x.device))
        ctx.blur = (taps, pad)
        return from_nhwc(_plain_conv_forward(ctx, xr, weight, bias, scale, kind, alpha, gain))

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        dxb, g_w, g_bias = _plain_conv_backward(ctx, gy)
        taps, (p0, p1) = ctx.blur
        dx = None
        if ctx.needs_input_grad[0]:
            kh = taps.shape[0]
            dx = upfirdn2d_raw(dxb, torch.flip(taps, [0, 1]), 1, 1, 1, 1, kh - 1 - p0, kh - 1 - p1, kh - 1 - p0, kh - 1 - p1)
        return (from_nhwc(dx) if dx is not None else None), g_w, g_bias, None, None, None, None, None, None


def env_blur_conv():
    """A/B switch: SR_BLUR_CONV=0 keeps Blur and the convolution as two autograd nodes (round-1 form)."""
    import os
    return os.environ.get("SR_BLUR_CONV", "1") != "0"


def blur_conv_supported(conv, blur, x):
    """Shapes BlurConvTC takes: a 4x4 FIR in front of a tensor-core-supported stride-2 EqualConv2d."""
    if blur.kernel.shape != (4, 4) or not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        return None
    p0, p1 = blur.pad
    bh, bw = x.shape[2] + p0 + p1 - 3, x.shape[3] + p0 + p1 - 3
    if bh < 1 or bw < 1:
        return None
    k = conv.weight.shape[2]
    cout, cin = conv.weight.shape[:2]
    kind = {(3, 2, 0): "s2", (1, 2, 0): "p2"}.get((k, conv.stride, conv.padding))
    if kind is None or not (tc.supported(cin, cout) and tc.supported(cout, cin)):
        return None
    if kind == "s2" and (bh % 2 == 0 or bw % 2 == 0):
        return None
    return kind


def blur_conv(conv, act, blur, x, kind):
    if act is not None:
        bias = act.bias if conv.bias is None else act.bias + conv.bias
        return BlurConvTC.apply(x, conv.weight, bias, conv.scale, kind, act.negative_slope, act.scale, blur.kernel, tuple(blur.pad))
    return BlurConvTC.apply(x, conv.weight, None, conv.scale, kind, 0.2, 1.0, blur.kernel, tuple(blur.pad))


def plain_conv_supported(conv, x):
    """EqualConv2d shapes the tensor-core path takes: 3x3 s1 p1, 3x3 s2 p0, 1x1 s2 p0 with channels in multiples of 128."""
    k = conv.weight.shape[2]
    cout, cin = conv.weight.shape[:2]
    kind = {(3, 1, 1): "s1", (3, 2, 0): "s2", (1, 2, 0): "p2"}.get((k, conv.stride, conv.padding))
    # forward GEMM: N = cout, dgrad GEMM: N = cin -> both multiples of 128 (which is also what the wgrad kernels need)
    ok = tc.supported(cin, cout) and tc.supported(cout, cin)
    if kind is None or not (x.is_cuda and x.dtype == torch.float32 and ok):
        return None
    if kind == "s2" and (x.shape[2] % 2 == 0 or x.shape[3] % 2 == 0):      # dgrad = transposed conv: needs odd (2m+1) inputs
        return None
    if kind == "p2" and (x.shape[2] < 1 or x.shape[3] < 1):
        return None
    return kind


def plain_conv(conv, act, x, kind):
    """ConvLayer body (EqualConv2d [+ FusedLeakyReLU]) on the tensor cores; `act` is the FusedLeakyReLU module or None."""
    if act is not None:      # this fork keeps a bias in the conv AND in the activation (reference layers.py:365-375): they add up
        bias = act.bias if conv.bias is None else act.bias + conv.bias
        return PlainConvTC.apply(x, conv.weight, bias, conv.scale, kind, act.negative_slope, act.scale)
    return PlainConvTC.apply(x, conv.weight, None, conv.scale, kind, 0.2, 1.0)


def mod_conv(mod, x, style):
    """ModulatedConv2d.forward on the tensor cores (see ModConvTC)."""
    s, d = mod.style_scales(style)
    taps = mod.blur.kernel if mod.upsample else s
    return ModConvTC.apply(x, mod.weight, s, d, mod.scale, mod.upsample, taps)


def styled_conv(mod_conv, noise_mod, act_mod, x, style, noise):
    """The tcgen05 fast path of StyledConv.forward (mod_conv: ModulatedConv2d, act_mod: FusedLeakyReLU)."""
    s, d = mod_conv.style_scales(style)
    b, _, h, w = x.shape
    oh, ow = (2 * h, 2 * w) if mod_conv.upsample else (h, w)
    if noise is None:
        noise = x.new_empty(b, 1, oh, ow).normal_()
    taps = mod_conv.blur.kernel if mod_conv.upsample else noise_mod.weight
    return StyledConvTC.apply(x, mod_conv.weight, s, d, noise, noise_mod.weight, act_mod.bias, mod_conv.scale,
                              mod_conv.upsample, taps, act_mod.negative_slope, act_mod.scale)


# ------------------------------------------------------------------------------------------------------------------
# Chained generator: consecutive StyledConv blocks hand each other the ALREADY MODULATED tf32 operand (written by the
# producer's epilogue), ToRGB rides the conv epilogue, and in the backward the gradient of a block is assembled inside
# its prologue pass from the next block's dgrad output and the ToRGB gradient -- no stand-alone modulate / scale / 1x1
# conv / gradient-accumulation passes remain.
class ModulateTC(Function):
    """xs = tf32(x * s[b,c]) as an autograd node (used once, for the constant 4x4 input)."""

    @staticmethod
    def forward(ctx, x, s):
        x_nhwc = to_nhwc(x)
        s = s.contiguous()
        ctx.save_for_backward(x_nhwc, s)
        xs = tc.modulate(x_nhwc, s)
        if xs.dtype != torch.float32:                    # bf16 operand mode: (fp32 stand-in for autograd, operand data)
            ctx.mark_non_differentiable(xs)
            return _proxy(xs), xs
        return from_nhwc(xs), None

    @staticmethod
    @once_differentiable
    def backward(ctx, gxs, _g_op=None):
        x_nhwc, s = ctx.saved_tensors
        gx, gs = tc.scale_dot(to_nhwc(gxs), x_nhwc, s, False)
        return from_nhwc(gx), gs


class StyledLayerTC(Function):
    """One StyledConv / StyledMapConv block of the chain.  Input: xs (already x * s, tf32).  Outputs: (main, rgb) with
    main = tf32(y * s_next) if s_next is given else y, rgb = sum_c y * rgb_weight (or None).  With `stylemap` [B,2,H,W]
    the block is a StyledMapConv (reference model.py:33-55): t = conv_d * map0 + map1 + noise + bias."""

    @staticmethod
    def forward(ctx, xs, weight, d, noise, noise_weight, act_bias, s_next, rgb_weight, scale, upsample, blur_taps, alpha,
                gain, wk=None, wkt=None, stylemap=None, xs_op=None):
        ctx.set_materialize_grads(False)                     # unused outputs arrive as None, not as zero tensors
        # xs_op: the bfloat16 operand [B,H,W,C] in bf16 operand mode (xs is then its fp32 stand-in, see _proxy)
        xs_nhwc = xs_op if xs_op is not None else to_nhwc(xs)
        b, h, w, cin = xs_nhwc.shape
        cout = weight.shape[1]
        d = d.contiguous()
        if wk is None:                                       # operands prepared by the caller (style.WeightPrepAll) or here
            wk = tc.weight_prep(weight[0], scale, 0)
        ctx.wkt = wkt
        s_next = s_next.contiguous() if s_next is not None else None
        rgb_weight = rgb_weight.contiguous() if rgb_weight is not None else None
        t = rgb = y2 = None
        if not upsample:
            y = torch.empty(b, h, w, cout, dtype=torch.float32, device=xs.device)
            y2 = tc.operand_like(y) if s_next is not None else None
            rgb = torch.empty(b, h, w, 3, dtype=torch.float32, device=xs.device) if rgb_weight is not None else None
            # StyledMapConv: `y` receives the demodulated conv output before the map affine (epilogue 2); the backward
            # prologue rebuilds the activated value from it -- a map that is exactly 0 (background of the rasterised
            # normals at default init) makes it unrecoverable from the activated one
            tc.conv3x3(xs_nhwc, wk, out=y, epilogue=2 if stylemap is not None else 1, rowscale=d, bias=act_bias, alpha=alpha,
                       gain=gain, noise=noise.reshape(-1, h, w).contiguous(), noise_weight=noise_weight, out2=y2,
                       scale2=s_next, rgb_weight=rgb_weight, rgb_out=rgb, stylemap=stylemap)
        else:
            assert rgb_weight is None
            t = tc.conv_transpose3x3_s2(xs_nhwc, wk, rowscale=d)
            if s_next is not None:
                y, y2 = tc.blur_styled(t, blur_taps, (1, 1), noise, noise_weight, act_bias, alpha, gain, scale2=s_next,
                                       stylemap=stylemap)
            else:
                y = tc.blur_styled(t, blur_taps, (1, 1), noise, noise_weight, act_bias, alpha, gain, stylemap=stylemap)
        ctx.save_for_backward(xs_nhwc, y, None, weight, d, noise, noise_weight, act_bias, blur_taps, s_next, rgb_weight, stylemap)
        ctx.cfg = (scale, upsample, alpha, gain)
        ctx.exact_fwd = tc._exact()
        op_out = None
        if s_next is None and stylemap is not None:          # y holds the pre-map value: no activated output to hand on
            main = xs.new_zeros(1)
            ctx.mark_non_differentiable(main)
        elif s_next is not None and y2.dtype != torch.float32:
            main, op_out = _proxy(y2), y2                    # bf16 operand mode
            ctx.mark_non_differentiable(op_out)
        else:
            main = from_nhwc(y2 if s_next is not None else y)
        if rgb is None:
            rgb = xs.new_zeros(1)
            ctx.mark_non_differentiable(rgb)
        return main, rgb, op_out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_main, g_rgb, _g_op=None):
        xs, y, t, weight, d, noise, noise_weight, act_bias, blur_taps, s_next, rgb_weight, stylemap = ctx.saved_tensors
        xs, wkt = _mode_operand(ctx, xs), _mode_operand(ctx, ctx.wkt)
        scale, upsample, alpha, gain = ctx.cfg
        b, h, w, cin = xs.shape
        cout = y.shape[3]
        g_main = to_nhwc(g_main) if g_main is not None else None
        src = dict(gy=None, gxs=None, s_next=None, g_rgb=None, rgb_weight=None)
        if g_main is not None:
            if s_next is not None:
                src.update(gxs=g_main, s_next=s_next)
            else:
                src.update(gy=g_main)
        if rgb_weight is not None and g_rgb is not None:
            src.update(g_rgb=g_rgb.contiguous(), rgb_weight=rgb_weight)
        g_map = None
        # frozen weights (latent inversion): the weight-gradient GEMM (a third of the backward
        # flops) and its layout pass are skipped; every other gradient is unchanged
        need_w = ctx.needs_input_grad[1]
        if not upsample:
            res = tc.bwd_prologue2(y, noise, noise_weight, act_bias, d, alpha, gain, True, stylemap=stylemap, **src)
            ga, g_bias, g_noise_w, e, ds_next, dwb = res[:6]
            g_map = res[6] if stylemap is not None else None
            dxs = tc.conv3x3(ga, wkt if wkt is not None else tc.weight_prep(weight[0], scale, 1))
            dwk = tc.wgrad3x3(ga, xs) if need_w else None
        else:
            # e = sum_p gp * (fir(t) * map0) comes from the prologue (fir(t) * map0 is recoverable from y, like the plain
            # block's conv output), so the FIR^T pass does not have to read the saved transposed-conv output t again
            res = tc.bwd_prologue2(y, noise, noise_weight, act_bias, None, alpha, gain, True, stylemap=stylemap, **src)
            g_pre, g_bias, g_noise_w, e, ds_next, dwb = res[:6]
            g_map = res[6] if stylemap is not None else None
            ga, _ = tc.blur_scaledot(g_pre, torch.flip(blur_taps, [0, 1]), (2, 2), d)       # FIR^T, * d, tf32
            dxs = tc.conv3x3_s2_gather(ga, wkt if wkt is not None else tc.weight_prep(weight[0], scale, 2), (h, w))
            dwk = tc.wgrad_transpose3x3_s2(ga, xs) if need_w else None
        g_d = e / d
        g_w = style.weight_grad_layout(dwk, scale, cout, cin, 3) if need_w else None
        return (from_nhwc(dxs), g_w, g_d, None, g_noise_w, g_bias, ds_next, dwb, None, None, None, None, None, None, None,
                g_map, None)


def chain_supported(gen, x, noise=None):
    from . import layers as L
    if L.get_conv_backend() != "tcgen05" or (torch.is_grad_enabled() and L.double_backward_requested()):
        return False
    if _noise_needs_grad(noise):
        return False
    blocks = [gen.conv1] + list(gen.convs)
    return all(type(m).__name__ in ("StyledConv", "StyledMapConv") and supported(m.conv, x) for m in blocks)


def _rgb_weights(to_rgb, s):
    """Per-sample modulated 1x1 weights of a ToRGB [B,3,C] (reference layers.py:296 with demodulate=False); s = its style."""
    conv = to_rgb.conv
    return (conv.weight[0, :, :, 0, 0] * conv.scale).unsqueeze(0) * s.unsqueeze(1)


def generator_chain_forward(gen, latent, noise, maps_fn=None):
    """Generator.forward body on the chained tensor-core blocks (same math as reference model.py:169-182).  With
    `maps_fn(k, h, w)` -> style map [B,2,h,w] (or None) for block k the blocks are StyledMapConv (GeneratorWithMap,
    reference model.py:259-285)."""
    blocks = [gen.conv1] + list(gen.convs)
    # ToRGB after conv1 (latent 1) and after the second conv of every resolution (block k even, latent k + 1)
    rgbs = {0: (gen.to_rgb1, 1)}
    rgbs.update({k: (gen.to_rgbs[k // 2 - 1], k + 1) for k in range(2, len(blocks), 2)})
    # all modulation / demodulation vectors of the network in one batched call (style.py): conv k uses latent[:, k]
    mods = [blk.conv for blk in blocks] + [rgbs[k][0].conv for k in sorted(rgbs)]
    lat_idx = list(range(len(blocks))) + [rgbs[k][1] for k in sorted(rgbs)]
    # one pass per conv weight: demodulation statistic + forward / dgrad GEMM operands (style.WeightPrepAll)
    prep = style.weight_prep_all_layers([(blk.conv.weight, blk.conv.scale, not blk.conv.upsample) for blk in blocks])
    sd = style.style_scales_all(latent, mods, lat_idx, [pr[0] for pr in prep] + [None] * len(rgbs))
    scales = sd[:len(blocks)]
    rgb_style = {k: sd[len(blocks) + j][0] for j, k in enumerate(sorted(rgbs))}
    x0 = gen.input(latent)
    xs, xs_op = ModulateTC.apply(x0, scales[0][0])
    skip = None
    for k, blk in enumerate(blocks):
        s_next = scales[k + 1][0] if k + 1 < len(blocks) else None
        to_rgb = rgbs[k][0] if k in rgbs else None
        wb = _rgb_weights(to_rgb, rgb_style[k]) if to_rgb is not None else None
        b, _, h, w = xs.shape
        oh, ow = (2 * h, 2 * w) if blk.conv.upsample else (h, w)
        smap = maps_fn(k, oh, ow) if maps_fn is not None else None
        nz = noise[k]
        if nz is None:
            nz = xs.new_empty(b, 1, oh, ow).normal_()
        taps = blk.conv.blur.kernel if blk.conv.upsample else blk.noise.weight
        xs, rgb, xs_op = StyledLayerTC.apply(xs, blk.conv.weight, scales[k][1], nz, blk.noise.weight, blk.activate.bias, s_next,
                                             wb, blk.conv.scale, blk.conv.upsample, taps, blk.activate.negative_slope,
                                             blk.activate.scale, prep[k][1], prep[k][2], smap, xs_op)
        if to_rgb is not None:
            out = rgb.permute(0, 3, 1, 2) + to_rgb.bias
            skip = out if skip is None else out + to_rgb.upsample(skip)
    return skip


# ------------------------------------------------------------------------------------------------- style-map networks
class StyleMapResBlockFn(Function):
    """ResBlock(3 -> 2 / 4, downsample=False) of GeneratorWithMap's style-map networks (reference model.py:194-216,
    layers.py:379-391) as ONE kernel per direction over [B,3,H,W] planes (csrc/stylemap_net.cu): forward writes only the
    block output, backward recomputes the intermediate activations from x and reduces every parameter gradient.
    The input gets no gradient (callers that need one take the composed path, see stylemap_resblock_supported)."""

    @staticmethod
    def forward(ctx, x, w1, b1c, b1a, w2, b2c, b2a, ws, alpha, gain):
        b, ci, h, w = x.shape
        co = w2.shape[0]
        out = torch.empty(b, co, h, w, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().sr_stylemap_resblock_forward_f32(
                _lib.ptr(out), _lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1c), _lib.ptr(b1a), _lib.ptr(w2), _lib.ptr(b2c),
                _lib.ptr(b2a), _lib.ptr(ws), b, ci, co, h, w, float(alpha), float(gain), _lib.stream_of(x))
        _lib.check(rc, "sr_stylemap_resblock_forward_f32")
        ctx.save_for_backward(x, w1, b1c, b1a, w2, b2c, b2a, ws)
        ctx.alpha, ctx.gain = alpha, gain
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w1, b1c, b1a, w2, b2c, b2a, ws = ctx.saved_tensors
        b, ci, h, w = x.shape
        co = w2.shape[0]
        n1, n2 = ci * ci * 9, co * ci * 9
        grads = torch.empty(n1 + ci + n2 + co + co * ci, dtype=torch.float32, device=x.device)
        g = g.contiguous()
        with torch.cuda.device(x.device):
            rc = _lib.lib().sr_stylemap_resblock_backward_f32(
                _lib.ptr(grads), _lib.ptr(g), _lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1c), _lib.ptr(b1a), _lib.ptr(w2),
                _lib.ptr(b2c), _lib.ptr(b2a), _lib.ptr(ws), b, ci, co, h, w, float(ctx.alpha), float(ctx.gain),
                _lib.stream_of(x))
        _lib.check(rc, "sr_stylemap_resblock_backward_f32")
        dw1, db1 = grads[:n1].view_as(w1), grads[n1:n1 + ci]
        dw2, db2 = grads[n1 + ci:n1 + ci + n2].view_as(w2), grads[n1 + ci + n2:n1 + ci + n2 + co]
        dws = grads[n1 + ci + n2 + co:].view_as(ws)
        return (None, dw1, db1 if b1c is not None else None, db1 if b1a is not None else None, dw2,
                db2 if b2c is not None else None, db2 if b2a is not None else None, dws, None, None)


def stylemap_resblock_supported(block, x):
    """The fused kernel covers the blocks GeneratorWithMap builds by default: 3 -> 2 / 4 channels, stride 1, fp32 NCHW CUDA
    input without gradient, first-order differentiation only."""
    from .layers import EqualConv2d, double_backward_requested
    from .op.fused_act import FusedLeakyReLU
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
        return False
    if (x.requires_grad and torch.is_grad_enabled()) or (torch.is_grad_enabled() and double_backward_requested()):
        return False
    c1, c2, sk = list(block.conv1), list(block.conv2), list(block.skip)
    if not (len(c1) == 2 and len(c2) == 2 and len(sk) == 1):
        return False
    if not (isinstance(c1[0], EqualConv2d) and isinstance(c2[0], EqualConv2d) and isinstance(sk[0], EqualConv2d)
            and isinstance(c1[1], FusedLeakyReLU) and isinstance(c2[1], FusedLeakyReLU)):
        return False
    if c1[1].negative_slope != c2[1].negative_slope or c1[1].scale != c2[1].scale:
        return False
    ok = lambda c, k: c.weight.shape[2:] == (k, k) and c.stride == 1 and c.padding == k // 2    # noqa: E731
    return (ok(c1[0], 3) and ok(c2[0], 3) and ok(sk[0], 1) and sk[0].bias is None and c1[0].weight.shape[:2] == (3, 3)
            and c2[0].weight.shape[0] in (2, 4) and c2[0].weight.shape[1] == 3 and sk[0].weight.shape[:2] == c2[0].weight.shape[:2])


def stylemap_resblock(block, x):
    c1, a1 = block.conv1[0], block.conv1[1]
    c2, a2 = block.conv2[0], block.conv2[1]
    return StyleMapResBlockFn.apply(x.contiguous(), c1.weight, c1.bias, a1.bias, c2.weight, c2.bias, a2.bias,
                                    block.skip[0].weight, a1.negative_slope, a1.scale)


# ------------------------------------------------------------------------------------------------- small-channel convs
def _small_conv_call(x, w):
    b, ci, h, wd = x.shape
    co, k = w.shape[0], w.shape[2]
    y = torch.empty(b, co, h, wd, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().sr_small_conv_f32(_lib.ptr(y), _lib.ptr(x), _lib.ptr(w), b, ci, co, k, h, wd, _lib.stream_of(x))
    _lib.check(rc, "sr_small_conv_f32")
    return y


def _flip_t(w):
    """Weights of the data gradient of a stride-1 'same' convolution: spatially flipped, channel axes swapped."""
    return w.flip(2, 3).transpose(0, 1)


class SmallConvFn(Function):
    """y = conv(x, w): 1..8 channels, 1x1 or 3x3, stride 1, zero padding k // 2 (csrc/stylemap_net.cu).  With SmallWgradFn
    it is differentiable to any order on the same two kernels: the style-map nets of GeneratorWithMap inside the
    regulariser iterations (reference train.py:335-354), where torch's double backward of a cuDNN convolution computes
    weight gradients as convolutions with image-sized kernels (12.5 ms per call at [8,3,256,256])."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        return _small_conv_call(x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = SmallConvFn.apply(gy, _flip_t(w)) if ctx.needs_input_grad[0] else None
        gw = SmallWgradFn.apply(gy, x, w.shape[2]) if ctx.needs_input_grad[1] else None
        return gx, gw


class SmallWgradFn(Function):
    """dw[o,i,ky,kx] = sum gy[n,o,p] x[n,i,p + (ky,kx) - k/2]: bilinear in (gy, x); its derivatives are convolutions."""

    @staticmethod
    def forward(ctx, gy, x, k):
        gy, x = gy.contiguous(), x.contiguous()
        ctx.save_for_backward(gy, x)
        b, ci, h, wd = x.shape
        co = gy.shape[1]
        dw = torch.empty(co, ci, k, k, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().sr_small_conv_wgrad_f32(_lib.ptr(dw), _lib.ptr(gy), _lib.ptr(x), b, ci, co, k, h, wd,
                                                    _lib.stream_of(x))
        _lib.check(rc, "sr_small_conv_wgrad_f32")
        return dw

    @staticmethod
    def backward(ctx, gdw):
        gy, x = ctx.saved_tensors
        ggy = SmallConvFn.apply(x, gdw) if ctx.needs_input_grad[0] else None
        gx = SmallConvFn.apply(gy, _flip_t(gdw)) if ctx.needs_input_grad[1] else None
        return ggy, gx, None


def small_conv_supported(conv, x):
    w = conv.weight
    k = w.shape[2]
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and w.shape[2] == w.shape[3] and k in (1, 3)
            and conv.stride == 1 and conv.padding == k // 2 and w.shape[0] <= 8 and w.shape[1] <= 8
            and w.shape[0] * w.shape[1] * k * k <= 256)


def small_conv(conv, x):
    out = SmallConvFn.apply(x, conv.weight * conv.scale)
    return out if conv.bias is None else out + conv.bias.view(1, -1, 1, 1)


# ------------------------------------------------------------------------------------------------- Discriminator stem
class StemConvFn(Function):
    """ConvLayer(3, C, 1) of the Discriminator (reference model.py:303): 1x1 EqualConv2d + bias + FusedLeakyReLU in one
    bandwidth pass per direction (csrc/stem_conv.cu).  Output is channels_last (logical [N,C,H,W]) for the tensor-core
    ResBlocks that follow.  First-order gradients (x, weight, both biases); the backward recomputes the activation mask
    from x, so nothing is saved but x itself."""

    @staticmethod
    def forward(ctx, x, w, b_conv, b_act, alpha, gain):
        n, ci, h, wd = x.shape
        co = w.shape[0]
        nhwc = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        if not nhwc:
            x = x.contiguous()
        y = torch.empty(n, h, wd, co, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().sr_stem_conv_forward_f32(_lib.ptr(y), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b_conv), _lib.ptr(b_act),
                                                     n, ci, co, h, wd, int(nhwc), float(alpha), float(gain), _lib.stream_of(x))
        _lib.check(rc, "sr_stem_conv_forward_f32")
        ctx.save_for_backward(x, w, b_conv, b_act)
        ctx.cfg = (nhwc, alpha, gain)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, b_conv, b_act = ctx.saved_tensors
        nhwc, alpha, gain = ctx.cfg
        n, ci, h, wd = x.shape
        co = w.shape[0]
        g = gy.permute(0, 2, 3, 1).contiguous()                              # a free view for channels_last gradients
        grads = torch.empty(co * 4, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None      # same layout as x
        with torch.cuda.device(x.device):
            rc = _lib.lib().sr_stem_conv_backward_f32(_lib.ptr(grads), _lib.ptr(dx), _lib.ptr(g), _lib.ptr(x), _lib.ptr(w),
                                                      _lib.ptr(b_conv), _lib.ptr(b_act), n, ci, co, h, wd, int(nhwc),
                                                      float(alpha), float(gain), _lib.stream_of(x))
        _lib.check(rc, "sr_stem_conv_backward_f32")
        dw = grads[:co * ci].view_as(w) if ctx.needs_input_grad[1] else None
        db = grads[co * ci:]
        return (dx, dw, db if (b_conv is not None and ctx.needs_input_grad[2]) else None,
                db if (b_act is not None and ctx.needs_input_grad[3]) else None, None, None)


def stem_conv_supported(conv, act, x):
    from .layers import double_backward_requested
    w = conv.weight
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and act is not None and w.shape[1] == 3 and x.shape[1] == 3
            and w.shape[2] == 1 and w.shape[3] == 1 and conv.stride == 1 and conv.padding == 0 and w.shape[0] % 4 == 0
            and 4 <= w.shape[0] <= 1024 and not (torch.is_grad_enabled() and double_backward_requested()))


def stem_conv(conv, act, x):
    return StemConvFn.apply(x, conv.weight, conv.bias, act.bias, act.negative_slope, act.scale)
