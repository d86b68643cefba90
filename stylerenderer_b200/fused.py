"""StyledConv as one differentiable block on the tcgen05 kernels (forward: modulate -> implicit-GEMM conv with the
demodulate/noise/bias/leaky-ReLU epilogue; backward: activation backward -> dgrad GEMM -> wgrad GEMM).

Replaces, for one StyledConv (reference model.py:26-32 + layers.py:293-323 + op/fused_act.py), the chain
  weight*style -> demod -> grouped conv [-> blur] -> + noise -> + bias, lrelu, *sqrt2
and its autograd graph.  Internal layout is NHWC; logical shapes stay NCHW (channels_last strides) so callers and
the state_dict are untouched.  First-order gradients only (the R1 / path-length regularisers need double backward
and run on the composed-op path, see layers.get_conv_backend()).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import tc_conv as tc
from .op.upfirdn2d import upfirdn2d_raw


def to_nhwc(x):
    """logical [N,C,H,W] -> contiguous [N,H,W,C] (a free view when x is channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()


def from_nhwc(x):
    return x.permute(0, 3, 1, 2)


def supported(mod, x):
    return (x.is_cuda and x.dtype == torch.float32 and mod.kernel_size == 3 and not mod.downsample and mod.demodulate
            and tc.supported(mod.in_channel, mod.out_channel) and tc.wgrad_supported(mod.in_channel, mod.out_channel))


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
        return from_nhwc(y)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x_nhwc, xs, y, t, weight, s, d, noise, noise_weight, act_bias, blur_taps = ctx.saved_tensors
        scale, upsample, alpha, gain = ctx.cfg
        b, h, w, cin = x_nhwc.shape
        cout = y.shape[3]
        oh, ow = y.shape[1], y.shape[2]
        gy = to_nhwc(gy)
        if not upsample:
            # one pass: activation backward, bias / noise-weight gradients, e = sum g_pre * (d * acc), ga = tf32(g_pre * d)
            ga, g_bias, g_noise_w, e = tc.bwd_prologue(gy, y, noise, noise_weight, act_bias, d, alpha, gain, True)
            dxs = tc.conv3x3(ga, tc.weight_prep(weight[0], scale, 1))
            dwk = tc.wgrad3x3(ga, xs)
        else:
            g_pre, g_bias, g_noise_w, _ = tc.bwd_prologue(gy, y, noise, noise_weight, act_bias, None, alpha, gain, False)
            gt = upfirdn2d_raw(g_pre, torch.flip(blur_taps, [0, 1]), 1, 1, 1, 1, 2, 2, 2, 2)   # transpose of the FIR
            ga, e = tc.scale_dot(gt, t, d, True)                       # ga = tf32(gt * d), e = sum gt * t  (t = d * acc)
            dxs = tc.conv3x3_s2_gather(ga, tc.weight_prep(weight[0], scale, 2), (h, w))
            dwk = tc.wgrad_transpose3x3_s2(ga, xs)
        g_d = e / d                                                     # dL/dd = sum g * acc = e / d
        g_x, g_s = tc.scale_dot(dxs, x_nhwc, s, False)                 # dx = dxs * s, ds = sum_p dxs * x
        g_x = from_nhwc(g_x)
        g_w = (dwk.view(cout, 3, 3, cin).permute(0, 3, 1, 2) * scale).unsqueeze(0)
        return g_x, g_w, g_s, g_d, None, g_noise_w, g_bias, None, None, None, None, None


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
