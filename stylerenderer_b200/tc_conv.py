"""Functional front-end of the tcgen05 implicit-GEMM convolution (csrc/modconv.cu) on NHWC float32 tensors.

Internal to the package: layers.ModulatedConv2d / model.StyledConv route here when conv_backend == "tcgen05"
and the layer shape is supported (cin % 32 == 0, cout % 128 == 0).  All tensors are plain contiguous
[batch, h, w, channels] views; callers obtain them from channels_last NCHW tensors with .permute(0, 2, 3, 1).
"""
import ctypes

import torch

from . import _lib

_I32, _I64, _F, _P = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class ConvArgs(ctypes.Structure):           # mirrors `sr_conv_args` in include/stylerenderer_b200.h
    _fields_ = [("in_", _P), ("batch", _I64), ("in_h", _I64), ("in_w", _I64), ("cin", _I64),
                ("weight", _P), ("cout", _I64), ("taps_total", _I64),
                ("num_taps", _I32), ("tap_dy", _I32 * 9), ("tap_dx", _I32 * 9), ("tap_w", _I32 * 9),
                ("in_stride", _I32),
                ("grid_h", _I64), ("grid_w", _I64), ("out_h", _I64), ("out_w", _I64),
                ("out_stride", _I32), ("out_y0", _I32), ("out_x0", _I32),
                ("out", _P), ("out2", _P), ("epilogue", _I32),
                ("rowscale", _P), ("scale2", _P), ("bias", _P), ("noise", _P), ("noise_weight", _P), ("stylemap", _P),
                ("noise_batch_stride", _I64), ("stylemap_batch_stride", _I64),
                ("alpha", _F), ("gain", _F), ("rgb_weight", _P), ("rgb_out", _P)]


class WgradArgs(ctypes.Structure):         # mirrors `sr_wgrad_args`
    _fields_ = [("g", _P), ("batch", _I64), ("g_h", _I64), ("g_w", _I64), ("cout", _I64),
                ("x", _P), ("x_h", _I64), ("x_w", _I64), ("cin", _I64),
                ("grid_h", _I64), ("grid_w", _I64), ("num_taps", _I32),
                ("g_dy", _I32 * 9), ("g_dx", _I32 * 9), ("x_dy", _I32 * 9), ("x_dx", _I32 * 9), ("tap_out", _I32 * 9),
                ("g_stride", _I32), ("x_stride", _I32), ("dw", _P), ("taps_total", _I64), ("zero_init", _I32)]


def wgrad_supported(cin, cout):
    return cin % 128 == 0 and cout % 128 == 0 and cin >= 128 and cout >= 128


def supported(cin, cout):
    return cin % 32 == 0 and cin >= 32 and cout % 128 == 0 and cout >= 128


# ------------------------------------------------------------------------------------------------- operand precision
# "tf32"   (default): GEMM operands are rounded to tf32 by the kernels that produce them (cvt.rna), one MMA per product:
#          the arithmetic class of the reference's cuDNN path under torch's default cudnn.allow_tf32 = True.
# "tf32x3" (parity mode): fp32-faithful split operands.  Every operand x is split into hi = tf32(x) and lo = x - hi and the
#          SAME tcgen05 pipelines accumulate  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  into one TMEM accumulator: the conv
#          kernels get the three terms as a three times longer K loop (operand channels [hi | lo | hi] against weight
#          columns [hi | hi | lo]), the weight-gradient kernel as three accumulating launches.  Products then carry
#          ~2^-21 relative error instead of 2^-11, so no leaky-ReLU mask flips and every gradient can be checked at the
#          1e-3 bar with generic inputs.  3x the tensor-core work plus unfused glue passes: a checking mode, not a fast one.
# "bf16"   (train-step mode, BASELINE.json configs[3]): the GEMM operands -- modulated activations, re-laid-out weights,
#          the gradient operand of dgrad / wgrad -- are bfloat16 tensors (tcgen05 kind::f16: twice the tensor-core rate,
#          half the operand bytes); accumulation, saved activations, epilogue vectors, reductions and every parameter
#          gradient stay fp32.  8-bit significands: outputs agree with the fp32 oracle to ~1e-2, not 1e-3.
_PRECISION = {"mode": "tf32"}


def set_precision(mode):
    assert mode in ("tf32", "tf32x3", "bf16")
    _PRECISION["mode"] = mode


def get_precision():
    return _PRECISION["mode"]


class precision:
    """with tc_conv.precision("tf32x3"): ... (restores the previous mode)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)


def _exact():
    return _PRECISION["mode"] == "tf32x3"


def _bf16():
    return _PRECISION["mode"] == "bf16"


def operand_dtype():
    """dtype of the GEMM operand tensors in the current mode."""
    return torch.bfloat16 if _bf16() else torch.float32


def operand_like(t):
    """An empty GEMM-operand tensor shaped like the fp32 NHWC tensor t."""
    return torch.empty(t.shape, dtype=operand_dtype(), device=t.device)


def split_tf32(x):
    """x -> (hi, lo): hi = x rounded to tf32 (nearest, ties away: the bit trick equals cvt.rna.tf32.f32), lo = x - hi
    (exact in fp32; the tensor core then truncates lo to its own 10 mantissa bits, a 2^-21 relative residual)."""
    xi = x.contiguous().view(torch.int32)
    hi = ((xi + 0x1000) & -0x2000).view(torch.float32)
    return hi, x - hi


def _lrelu(t, alpha, gain):
    return torch.where(t > 0, t, t * alpha) * gain


def _styled_tail(pre, noise, noise_weight, bias, stylemap, alpha, gain):
    """y = lrelu(pre (* map0 + map1) + noise_weight * noise + bias) * gain on NHWC tensors (tf32x3 glue only)."""
    b, h, w, _ = pre.shape
    t = pre
    if stylemap is not None:
        t = t * stylemap[:, 0].reshape(b, h, w, 1) + stylemap[:, 1].reshape(b, h, w, 1)
    if noise is not None:
        t = t + noise_weight * noise.reshape(-1, h, w, 1)
    if bias is not None:
        t = t + bias
    return _lrelu(t, alpha, gain)


def _check_nhwc(t, name, operand=False):
    ok_dtype = t.dtype == torch.float32 or (operand and t.dtype == torch.bfloat16)
    if not (t.is_cuda and ok_dtype and t.dim() == 4 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous float32{' / bfloat16' if operand else ''} CUDA tensor "
                           "[batch, h, w, channels]")


def weight_prep(w, scale, mode):
    """Reference-layout weight [cout, cin, kh, kw] -> GEMM B operand [rows, kh*kw, cols], scaled, tf32-rounded.
    mode 0: rows=cout, cols=cin (forward);  1: rows=cin, cols=cout, taps flipped (dgrad of the plain conv);
    mode 2: rows=cin, cols=cout, taps as is (dgrad of the stride-2 transposed conv)."""
    cout, cin, kh, kw = w.shape
    if _exact() or _bf16():                                  # same layouts: unrounded (split by the GEMM wrappers) / bfloat16
        ws = w.detach() * scale
        if mode in (0, 3):
            ws = ws.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
        else:
            ws = (ws.flip(2, 3) if mode == 1 else ws).permute(1, 2, 3, 0).reshape(cin, kh * kw, cout)
        return ws.to(operand_dtype()).contiguous()
    rows, cols = (cout, cin) if mode in (0, 3) else (cin, cout)
    dst = torch.empty(rows, kh * kw, cols, dtype=torch.float32, device=w.device)
    wc = w.contiguous()
    with torch.cuda.device(w.device):
        rc = _lib.lib().sr_conv_weight_prep_tf32(_lib.ptr(dst), _lib.ptr(wc), float(scale), cout, cin, kh, kw, mode,
                                                 _lib.stream_of(w))
    _lib.check(rc, "sr_conv_weight_prep_tf32")
    return dst


def weight_prep_dual(w, scale, flip_transposed, want_wsq=True):
    """One pass over a [cout, cin, k, k] weight -> (fwd [cout, k*k, cin], tr [cin, k*k, cout], wsq [cout, cin] or None):
    weight_prep(mode 0), weight_prep(mode 1 if flip_transposed else 2) and scale^2 * sum_taps w^2 together."""
    cout, cin, kh, kw = w.shape
    if _exact():
        ws = w.detach() * scale
        return (weight_prep(w, scale, 0), weight_prep(w, scale, 1 if flip_transposed else 2),
                ws.pow(2).sum((2, 3)) if want_wsq else None)
    wc = w.contiguous()
    fwd = torch.empty(cout, kh * kw, cin, dtype=operand_dtype(), device=w.device)
    tr = torch.empty(cin, kh * kw, cout, dtype=operand_dtype(), device=w.device)
    wsq = torch.empty(cout, cin, dtype=torch.float32, device=w.device) if want_wsq else None
    fn = _lib.lib().sr_conv_weight_prep_dual_bf16 if _bf16() else _lib.lib().sr_conv_weight_prep_dual_tf32
    with torch.cuda.device(w.device):
        rc = fn(_lib.ptr(fwd), _lib.ptr(tr), _lib.ptr(wsq), _lib.ptr(wc), float(scale), cout, cin, kh * kw,
                1 if flip_transposed else 0, _lib.stream_of(w))
    _lib.check(rc, "sr_conv_weight_prep_dual")
    return fwd, tr, wsq


def modulate(x, style=None):
    """xs = tf32_round(x * style[b, c]) for NHWC x [B,H,W,C]; style [B,C] or None (rounding only)."""
    _check_nhwc(x, "modulate")
    b, h, w, c = x.shape
    if _exact():
        return x * style.reshape(b, 1, 1, c) if style is not None else x.clone()
    xs = operand_like(x)
    s = style.contiguous() if style is not None else None
    fn = _lib.lib().sr_modulate_bf16 if _bf16() else _lib.lib().sr_modulate_tf32
    with torch.cuda.device(x.device):
        rc = fn(_lib.ptr(xs), _lib.ptr(x), _lib.ptr(s), b, h * w, c, _lib.stream_of(x))
    _lib.check(rc, "sr_modulate")
    return xs


def _fill_args(a, x, wmat, taps, out, in_stride, grid, out_stride, out_origin, epilogue, rowscale, out2, scale2, bias,
               noise, noise_weight, stylemap, alpha, gain, rgb_weight=None, rgb_out=None):
    a.rgb_weight, a.rgb_out = _lib.ptr(rgb_weight), _lib.ptr(rgb_out)
    a.in_ = _lib.ptr(x)
    a.batch, a.in_h, a.in_w, a.cin = x.shape
    a.weight = _lib.ptr(wmat)
    a.cout, a.taps_total = wmat.shape[0], wmat.shape[1]
    a.num_taps = len(taps)
    for i, (dy, dx, wi) in enumerate(taps):
        a.tap_dy[i], a.tap_dx[i], a.tap_w[i] = dy, dx, wi
    a.in_stride = in_stride
    gh, gw = grid if grid is not None else (out.shape[1], out.shape[2])
    a.grid_h, a.grid_w, a.out_h, a.out_w = gh, gw, out.shape[1], out.shape[2]
    a.out_stride, a.out_y0, a.out_x0 = out_stride, out_origin[0], out_origin[1]
    a.out = _lib.ptr(out)
    a.out2 = _lib.ptr(out2)
    a.epilogue = epilogue
    a.rowscale, a.scale2, a.bias = _lib.ptr(rowscale), _lib.ptr(scale2), _lib.ptr(bias)
    a.noise, a.noise_weight, a.stylemap = _lib.ptr(noise), _lib.ptr(noise_weight), _lib.ptr(stylemap)
    if noise is not None:
        assert noise.is_contiguous() and noise.shape[-2:] == out.shape[1:3]
        plane = out.shape[1] * out.shape[2]
        if noise.numel() not in (plane, plane * out.shape[0]):
            raise RuntimeError(f"noise batch must be 1 or {out.shape[0]} (got {noise.numel() // max(plane, 1)} planes)")
        a.noise_batch_stride = 0 if noise.numel() == plane else plane
    if stylemap is not None:
        assert stylemap.shape[1] == 2 and stylemap.stride(3) == 1 and stylemap.stride(2) == out.shape[2] \
            and stylemap.stride(1) == out.shape[1] * out.shape[2]
        a.stylemap_batch_stride = stylemap.stride(0)
    a.alpha, a.gain = alpha, gain


def conv_igemm_multi(x, wmat, phases, out, *, in_stride=1, out_stride=1, epilogue=0, rowscale=None, out2=None,
                     scale2=None, bias=None, noise=None, noise_weight=None, stylemap=None, alpha=0.2, gain=2 ** 0.5,
                     rgb_weight=None, rgb_out=None):
    """One persistent launch over up to 4 phases [(taps, grid, out_origin)] sharing all tensors (see conv_igemm)."""
    _check_nhwc(x, "conv input", operand=True)
    _check_nhwc(out, "conv output")
    assert wmat.is_contiguous() and wmat.shape[2] == x.shape[3] and out.shape[3] == wmat.shape[0] and out.shape[0] == x.shape[0]
    op16 = x.dtype == torch.bfloat16
    if wmat.dtype != x.dtype or (out2 is not None and out2.dtype != x.dtype):
        raise RuntimeError("conv: the activation operand, the weights and the second output must share the operand dtype")
    if _exact():                     # split operands: K = [hi | lo | hi] x [hi | hi | lo] through the same kernels
        xh, xl = split_tf32(x)
        wh, wl = split_tf32(wmat)
        x = torch.cat([xh, xl, xh], 3)
        wmat = torch.cat([wh, wh, wl], 2)
    arr = (ConvArgs * len(phases))()
    for a, (taps, grid, origin) in zip(arr, phases):
        _fill_args(a, x, wmat, taps, out, in_stride, grid, out_stride, origin, epilogue, rowscale, out2, scale2, bias, noise,
                   noise_weight, stylemap, alpha, gain, rgb_weight, rgb_out)
    with torch.cuda.device(x.device):
        fn = _lib.lib().sr_conv_igemm_multi_bf16 if op16 else _lib.lib().sr_conv_igemm_multi_tf32
        rc = fn(arr, len(phases), _lib.stream_of(x))
    _lib.check(rc, "sr_conv_igemm_multi")
    if _exact() and out2 is not None:   # the next layer's operand, unrounded (the kernel wrote its tf32 rounding)
        y = out if epilogue != 2 else _styled_tail(out, noise, noise_weight, bias, stylemap, alpha, gain)
        out2.copy_(y * scale2.reshape(out.shape[0], 1, 1, -1))
    return out


def conv_igemm(x, wmat, taps, out, *, grid=None, out_origin=(0, 0), **kw):
    """out[n, y0 + gy*os, x0 + gx*os, :] = epilogue( sum_t x[n, gy*is + dy_t, gx*is + dx_t, :] @ wmat[:, w_t, :]^T ).
    taps: list of (dy, dx, weight_tap_index)."""
    return conv_igemm_multi(x, wmat, [(taps, grid, out_origin)], out, **kw)


TAPS_3X3 = [(ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]


def conv3x3(x, wmat, out=None, **kw):
    """3x3, stride 1, zero padding 1 (cross-correlation like F.conv2d); wmat from weight_prep(mode 0 or 1)."""
    if out is None:
        out = torch.empty(x.shape[0], x.shape[1], x.shape[2], wmat.shape[0], dtype=torch.float32, device=x.device)
    return conv_igemm(x, wmat, TAPS_3X3, out, **kw)


def conv_transpose3x3_s2(x, wmat, out=None, rowscale=None):
    """Stride-2 transposed 3x3 conv, no padding: [B,H,W,Cin] -> [B,2H+1,2W+1,Cout] as four phase GEMMs
    (reference layers.py:301-309: F.conv_transpose2d(stride=2, padding=0)); wmat from weight_prep(mode 0)."""
    b, h, w, _ = x.shape
    if out is None:
        out = torch.empty(b, 2 * h + 1, 2 * w + 1, wmat.shape[0], dtype=torch.float32, device=x.device)
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            taps = [(-(ky - py) // 2, -(kx - px) // 2, ky * 3 + kx)
                    for ky in range(py, 3, 2) for kx in range(px, 3, 2)]
            phases.append((taps, (h + 1 - py, w + 1 - px), (py, px)))
    return conv_igemm_multi(x, wmat, phases, out, out_stride=2, rowscale=rowscale)


def conv3x3_s2_gather(g, wmat, out_hw, out=None, rowscale=None):
    """dgrad of the transposed conv: dx[m,n] = sum_{ky,kx} g[2m+ky, 2n+kx] @ W  (stride-2 gather through TMA
    element strides); g [B,2H+1,2W+1,Cout], wmat from weight_prep(mode 2) -> [B,H,W,Cin]."""
    h, w = out_hw
    if out is None:
        out = torch.empty(g.shape[0], h, w, wmat.shape[0], dtype=torch.float32, device=g.device)
    taps = [(ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]
    return conv_igemm(g, wmat, taps, out, in_stride=2, rowscale=rowscale)


def wgrad(g, x, taps, grid, *, g_stride=1, x_stride=1, taps_total=9, dw=None):
    """dw[co, t_out, ci] = sum_{n,gy,gx} g[n, gy*gs + gdy, gx*gs + gdx, co] * x[n, gy*xs + xdy, gx*xs + xdx, ci];
    taps: list of (g_dy, g_dx, x_dy, x_dx, t_out) -> [cout, taps_total, cin]."""
    _check_nhwc(g, "wgrad g", operand=True)
    _check_nhwc(x, "wgrad x", operand=True)
    if g.dtype != x.dtype:
        raise RuntimeError("wgrad: both operands must share the operand dtype")
    wg_fn = _lib.lib().sr_conv_wgrad_bf16 if g.dtype == torch.bfloat16 else _lib.lib().sr_conv_wgrad_tf32
    a = WgradArgs()
    a.g, a.x = _lib.ptr(g), _lib.ptr(x)
    a.batch, a.g_h, a.g_w, a.cout = g.shape
    _, a.x_h, a.x_w, a.cin = x.shape
    a.grid_h, a.grid_w = grid
    a.num_taps = len(taps)
    for i, (gdy, gdx, xdy, xdx, to) in enumerate(taps):
        a.g_dy[i], a.g_dx[i], a.x_dy[i], a.x_dx[i], a.tap_out[i] = gdy, gdx, xdy, xdx, to
    a.g_stride, a.x_stride = g_stride, x_stride
    if dw is None:
        dw = torch.empty(g.shape[3], taps_total, x.shape[3], dtype=torch.float32, device=g.device)
    a.dw, a.taps_total, a.zero_init = _lib.ptr(dw), taps_total, 1
    if _exact():                     # G_hi X_hi + G_lo X_hi + G_hi X_lo accumulated by three launches of the same kernel
        gh, gl = split_tf32(g)
        xh, xl = split_tf32(x)
        terms = [(gh, xh), (gl, xh), (gh, xl)]
    else:
        terms = [(g, x)]
    with torch.cuda.device(g.device):
        for i, (gt, xt) in enumerate(terms):
            a.g, a.x, a.zero_init = _lib.ptr(gt), _lib.ptr(xt), 1 if i == 0 else 0
            rc = wg_fn(ctypes.byref(a), _lib.stream_of(g))
            _lib.check(rc, "sr_conv_wgrad")
    return dw


def wgrad3x3(g, x):
    """Weight gradient of conv3x3 (stride 1, pad 1): dw[co, ky*3+kx, ci] = sum g[n,y,x,co] * x[n,y+ky-1,x+kx-1,ci]."""
    taps = [(0, 0, ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]
    return wgrad(g, x, taps, (g.shape[1], g.shape[2]))


def wgrad_transpose3x3_s2(g, x):
    """Weight gradient of the stride-2 transposed conv: dw[co, ky*3+kx, ci] = sum g[n,2m+ky,2n+kx,co] * x[n,m,n,ci]."""
    taps = [(ky, kx, 0, 0, ky * 3 + kx) for ky in range(3) for kx in range(3)]
    return wgrad(g, x, taps, (x.shape[1], x.shape[2]), g_stride=2)


def _noise_args(noise, oh, ow, batch=None):
    if noise is None:
        return None, 0
    nz = noise.reshape(-1, oh, ow).contiguous()
    if batch is not None and nz.shape[0] not in (1, batch):
        raise RuntimeError(f"noise batch must be 1 or {batch} (got {nz.shape[0]})")
    return nz, (0 if nz.shape[0] == 1 else oh * ow)


def _map_args(stylemap, oh, ow):
    """[B,2,oh,ow] style map (possibly a channel slice of a wider tensor) -> (tensor, batch stride in floats)."""
    if stylemap is None:
        return None, 0
    assert stylemap.shape[1] == 2 and stylemap.shape[2:] == (oh, ow) and stylemap.stride(3) == 1 and \
        stylemap.stride(2) == ow and stylemap.stride(1) == oh * ow, "stylemap must be [B,2,H,W] with contiguous planes"
    return stylemap, stylemap.stride(0)


def blur_styled(t, taps, pad, noise, noise_weight, bias, alpha, gain, scale2=None, stylemap=None):
    """NHWC 4x4 FIR (* map0 + map1) + noise + bias + leaky-ReLU*gain in one pass: [B,IH,IW,C] -> [B,OH,OW,C];
    with scale2 [B,C] also returns out2 = tf32(out * scale2) (the next layer's GEMM operand)."""
    _check_nhwc(t, "blur_styled")
    b, ih, iw, c = t.shape
    oh, ow = ih + pad[0] + pad[1] - 3, iw + pad[0] + pad[1] - 3
    out = torch.empty(b, oh, ow, c, dtype=torch.float32, device=t.device)
    out2 = operand_like(out) if scale2 is not None else None
    nz, nbs = _noise_args(noise, oh, ow, b)
    sm, sms = _map_args(stylemap, oh, ow)
    fn = _lib.lib().sr_blur_nhwc_styled3_bf16 if _bf16() else _lib.lib().sr_blur_nhwc_styled3_f32
    with torch.cuda.device(t.device):
        rc = fn(_lib.ptr(out), _lib.ptr(out2), _lib.ptr(scale2), _lib.ptr(t), _lib.ptr(taps.contiguous()), b, ih, iw, c,
                pad[0], pad[1], _lib.ptr(nz), nbs, _lib.ptr(noise_weight), _lib.ptr(bias), float(alpha), float(gain),
                _lib.ptr(sm), sms, _lib.stream_of(t))
    _lib.check(rc, "sr_blur_nhwc_styled3_f32")
    if _exact() and scale2 is not None:
        y = out if stylemap is None else _styled_tail(out, noise, noise_weight, bias, stylemap, alpha, gain)
        out2.copy_(y * scale2.reshape(b, 1, 1, c))
    return out if scale2 is None else (out, out2)


def bwd_prologue(gy, y, noise, noise_weight, bias, d, alpha, gain, want_e):
    """-> (ga, g_bias [C], g_noise_w [1], e [B,C] or None); see sr_styled_bwd_prologue_f32."""
    _check_nhwc(gy, "bwd_prologue gy")
    _check_nhwc(y, "bwd_prologue y")
    b, h, w, c = y.shape
    dev = y.device
    d_k = None if _exact() else d                       # tf32x3: the kernel leaves g_pre unrounded, * d below
    ga = operand_like(y) if d_k is not None else torch.empty_like(y)      # the GEMM operand only when d is applied here
    g_bias = torch.empty(c, dtype=torch.float32, device=dev)
    g_nw = torch.empty(1, dtype=torch.float32, device=dev)
    e = torch.empty(b, c, dtype=torch.float32, device=dev) if want_e else None
    nz, nbs = _noise_args(noise, h, w, b)
    fn = _lib.lib().sr_styled_bwd_prologue3_bf16 if _bf16() else _lib.lib().sr_styled_bwd_prologue3_f32
    with torch.cuda.device(dev):
        rc = fn(_lib.ptr(ga), _lib.ptr(g_bias), _lib.ptr(g_nw), _lib.ptr(e), None, None, _lib.ptr(gy), None, None, None, None,
                _lib.ptr(y), _lib.ptr(nz), nbs, _lib.ptr(noise_weight), _lib.ptr(bias), _lib.ptr(d_k), b, h * w, c,
                float(alpha), float(gain), None, 0, None, _lib.stream_of(y))
    _lib.check(rc, "sr_styled_bwd_prologue3")
    if d is not None and d_k is None:
        ga = ga * d.reshape(b, 1, 1, c)
    return ga, g_bias, g_nw, e


def scale_dot(a, other, scale, round_out, want_out=True):
    """-> (a * scale[b,c] (tf32-rounded if round_out) or None, dot[b,c] = sum_p a*other or None)."""
    _check_nhwc(a, "scale_dot a")
    b, h, w, c = a.shape
    out = torch.empty_like(a) if want_out else None
    dot = torch.empty(b, c, dtype=torch.float32, device=a.device) if other is not None else None
    if _exact():
        round_out = False
    with torch.cuda.device(a.device):
        rc = _lib.lib().sr_scale_dot_nhwc_f32(_lib.ptr(out), _lib.ptr(dot), _lib.ptr(a), _lib.ptr(other),
                                              _lib.ptr(scale.contiguous() if scale is not None else None), b, h * w, c,
                                              int(bool(round_out)), _lib.stream_of(a))
    _lib.check(rc, "sr_scale_dot_nhwc_f32")
    if round_out and _bf16() and out is not None:        # this (unchained) path gets its bf16 operand from a torch cast
        out = out.to(torch.bfloat16)
    return out, dot


def bwd_prologue2(y, noise, noise_weight, bias, d, alpha, gain, want_e, gy=None, gxs=None, s_next=None, g_rgb=None,
                  rgb_weight=None, stylemap=None):
    """Chained backward prologue (sr_styled_bwd_prologue3_f32) -> (ga, g_bias, g_noise_w, e, ds_next, d_rgb_weight[, g_map]);
    g_map [B,2,H,W] is appended when a stylemap is given."""
    _check_nhwc(y, "bwd_prologue2 y")
    b, h, w, c = y.shape
    dev = y.device
    d_k = None if _exact() else d                       # tf32x3: the kernel leaves g_pre unrounded, * d below
    ga = operand_like(y) if d_k is not None else torch.empty_like(y)      # the GEMM operand only when d is applied here
    g_bias = torch.empty(c, dtype=torch.float32, device=dev)
    g_nw = torch.empty(1, dtype=torch.float32, device=dev)
    e = torch.empty(b, c, dtype=torch.float32, device=dev) if want_e else None
    ds_next = torch.empty(b, c, dtype=torch.float32, device=dev) if gxs is not None else None
    dwb = torch.empty(b, 3, c, dtype=torch.float32, device=dev) if g_rgb is not None else None
    nz, nbs = _noise_args(noise, h, w, b)
    for t_ in (gy, gxs):
        if t_ is not None:
            _check_nhwc(t_, "bwd_prologue2 gradient")
    if g_rgb is not None:
        assert g_rgb.is_contiguous() and g_rgb.shape == (b, h, w, 3) and rgb_weight.is_contiguous()
    sm, sms = _map_args(stylemap, h, w)
    g_map = torch.empty(b, 2, h, w, dtype=torch.float32, device=dev) if sm is not None else None
    fn = _lib.lib().sr_styled_bwd_prologue3_bf16 if _bf16() else _lib.lib().sr_styled_bwd_prologue3_f32
    with torch.cuda.device(dev):
        rc = fn(
            _lib.ptr(ga), _lib.ptr(g_bias), _lib.ptr(g_nw), _lib.ptr(e), _lib.ptr(ds_next), _lib.ptr(dwb), _lib.ptr(gy),
            _lib.ptr(gxs), _lib.ptr(s_next), _lib.ptr(g_rgb), _lib.ptr(rgb_weight), _lib.ptr(y), _lib.ptr(nz), nbs,
            _lib.ptr(noise_weight), _lib.ptr(bias), _lib.ptr(d_k), b, h * w, c, float(alpha), float(gain), _lib.ptr(sm), sms,
            _lib.ptr(g_map), _lib.stream_of(y))
    _lib.check(rc, "sr_styled_bwd_prologue3_f32")
    if d is not None and d_k is None:
        ga = ga * d.reshape(b, 1, 1, c)
    if sm is not None:
        return ga, g_bias, g_nw, e, ds_next, dwb, g_map
    return ga, g_bias, g_nw, e, ds_next, dwb


def blur_scaledot(x, taps, pad, scale, other=None):
    """(tf32(fir(x) * scale[b,c]), dot[b,c] = sum_p fir(x) * other): the up-sampling block's backward FIR with its tail.
    other = None: scale only (dot is None) -- the block gets its demodulation gradient from the prologue instead."""
    _check_nhwc(x, "blur_scaledot")
    b, ih, iw, c = x.shape
    oh, ow = ih + pad[0] + pad[1] - 3, iw + pad[0] + pad[1] - 3
    assert other is None or (other.shape == (b, oh, ow, c) and other.is_contiguous())
    if _exact():                                         # plain FIR kernel + unrounded torch tail
        from .op.upfirdn2d import upfirdn2d_raw
        f = upfirdn2d_raw(x, taps, 1, 1, 1, 1, pad[0], pad[1], pad[0], pad[1])
        return f * scale.reshape(b, 1, 1, c), ((f * other).sum((1, 2)) if other is not None else None)
    out = torch.empty(b, oh, ow, c, dtype=operand_dtype(), device=x.device)
    dot = torch.empty(b, c, dtype=torch.float32, device=x.device) if other is not None else None
    fn = _lib.lib().sr_blur_nhwc_scaledot_bf16 if _bf16() else _lib.lib().sr_blur_nhwc_scaledot_f32
    with torch.cuda.device(x.device):
        rc = fn(_lib.ptr(out), _lib.ptr(dot), _lib.ptr(x), _lib.ptr(taps.contiguous()), _lib.ptr(scale.contiguous()),
                _lib.ptr(other), b, ih, iw, c, pad[0], pad[1], _lib.stream_of(x))
    _lib.check(rc, "sr_blur_nhwc_scaledot")
    return out, dot
