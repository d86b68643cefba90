"""Style path of every modulated convolution of a network in a handful of launches (csrc/style_ops.cu).

Reference per layer (layers.py:232-239, 295-299): `s = modulation(style)` (EqualLinear) and the demodulation
`rsqrt(sum w^2 + 1e-8)`.  Here (activation-scaling form, see layers.ModulatedConv2d.style_scales):
    s[b,i] = c * latent[b, li] . Wm[i] + bm[i],      d[b,o] = rsqrt(sum_i s[b,i]^2 * Wsq[o,i] + eps)
computed for ALL layers by `style_scales_all` (2 launches forward, 4 backward) instead of ~35 torch launches per layer.
"""
import ctypes
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_P, _I32 = ctypes.c_void_p, ctypes.c_int32


class StyleLayer(ctypes.Structure):          # mirrors `sr_style_layer` in include/stylerenderer_b200.h
    _fields_ = [("mod_weight", _P), ("mod_bias", _P), ("wsq", _P), ("s", _P), ("d", _P),
                ("cin", _I32), ("cout", _I32), ("latent_index", _I32), ("reserved", _I32),
                ("g_s", _P), ("g_d", _P), ("g_mod_weight", _P), ("g_mod_bias", _P), ("g_wsq", _P),
                ("gs_total", _P), ("du", _P)]


class WeightSq(Function):
    """wsq[o,i] = scale^2 * sum_taps W[0,o,i,:,:]^2 -- the weight statistic the demodulation needs (layers.py:297)."""

    @staticmethod
    def forward(ctx, weight, scale):
        _, cout, cin, kh, kw = weight.shape
        w = weight.contiguous()
        wsq = torch.empty(cout, cin, dtype=torch.float32, device=weight.device)
        with torch.cuda.device(weight.device):
            rc = _lib.lib().sr_weight_sq_f32(_lib.ptr(wsq), _lib.ptr(w), float(scale), cout, cin, kh * kw, _lib.stream_of(w))
        _lib.check(rc, "sr_weight_sq_f32")
        ctx.save_for_backward(w)
        ctx.scale = float(scale)
        return wsq

    @staticmethod
    @once_differentiable
    def backward(ctx, g_wsq):
        w, = ctx.saved_tensors
        _, cout, cin, kh, kw = w.shape
        gw = torch.empty_like(w)
        g = g_wsq.contiguous()
        with torch.cuda.device(w.device):
            rc = _lib.lib().sr_weight_sq_backward_f32(_lib.ptr(gw), _lib.ptr(w), _lib.ptr(g), ctx.scale, cout, cin, kh * kw,
                                                      _lib.stream_of(w))
        _lib.check(rc, "sr_weight_sq_backward_f32")
        return gw, None


class WeightPrepAll(Function):
    """(wsq, wk_fwd, wk_tr) of a 3x3 modulated-conv weight in ONE pass over it (sr_conv_weight_prep_dual_tf32): the
    demodulation statistic (differentiable, see WeightSq) and the two tf32 GEMM operand layouts the tensor-core block
    uses in its forward and backward (constants of the step, not differentiable)."""

    @staticmethod
    def forward(ctx, weight, scale, flip_transposed):
        from . import tc_conv as tc
        w = weight.contiguous()
        wk_f, wk_t, wsq = tc.weight_prep_dual(w[0], scale, flip_transposed)
        ctx.save_for_backward(w)
        ctx.scale = float(scale)
        ctx.mark_non_differentiable(wk_f, wk_t)
        return wsq, wk_f, wk_t

    @staticmethod
    @once_differentiable
    def backward(ctx, g_wsq, _gf, _gt):
        if g_wsq is None:
            return None, None, None
        w, = ctx.saved_tensors
        _, cout, cin, kh, kw = w.shape
        gw = torch.empty_like(w)
        g = g_wsq.contiguous()
        with torch.cuda.device(w.device):
            rc = _lib.lib().sr_weight_sq_backward_f32(_lib.ptr(gw), _lib.ptr(w), _lib.ptr(g), ctx.scale, cout, cin, kh * kw,
                                                      _lib.stream_of(w))
        _lib.check(rc, "sr_weight_sq_backward_f32")
        return gw, None, None


class WeightPrepItem(ctypes.Structure):      # mirrors `sr_weight_prep_item` in include/stylerenderer_b200.h
    _fields_ = [("fwd", _P), ("tr", _P), ("wsq", _P), ("w", _P), ("gw", _P), ("g_wsq", _P), ("scale", ctypes.c_float),
                ("cout", _I32), ("cin", _I32), ("taps", _I32), ("flip_transposed", _I32), ("reserved", _I32)]


WEIGHT_PREP_MAX = 16


class WeightPrepAllLayers(Function):
    """WeightPrepAll for every conv weight of a network in ONE launch (sr_conv_weight_prep_multi_*), and one launch for all
    the demodulation-statistic gradients in the backward (sr_weight_sq_backward_multi_f32).
    forward(cfgs, *weights) with cfgs = ((scale, flip_transposed), ...) -> (wsq_0, wk_f_0, wk_t_0, wsq_1, ...)."""

    @staticmethod
    def forward(ctx, cfgs, *weights):
        from . import tc_conv as tc
        n = len(weights)
        assert 1 <= n <= WEIGHT_PREP_MAX and len(cfgs) == n
        ws = [w.contiguous() for w in weights]
        arr = (WeightPrepItem * n)()
        outs = []
        for a, w, (scale, flip) in zip(arr, ws, cfgs):
            _, cout, cin, kh, kw = w.shape
            wk_f = torch.empty(cout, kh * kw, cin, dtype=tc.operand_dtype(), device=w.device)
            wk_t = torch.empty(cin, kh * kw, cout, dtype=tc.operand_dtype(), device=w.device)
            wsq = torch.empty(cout, cin, dtype=torch.float32, device=w.device)
            a.fwd, a.tr, a.wsq, a.w = _lib.ptr(wk_f), _lib.ptr(wk_t), _lib.ptr(wsq), _lib.ptr(w)
            a.scale, a.cout, a.cin, a.taps, a.flip_transposed = float(scale), cout, cin, kh * kw, int(bool(flip))
            outs += [wsq, wk_f, wk_t]
        fn = _lib.lib().sr_conv_weight_prep_multi_bf16 if tc._bf16() else _lib.lib().sr_conv_weight_prep_multi_tf32
        with torch.cuda.device(ws[0].device):
            rc = fn(arr, n, _lib.stream_of(ws[0]))
        _lib.check(rc, "sr_conv_weight_prep_multi")
        ctx.save_for_backward(*ws)
        ctx.cfgs = cfgs
        ctx.mark_non_differentiable(*[o for i, o in enumerate(outs) if i % 3])
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        ws = ctx.saved_tensors
        live = [(i, w, grads[3 * i]) for i, w in enumerate(ws) if grads[3 * i] is not None and ctx.needs_input_grad[1 + i]]
        out = [None] * len(ws)
        if live:
            arr = (WeightPrepItem * len(live))()
            keep = []
            for a, (i, w, g) in zip(arr, live):
                _, cout, cin, kh, kw = w.shape
                gw, g = torch.empty_like(w), g.contiguous()
                a.w, a.gw, a.g_wsq = _lib.ptr(w), _lib.ptr(gw), _lib.ptr(g)
                a.scale, a.cout, a.cin, a.taps = float(ctx.cfgs[i][0]), cout, cin, kh * kw
                out[i] = gw
                keep.append(g)
            with torch.cuda.device(ws[0].device):
                rc = _lib.lib().sr_weight_sq_backward_multi_f32(arr, len(live), _lib.stream_of(ws[0]))
            _lib.check(rc, "sr_weight_sq_backward_multi_f32")
        return (None, *out)


def weight_prep_all_layers(mods_cfg):
    """[(weight [1,cout,cin,k,k], scale, flip_transposed), ...] -> [(wsq, wk_fwd, wk_tr), ...]: one launch when the operand mode
    allows it (tf32 / bf16), per-layer WeightPrepAll in the fp32-faithful parity mode or beyond WEIGHT_PREP_MAX layers."""
    from . import tc_conv as tc
    if tc._exact() or not 1 <= len(mods_cfg) <= WEIGHT_PREP_MAX or os.environ.get("SR_WEIGHT_PREP_MULTI", "1") == "0":
        return [WeightPrepAll.apply(w, sc, fl) for w, sc, fl in mods_cfg]
    flat = WeightPrepAllLayers.apply(tuple((float(sc), bool(fl)) for _, sc, fl in mods_cfg), *[w for w, _, _ in mods_cfg])
    return [tuple(flat[3 * i:3 * i + 3]) for i in range(len(mods_cfg))]


def weight_grad_layout(dwk, scale, cout, cin, k):
    """[cout, k*k, cin] (wgrad kernels) * scale -> [1, cout, cin, k, k] (reference weight layout)."""
    gw = torch.empty(1, cout, cin, k, k, dtype=torch.float32, device=dwk.device)
    with torch.cuda.device(dwk.device):
        rc = _lib.lib().sr_weight_grad_layout_f32(_lib.ptr(gw), _lib.ptr(dwk), float(scale), cout, cin, k * k,
                                                  _lib.stream_of(dwk))
    _lib.check(rc, "sr_weight_grad_layout_f32")
    return gw


class StyleScalesAll(Function):
    """forward(latent [B, L, K], cfg, *tensors) with tensors = (Wm_0, bm_0, wsq_0, Wm_1, ...), wsq_l = None for layers
    without demodulation; cfg = (latent_indices, mod_scale, lr_mul, eps).  Returns (s_0[, d_0], s_1[, d_1], ...)."""

    @staticmethod
    def forward(ctx, latent, cfg, *tensors):
        lat_idx, mod_scale, lr_mul, eps = cfg
        n = len(lat_idx)
        assert len(tensors) == 3 * n
        ctx.set_materialize_grads(False)
        lat = latent.contiguous()
        b, n_latent, k = lat.shape
        arr = (StyleLayer * n)()
        outs, keep = [], []
        for l in range(n):
            wm, bm, wsq = tensors[3 * l], tensors[3 * l + 1], tensors[3 * l + 2]
            wm, bm = wm.contiguous(), bm.contiguous()
            cin = wm.shape[0]
            s = torch.empty(b, cin, dtype=torch.float32, device=lat.device)
            a = arr[l]
            a.mod_weight, a.mod_bias, a.s = _lib.ptr(wm), _lib.ptr(bm), _lib.ptr(s)
            a.cin, a.latent_index = cin, int(lat_idx[l])
            outs.append(s)
            d = None
            if wsq is not None:
                wsq = wsq.contiguous()
                cout = wsq.shape[0]
                d = torch.empty(b, cout, dtype=torch.float32, device=lat.device)
                a.wsq, a.d, a.cout = _lib.ptr(wsq), _lib.ptr(d), cout
                outs.append(d)
            keep += [wm, bm, wsq, s, d]
        with torch.cuda.device(lat.device):
            rc = _lib.lib().sr_style_scales_forward_f32(arr, n, _lib.ptr(lat), b, n_latent, k, float(mod_scale), float(lr_mul),
                                                        float(eps), _lib.stream_of(lat))
        _lib.check(rc, "sr_style_scales_forward_f32")
        ctx.save_for_backward(lat, *[t for t in keep if t is not None])
        ctx.layout = [t is not None for t in keep]
        ctx.cfg = cfg
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        lat_idx, mod_scale, lr_mul, eps = ctx.cfg
        n = len(lat_idx)
        saved = list(ctx.saved_tensors)
        lat = saved.pop(0)
        keep = [saved.pop(0) if present else None for present in ctx.layout]
        b, n_latent, k = lat.shape
        dev = lat.device
        arr = (StyleLayer * n)()
        g_latent = torch.empty_like(lat)
        result, hold = [], []
        gi = 0
        for l in range(n):
            wm, bm, wsq, s, d = keep[5 * l:5 * l + 5]
            cin = wm.shape[0]
            g_s = grads[gi]; gi += 1
            g_d = None
            if wsq is not None:
                g_d = grads[gi]; gi += 1
            g_s = g_s.contiguous() if g_s is not None else None
            g_d = g_d.contiguous() if g_d is not None else None
            a = arr[l]
            a.mod_weight, a.mod_bias, a.s = _lib.ptr(wm), _lib.ptr(bm), _lib.ptr(s)
            a.cin, a.latent_index = cin, int(lat_idx[l])
            g_wm, g_bm = torch.empty_like(wm), torch.empty_like(bm)
            gs_total = torch.empty(b, cin, dtype=torch.float32, device=dev)
            a.g_s, a.g_mod_weight, a.g_mod_bias, a.gs_total = _lib.ptr(g_s), _lib.ptr(g_wm), _lib.ptr(g_bm), _lib.ptr(gs_total)
            g_wsq = None
            if wsq is not None:
                cout = wsq.shape[0]
                g_wsq = torch.empty_like(wsq)
                du = torch.empty(b, cout, dtype=torch.float32, device=dev)
                a.wsq, a.d, a.cout = _lib.ptr(wsq), _lib.ptr(d), cout
                a.g_d, a.g_wsq, a.du = _lib.ptr(g_d), _lib.ptr(g_wsq), _lib.ptr(du)
                hold.append(du)
            hold += [g_s, g_d, gs_total]
            result += [g_wm, g_bm, g_wsq]
        with torch.cuda.device(dev):
            rc = _lib.lib().sr_style_scales_backward_f32(arr, n, _lib.ptr(lat), _lib.ptr(g_latent), b, n_latent, k,
                                                         float(mod_scale), float(lr_mul), _lib.stream_of(lat))
        _lib.check(rc, "sr_style_scales_backward_f32")
        return (g_latent, None, *result)


def style_scales_all(latent, mods, lat_idx, wsqs=None):
    """mods: list of ModulatedConv2d; lat_idx: latent index per module; wsqs: optional precomputed weight statistics
    (WeightPrepAll) per module, None entries are computed here.  -> list of (s, d or None)."""
    m0 = mods[0].modulation
    assert all(m.modulation.scale == m0.scale and m.modulation.lr_mul == m0.lr_mul and m.modulation.activation is None
               for m in mods)
    tensors = []
    for j, m in enumerate(mods):
        wsq = None
        if m.demodulate:
            wsq = wsqs[j] if (wsqs is not None and wsqs[j] is not None) else WeightSq.apply(m.weight, m.scale)
        tensors += [m.modulation.weight, m.modulation.bias, wsq]
    cfg = (tuple(int(i) for i in lat_idx), m0.scale, m0.lr_mul, mods[0].eps)
    outs = list(StyleScalesAll.apply(latent, cfg, *tensors))
    res = []
    for m in mods:
        s = outs.pop(0)
        d = outs.pop(0) if m.demodulate else None
        res.append((s, d))
    return res
