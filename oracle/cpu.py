"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/sr_oracle.c (see that file's header).

Every wrapper takes / returns CPU torch tensors so parity tests read like the reference's calls:
  upfirdn2d(x, k, up, down, pad)            <-> op/upfirdn2d.py:145-157 (CPU branch, `upfirdn2d_native`)
  fused_bias_act(x, b, ref, act, grad, a, s) <-> op/fused_bias_act.cpp:5-30
  rasterize_forward(v, tri, h, w, persp, eps) <-> op/rasterize.cpp:97-178 (`rasterize.forward`)
  rasterize_backward(v, ind, persp, eps)      <-> op/rasterize.cpp:179-241 (`rasterize.backward`)
  rasterize(v, tex, tri, h, ...) / rasterize_grads(...) <-> op/rasterize.py:19-80 (autograd Function)
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsr_oracle.so")
_SRC = [os.path.join(_HERE, "sr_oracle.c"), os.path.join(_HERE, "raster_body.inc")]
_lib = None


def build(force: bool = False) -> str:
    """gcc the C restatement.  No -march/-ffast-math: one rounding per operation."""
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in _SRC)):
        return _LIB_PATH
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB_PATH,
                           _SRC[0], "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name in ("sr_oracle_rasterize_f32", "sr_oracle_rasterize_f64",
                     "sr_oracle_rasterize_dcoeff_f32", "sr_oracle_rasterize_dcoeff_f64"):
            getattr(_lib, name).restype = ctypes.c_int64
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_I64 = ctypes.c_int64


def upfirdn2d_raw(x4, k, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """x4: [major, in_h, in_w, minor] float32 -> [major, out_h, out_w, minor]."""
    x4 = x4.contiguous().float()
    k = k.contiguous().float()
    major, in_h, in_w, minor = x4.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1
    out = torch.empty(major, out_h, out_w, minor, dtype=torch.float32)
    lib().sr_oracle_upfirdn2d_f32(_p(out), _p(x4), _p(k), _I64(major), _I64(in_h), _I64(in_w), _I64(minor),
                                  _I64(kh), _I64(kw), _I64(up_x), _I64(up_y), _I64(down_x), _I64(down_y),
                                  _I64(px0), _I64(px1), _I64(py0), _I64(py1))
    return out


def upfirdn2d(x, k, up=1, down=1, pad=(0, 0)):
    """NCHW front-end with the reference signature (op/upfirdn2d.py:145)."""
    n, c, h, w = x.shape
    out = upfirdn2d_raw(x.reshape(n * c, h, w, 1), k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.view(n, c, out.shape[1], out.shape[2])


def fused_bias_act(x, bias, ref, act, grad, alpha, scale):
    x = x.contiguous().float()
    b = bias.contiguous().float() if bias is not None and bias.numel() else None
    r = ref.contiguous().float() if ref is not None and ref.numel() else None
    step_b = 1
    for d in x.shape[2:]:
        step_b *= d
    out = torch.empty_like(x)
    lib().sr_oracle_fused_bias_act_f32(_p(out), _p(x), _p(b), _p(r), ctypes.c_int(act), ctypes.c_int(grad),
                                       ctypes.c_float(alpha), ctypes.c_float(scale), _I64(x.numel()),
                                       _I64(step_b), _I64(b.numel() if b is not None else 1))
    return out


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """GPU-branch semantics of op/fused_act.py:86-97 (the slope argument is honoured)."""
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_backward(grad_out, out, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:20-41: (grad_input, grad_bias)."""
    gx = fused_bias_act(grad_out, None, out, 3, 1, negative_slope, scale)
    step_b = 1
    for d in gx.shape[2:]:
        step_b *= d
    c = gx.shape[1]
    gb = torch.empty(c, dtype=torch.float64)
    lib().sr_oracle_bias_grad_f32(_p(gb), _p(gx), _I64(gx.numel()), _I64(step_b), _I64(c))
    return gx, gb


def _raster_fn(base, dtype):
    return getattr(lib(), base + ("_f32" if dtype == torch.float32 else "_f64"))


def _creal(dtype, v):
    return ctypes.c_float(v) if dtype == torch.float32 else ctypes.c_double(v)


def rasterize_forward(v, tri, h, w=0, perspective=False, eps=1e-9):
    """-> (index int64 [b,h,w,3], coefficient [b,h,w,3]); shape rules of op/rasterize.cpp:103-124."""
    assert v.dtype in (torch.float32, torch.float64) and tri.dtype == torch.int64
    h = 1 if h <= 0 else h
    w = h if w <= 0 else w
    v = v.contiguous()
    tri = tri.contiguous()
    shared_v = v.dim() == 2
    shared_f = tri.dim() == 2
    b = 1 if shared_v else v.shape[0]
    if tri.dim() == 3 and (tri.shape[0] == b or shared_v):
        b = tri.shape[0]
    nv = v.shape[-2]
    nf = tri.shape[-2]
    lead = (h, w) if (shared_v and shared_f) else (b, h, w)
    ind = torch.zeros(*lead, 3, dtype=torch.int64)
    coeff = torch.zeros(*lead, 3, dtype=v.dtype)
    big = torch.finfo(v.dtype).max
    zbuf = torch.full(lead, -big, dtype=v.dtype)
    _raster_fn("sr_oracle_rasterize", v.dtype)(
        _I64(b), _I64(nv), _I64(nf), _I64(h), _I64(w), ctypes.c_int(shared_v), ctypes.c_int(shared_f),
        ctypes.c_int(bool(perspective)), _p(v), _p(tri), _p(ind), _p(coeff), _p(zbuf), _creal(v.dtype, abs(eps)))
    return ind, coeff, zbuf


def rasterize_backward(v, ind, perspective=False, eps=1e-9):
    """-> dcoeff [b,h,w,3,9]  (op/rasterize.cpp:179-241)."""
    v = v.contiguous()
    ind = ind.contiguous()
    if ind.dim() == 3:
        b, (h, w) = 1, ind.shape[:2]
    else:
        b, h, w = ind.shape[:3]
    n = v.shape[-2]
    dc = torch.zeros(*ind.shape, 9, dtype=v.dtype)
    _raster_fn("sr_oracle_rasterize_dcoeff", v.dtype)(
        _I64(b), _I64(n), _I64(h), _I64(w), ctypes.c_int(bool(perspective)), _p(v), _p(ind), _p(dc),
        _creal(v.dtype, abs(eps)))
    return dc


def rasterize(v, tex, tri, h=256, w=0, perspective=False, eps=1e-6):
    """Forward of op/rasterize.py:19-37 -> (out [b,h,w,c] or [b,h,w], ind, coeff)."""
    ind, coeff, _ = rasterize_forward(v, tri, h, w, perspective, eps)
    scalar_tex = tex.dim() == v.dim() - 1
    c = 1 if scalar_tex else tex.shape[-1]
    tex = tex.contiguous()
    out = torch.empty(*ind.shape[:-1], c, dtype=v.dtype)
    _raster_fn("sr_oracle_raster_interp", v.dtype)(_I64(ind.numel() // 3), _I64(c), _p(ind), _p(coeff), _p(tex), _p(out))
    return (out[..., 0] if scalar_tex else out), ind, coeff


def rasterize_grads(v, tex, ind, coeff, grad_out, perspective=False, eps=1e-6):
    """Backward of op/rasterize.py:39-80 -> (grad_v like v, grad_tex like tex), float64 accumulated."""
    dc = rasterize_backward(v, ind, perspective, eps)
    scalar_tex = tex.dim() == v.dim() - 1
    c = 1 if scalar_tex else tex.shape[-1]
    tex_c = tex.contiguous()
    g = grad_out.contiguous().to(v.dtype)
    gv = torch.zeros(v.numel(), dtype=torch.float64)
    gt = torch.zeros(tex.numel(), dtype=torch.float64)
    _raster_fn("sr_oracle_raster_scatter", v.dtype)(_I64(ind.numel() // 3), _I64(c), _p(ind), _p(coeff.contiguous()),
                                                    _p(dc), _p(tex_c), _p(g), _p(gv), _p(gt))
    return gv.view(v.shape).to(v.dtype), gt.view(tex.shape).to(tex.dtype)


def mesh_point_normal(v, tri):
    """<-> utils_3d.py:379-404 (`mesh_point_normal`): v [b, n, 3] float32, tri int64 [f, 3] -> unit normals [b, n, 3]."""
    v = v[..., :3].contiguous().float()
    tri = tri.contiguous().long()
    b, n, _ = v.shape
    out = torch.empty(b, n, 3, dtype=torch.float32)
    scratch = torch.empty(n * 3, dtype=torch.float32)
    lib().sr_oracle_vertex_normals_f32(_p(out), _p(v), _p(tri), _I64(b), _I64(n), _I64(tri.shape[0]), _p(scratch),
                                       ctypes.c_float(1e-8))
    return out
