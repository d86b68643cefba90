"""rasterize behind the reference's signature (reference op/rasterize.py).

`rasterize_forward` / `rasterize_backward` mirror the reference's pybind functions
(`rasterize.forward` / `rasterize.backward`, reference op/rasterize.cpp:97-245).  The autograd
Function fuses what the reference does in Python around them: attribute interpolation runs inside
the resolve kernel, and the backward contracts d(coeff)/d(vertex) in registers and scatter-adds with
atomics instead of materialising a [b,h,w,3,9] tensor and a host-built sparse matrix
(reference op/rasterize.py:39-80).
"""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib


def _suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise RuntimeError("rasterize: vertices must be float32 or float64")


def _problem(vertices, triangles, height, width):
    """Shape bookkeeping of the reference shim (reference op/rasterize.cpp:103-124)."""
    if triangles.dtype != torch.int64:
        raise RuntimeError("rasterize: triangles must be int64")
    if vertices.device != triangles.device:
        raise RuntimeError("rasterize: cuda input error (vertices and triangles on different devices)")
    h = 1 if height <= 0 else int(height)
    w = h if width <= 0 else int(width)
    if vertices.dim() not in (2, 3) or vertices.shape[-1] != 3:
        raise RuntimeError("rasterize: vertices input error (expected [b,n,3] or [n,3])")
    shared_v = vertices.dim() == 2
    b = 1 if shared_v else vertices.shape[0]
    nv = vertices.shape[-2]
    if triangles.dim() == 3 and triangles.shape[2] == 3 and (triangles.shape[0] == b or shared_v):
        b, shared_f = triangles.shape[0], False
    elif triangles.dim() == 2 and triangles.shape[1] == 3:
        shared_f = True
    else:
        raise RuntimeError("rasterize: triangles input error (expected [f,3] or [b,f,3])")
    nf = triangles.shape[-2]
    lead = (h, w) if (shared_v and shared_f) else (b, h, w)
    return b, nv, nf, h, w, shared_v, shared_f, lead


def _forward(vertices, triangles, height, width, perspective, eps, tex=None, c=0):
    _lib.require_cuda(vertices, "rasterize")
    sfx = _suffix(vertices.dtype)
    v = vertices.contiguous()
    tri = triangles.contiguous()
    b, nv, nf, h, w, shared_v, shared_f, lead = _problem(v, tri, height, width)
    dev = v.device
    ind = torch.empty(*lead, 3, dtype=torch.int64, device=dev)
    coeff = torch.empty(*lead, 3, dtype=v.dtype, device=dev)
    L = _lib.lib()
    ws = torch.empty(L.sr_rasterize_workspace_bytes(b, nv, nf, h, w, int(sfx == "f64")) // 8 + 1, dtype=torch.int64, device=dev)
    out = None
    if tex is not None:
        tex = tex.contiguous()
        out = torch.empty(*lead, c, dtype=v.dtype, device=dev)
    fn = getattr(L, "sr_rasterize_forward_" + sfx)
    with torch.cuda.device(dev):
        rc = fn(b, nv, nf, h, w, int(shared_v), int(shared_f), int(bool(perspective)), _lib.ptr(v), _lib.ptr(tri),
                _lib.ptr(ind), _lib.ptr(coeff), _lib.ptr(ws), abs(float(eps)), _lib.ptr(tex), c, _lib.ptr(out),
                _lib.stream_of(v))
    _lib.check(rc, "sr_rasterize_forward_" + sfx)
    return ind, coeff, out


def rasterize_forward(vertices, triangles, height, width=0, perspective=False, eps=1e-9):
    """Mirror of the reference pybind `rasterize.forward` -> [index int64 [b,h,w,3], coefficient [b,h,w,3]]."""
    ind, coeff, _ = _forward(vertices, triangles, height, width, perspective, eps)
    return [ind, coeff]


def rasterize_backward(vertices, index, perspective=False, eps=1e-9):
    """Mirror of the reference pybind `rasterize.backward` -> dcoeff [b,h,w,3,9] (reference op/rasterize.cpp:179-241)."""
    _lib.require_cuda(vertices, "rasterize_backward")
    sfx = _suffix(vertices.dtype)
    v = vertices.contiguous()
    ind = index.contiguous()
    if ind.dim() == 3:
        b, (h, w) = 1, ind.shape[:2]
    else:
        b, h, w = ind.shape[:3]
    n = v.shape[-2]
    dc = torch.zeros(*ind.shape, 9, dtype=v.dtype, device=v.device)
    fn = getattr(_lib.lib(), "sr_rasterize_dcoeff_" + sfx)
    with torch.cuda.device(v.device):
        rc = fn(b, n, h, w, int(bool(perspective)), _lib.ptr(v), _lib.ptr(ind), _lib.ptr(dc), abs(float(eps)),
                _lib.stream_of(v))
    _lib.check(rc, "sr_rasterize_dcoeff_" + sfx)
    return dc


class Rasterize(Function):                          # reference op/rasterize.py:18-80
    @staticmethod
    def forward(ctx, v, tex, tri, h, w, perspective, eps):
        scalar_tex = tex.dim() == v.dim() - 1
        c = 1 if scalar_tex else int(tex.shape[-1])
        tex = tex.to(v.dtype)
        ind, coeff, out = _forward(v, tri, h, w, perspective, eps, tex=tex, c=c)
        ctx.save_for_backward(v, tex, ind, coeff)
        ctx.perspective, ctx.eps, ctx.c, ctx.scalar_tex = perspective, eps, c, scalar_tex
        ctx.mark_non_differentiable(ind)
        return (out[..., 0] if scalar_tex else out), ind, coeff

    @staticmethod
    def backward(ctx, grad_out, _grad_ind, _grad_coeff):
        v, tex, ind, coeff = ctx.saved_tensors
        need_v, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_v or need_t):
            return (None,) * 7
        sfx = _suffix(v.dtype)
        vc = v.contiguous()
        texc = tex.contiguous()
        if ind.dim() == 3:
            b, (h, w) = 1, ind.shape[:2]
        else:
            b, h, w = ind.shape[:3]
        n = vc.shape[-2]
        g = grad_out.to(v.dtype).contiguous()
        grad_v = torch.zeros_like(vc) if need_v else None
        grad_t = torch.zeros_like(texc) if need_t else None
        fn = getattr(_lib.lib(), "sr_rasterize_backward_" + sfx)
        with torch.cuda.device(v.device):
            rc = fn(b, n, h, w, ctx.c, int(bool(ctx.perspective)), _lib.ptr(vc), _lib.ptr(texc), _lib.ptr(ind),
                    _lib.ptr(coeff), _lib.ptr(g), _lib.ptr(grad_v), _lib.ptr(grad_t), abs(float(ctx.eps)),
                    _lib.stream_of(v))
        _lib.check(rc, "sr_rasterize_backward_" + sfx)
        return grad_v, grad_t, None, None, None, None, None


def rasterize(v, tex, tri, h=256, w=0, perspective=False, eps=1e-6, return_buffers=False):
    """reference op/rasterize.py:81-82: -> [b,h,w,c] (channels-last) interpolated attributes.
    `return_buffers=True` additionally returns the (index, coefficient) buffers."""
    _lib.require_cuda(v, "rasterize")
    out, ind, coeff = Rasterize.apply(v, tex, tri, h, w, perspective, eps)
    return (out, ind, coeff) if return_buffers else out


# ---------------------------------------------------------------------------------------------------------------------
# Resolution pyramid: the same mesh at several square sizes in one triangle pass / resolve pass / scatter pass
# (sr_rasterize_pyramid_*; GeneratorWithMap renders 4, 8, ..., 256 every forward, reference model.py:260-270).
MAX_LEVELS = 8


class RasterLevel(ctypes.Structure):                 # mirrors `sr_raster_level` in include/stylerenderer_b200.h
    _fields_ = [("size", ctypes.c_int64), ("ids", ctypes.c_void_p), ("bary", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("gout", ctypes.c_void_p)]


def _levels(sizes, inds, coeffs, outs=None, gouts=None):
    arr = (RasterLevel * len(sizes))()
    for i, sz in enumerate(sizes):
        a = arr[i]
        a.size, a.ids, a.bary = int(sz), _lib.ptr(inds[i]), _lib.ptr(coeffs[i])
        a.out = _lib.ptr(outs[i]) if outs is not None else None
        a.gout = _lib.ptr(gouts[i]) if gouts is not None else None
    return arr


class RasterizePyramid(Function):
    """(v, tex, tri) -> one interpolated map [b,s,s,c] per size s, each bit-identical to `Rasterize` at that size."""

    @staticmethod
    def forward(ctx, v, tex, tri, sizes, perspective, eps):
        _lib.require_cuda(v, "rasterize_pyramid")
        if v.dtype != torch.float32:
            raise RuntimeError("rasterize_pyramid: float32 vertices only (use rasterize() per size for float64)")
        if not 1 <= len(sizes) <= MAX_LEVELS:
            raise RuntimeError(f"rasterize_pyramid: 1..{MAX_LEVELS} sizes")
        ctx.set_materialize_grads(False)
        scalar_tex = tex.dim() == v.dim() - 1
        c = 1 if scalar_tex else int(tex.shape[-1])
        tex = tex.to(v.dtype).contiguous()
        vc, tric = v.contiguous(), tri.contiguous()
        dev = vc.device
        inds, coeffs, outs = [], [], []
        b = nv = nf = shared_v = shared_f = None
        for sz in sizes:
            b, nv, nf, h, w, shared_v, shared_f, lead = _problem(vc, tric, sz, sz)
            inds.append(torch.empty(*lead, 3, dtype=torch.int64, device=dev))
            coeffs.append(torch.empty(*lead, 3, dtype=vc.dtype, device=dev))
            outs.append(torch.empty(*lead, c, dtype=vc.dtype, device=dev))
        L = _lib.lib()
        csz = (ctypes.c_int64 * len(sizes))(*[int(s) for s in sizes])
        ws = torch.empty(L.sr_rasterize_pyramid_workspace_bytes(b, nv, nf, len(sizes), csz) // 8 + 1, dtype=torch.int64, device=dev)
        arr = _levels(sizes, inds, coeffs, outs)
        with torch.cuda.device(dev):
            rc = L.sr_rasterize_pyramid_forward_f32(b, nv, nf, len(sizes), arr, int(shared_v), int(shared_f),
                                                    int(bool(perspective)), _lib.ptr(vc), _lib.ptr(tric), _lib.ptr(ws),
                                                    abs(float(eps)), _lib.ptr(tex), c, _lib.stream_of(vc))
        _lib.check(rc, "sr_rasterize_pyramid_forward_f32")
        ctx.save_for_backward(vc, tex, *inds, *coeffs)
        ctx.cfg = (tuple(int(s) for s in sizes), bool(perspective), float(eps), c, b)
        return tuple(o[..., 0] if scalar_tex else o for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        sizes, perspective, eps, c, b = ctx.cfg
        saved = ctx.saved_tensors
        vc, tex = saved[0], saved[1]
        n_lv = len(sizes)
        inds, coeffs = saved[2:2 + n_lv], saved[2 + n_lv:2 + 2 * n_lv]
        need_v, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_v or need_t) or all(g is None for g in grads):
            return (None,) * 6
        gouts = [g.to(vc.dtype).contiguous() if g is not None else None for g in grads]
        grad_v = torch.zeros_like(vc) if need_v else None
        grad_t = torch.zeros_like(tex) if need_t else None
        arr = _levels(sizes, inds, coeffs, None, gouts)
        with torch.cuda.device(vc.device):
            rc = _lib.lib().sr_rasterize_pyramid_backward_f32(b, vc.shape[-2], n_lv, arr, c, int(perspective), _lib.ptr(vc),
                                                              _lib.ptr(tex), _lib.ptr(grad_v), _lib.ptr(grad_t),
                                                              abs(eps), _lib.stream_of(vc))
        _lib.check(rc, "sr_rasterize_pyramid_backward_f32")
        return grad_v, grad_t, None, None, None, None


def rasterize_pyramid(v, tex, tri, sizes, perspective=False, eps=1e-6):
    """[rasterize(v, tex, tri, s) for s in sizes] in three launches instead of 3 * len(sizes) (same values, bit for bit)."""
    return list(RasterizePyramid.apply(v, tex, tri, tuple(sizes), perspective, eps))


def rasterize_pyramid_maps(v, tex, tri, sizes, perspective=False, eps=1e-6, planar=True):
    """Forward-only pyramid: only the interpolated maps, no index / coefficient buffers (36 of the 48 bytes per pixel the
    full forward writes), as [b, c, s, s] planes when `planar` (NCHW, what the style-map nets consume).  Same values as
    `rasterize_pyramid`.  For meshes that need no gradient -- the training loop samples them under no_grad (reference
    train.py:249-251)."""
    _lib.require_cuda(v, "rasterize_pyramid_maps")
    if v.dtype != torch.float32 or tex.dim() != v.dim():
        raise RuntimeError("rasterize_pyramid_maps: float32 vertices and [.., n, c] attributes only")
    if not 1 <= len(sizes) <= MAX_LEVELS:
        raise RuntimeError(f"rasterize_pyramid_maps: 1..{MAX_LEVELS} sizes")
    c = int(tex.shape[-1])
    vc, texc, tric = v.detach().contiguous(), tex.detach().to(v.dtype).contiguous(), tri.contiguous()
    dev = vc.device
    outs = []
    b = nv = nf = shared_v = shared_f = None
    for sz in sizes:
        b, nv, nf, h, w, shared_v, shared_f, lead = _problem(vc, tric, sz, sz)
        shape = (*lead[:-2], c, sz, sz) if planar else (*lead, c)
        outs.append(torch.empty(shape, dtype=vc.dtype, device=dev))
    L = _lib.lib()
    csz = (ctypes.c_int64 * len(sizes))(*[int(s) for s in sizes])
    ws = torch.empty(L.sr_rasterize_pyramid_workspace_bytes(b, nv, nf, len(sizes), csz) // 8 + 1, dtype=torch.int64, device=dev)
    arr = (RasterLevel * len(sizes))()
    for i, sz in enumerate(sizes):
        arr[i].size, arr[i].out = int(sz), _lib.ptr(outs[i])
    with torch.cuda.device(dev):
        rc = L.sr_rasterize_pyramid_maps_f32(b, nv, nf, len(sizes), arr, int(shared_v), int(shared_f), int(bool(perspective)),
                                             _lib.ptr(vc), _lib.ptr(tric), _lib.ptr(ws), abs(float(eps)), _lib.ptr(texc), c,
                                             1 if planar else 0, _lib.stream_of(vc))
    _lib.check(rc, "sr_rasterize_pyramid_maps_f32")
    return outs
