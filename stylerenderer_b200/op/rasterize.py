"""rasterize behind the reference's signature (reference op/rasterize.py).

`rasterize_forward` / `rasterize_backward` mirror the reference's pybind functions
(`rasterize.forward` / `rasterize.backward`, reference op/rasterize.cpp:97-245).  The autograd
Function fuses what the reference does in Python around them: attribute interpolation runs inside
the resolve kernel, and the backward contracts d(coeff)/d(vertex) in registers and scatter-adds with
atomics instead of materialising a [b,h,w,3,9] tensor and a host-built sparse matrix
(reference op/rasterize.py:39-80).
"""
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
    ws = torch.empty(L.sr_rasterize_workspace_bytes(b, h, w, int(sfx == "f64")) // 8 + 1, dtype=torch.int64, device=dev)
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
