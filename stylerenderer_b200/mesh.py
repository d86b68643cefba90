"""Mesh front-end of the rasteriser: what produces `rasterize`'s inputs.

Public names and argument meaning follow the reference (`LinearMorphableModel` face_model.py:4-74, `euler_mat`
utils_3d.py:43-80, `random_apply_pose3D` utils_3d.py:360-378, `mesh_point_normal` utils_3d.py:379-404) so callers and
checkpoints (`fc.weight`, `fc.bias`, `sigma`) carry over; the data path is laid out for the GPU:

  morphable model   one GEMM (cuBLAS through nn.Linear -- a plain library GEMM, [b, k] x [k, 3n])
  pose              the 3x4 matrices [s R | t] are built on the host from 7 numbers per sample (same draws, same op
                    order as the reference: bit-identical), applied by `sr_mesh_pose_apply_f32` (one thread per vertex)
  vertex normals    `sr_mesh_vertex_normals_f32`: one thread per face scatters its area-weighted normal, one per vertex
                    normalises -- instead of three sparse matrix products whose index tensors the reference builds
                    from a Python `range` on the host at every call
  normal maps       `normal_pyramid()`: pose -> normals -> the map at EVERY generator resolution as [b, 3, s, s] planes in
                    one C call (`sr_mesh_normal_pyramid_f32`, 5 launches), with no index / coefficient buffers in between
                    (reference model.py:260-270: seven independent rasterize calls, each writing int64 ids + fp32
                    coefficients and gathering the normals back through them)

CPU tensors and tensors that need gradients take plain differentiable torch ops (the reference's own formulation is
differentiable; the kernels here are forward-only, which is the state of the training loop: train.py:249-251 samples the
mesh under no_grad).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib

DEFAULT_POSE_SIGMA = (.5, .1, .05, .1, .1, .1, .15)        # yaw, pitch, roll, tx, ty, tz, log-scale (utils_3d.py:360)


# ------------------------------------------------------------------------------------------------ morphable model
def _vertex_rows(mean):
    """Any of the layouts the reference accepts for a mean shape ([3, n], [n, c >= 3], flat [3n]) -> rows of vertices."""
    a = np.asarray(mean, dtype=np.float32)
    if a.shape[0] == 3:
        return a.reshape(3, -1).T
    if a.ndim > 1:
        return a.reshape(-1, a.shape[-1])
    return a.reshape(-1, 3)


def _basis_rows(basis, want_rows, flat_len):
    """A deformation basis as [components, 3n]: the reference also accepts the transposed [3n, components] layout."""
    a = np.asarray(basis, dtype=np.float32)
    a = a.reshape(-1, a.shape[-1])
    return a.T if (a.shape[0] == flat_len and a.shape[1] >= want_rows) else a


def _sigma_list(spec, count):
    """`count` standard deviations from a scalar / short list: entry i, else the last one, else 1 (face_model.py:56-62)."""
    vals = [] if spec is None else [abs(float(x)) for x in np.reshape(spec, -1)]
    return [vals[i] if i < len(vals) else (vals[-1] if vals else 1.0) for i in range(count)]


class LinearMorphableModel(nn.Module):
    """vertices = mean + parameters @ basis as one `nn.Linear` (reference face_model.py:4-74: same constructor arguments,
    same random initialisation stream -- numpy first the mean, then the basis -- and the same state_dict keys)."""

    def __init__(self, vertices_num, shape_dim=0, expression_dim=0, vertices_mean=None, w_shape_numpy=None,
                 w_expression_numpy=None, sigma_shape=1, sigma_expression=.01, learnable=False):
        super().__init__()
        n = max(int(vertices_num), 1)
        dims = (max(int(shape_dim), 0), max(int(expression_dim), 0))
        k = sum(dims)
        spread = np.sqrt(k)
        # random fill first (this consumes numpy's global stream exactly like the reference), then the given pieces
        # (dtype flow as in face_model.py:16-19: the mean is drawn as float32 and scaled by numpy's float64 sqrt, the basis
        # is drawn in WHATEVER dtype that product has under the installed numpy's promotion rules)
        mean = (np.random.rand(3 * n).astype(np.float32) * 2 - 1) * spread
        basis = (np.random.rand(k, 3 * n).astype(mean.dtype) * 2 - 1) * spread
        if vertices_mean is not None:
            rows = _vertex_rows(vertices_mean)
            m = min(n, rows.shape[0])
            mean[:3 * m] = rows[:m, :3].reshape(-1)
        first = 0
        for given, dim in zip((w_shape_numpy, w_expression_numpy), dims):
            if given is not None and dim > 0:
                rows = _basis_rows(given, dim, 3 * n)
                d, m = min(dim, rows.shape[0]), min(n, rows.shape[1] // 3)
                basis[first:first + d, :3 * m] = rows[:d, :3 * m]
            first += dim
        self.dim = [dims[0], dims[1], 3 * n]
        self.fc = nn.Linear(k, 3 * n, bias=True)
        self.sigma = nn.Parameter(torch.tensor(_sigma_list(sigma_shape, dims[0]) + _sigma_list(sigma_expression, dims[1]),
                                               dtype=torch.float32), requires_grad=False)
        with torch.no_grad():
            self.fc.weight.copy_(torch.from_numpy(np.ascontiguousarray(basis.T)))
            self.fc.bias.copy_(torch.from_numpy(mean))
        self.fc.weight.requires_grad = self.fc.bias.requires_grad = bool(learnable)

    def random_input(self, batch_size=1):
        """Parameters ~ N(0, sigma^2) (face_model.py:69-70; drawn on sigma's device)."""
        return torch.normal(mean=0, std=self.sigma.unsqueeze(0).expand(batch_size, -1))

    def forward(self, x):
        return self.fc(x).reshape(-1, self.dim[2] // 3, 3)

    def regulation(self, x):
        return ((x / self.sigma.unsqueeze(0)) ** 2).sum()


# ------------------------------------------------------------------------------------------------ pose
_ROT_PLANE = {"x": (1, 2), "y": (2, 0), "z": (0, 1)}        # a rotation about an axis turns this coordinate plane


def _axis_rotation(c, s, axis):
    """[b] cosines / sines -> [b, 3, 3] rotation about `axis`: R[i,i] = R[j,j] = c, R[i,j] = -s, R[j,i] = s on its plane
    (i, j), identity elsewhere (the matrices of reference utils_3d.py:55-70)."""
    i, j = _ROT_PLANE[axis]
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    cell = {(i, i): c, (j, j): c, (i, j): -s, (j, i): s}
    rows = [torch.stack([cell.get((r, q), one if r == q else zero) for q in range(3)], -1) for r in range(3)]
    return torch.stack(rows, -2)


def euler_mat(angle, _type="yxz"):
    """Euler angles [b, 3] (or [3]) -> rotation matrices; letter i of `_type` is the axis of angle[:, i] and later
    rotations multiply from the left (reference utils_3d.py:43-80; unknown letters are skipped there too)."""
    single = angle.dim() == 1
    a = angle.reshape(1, -1) if single else angle
    c, s = torch.cos(a), torch.sin(a)
    total = None
    for i, letter in enumerate(_type[:3]):
        if letter.lower() not in _ROT_PLANE:
            continue
        r = _axis_rotation(c[:, i], s[:, i], letter.lower())
        total = r if total is None else torch.matmul(r, total)
    return total[0] if single else total


def pose_matrices(z):
    """[b, 7] pose draws (yaw, pitch, roll, tx, ty, tz, log-scale) -> [b, 3, 4] matrices [exp(z6) * R_yxz | t]."""
    return torch.cat((torch.exp(z[:, 6]).reshape(-1, 1, 1) * euler_mat(z[:, :3], "yxz"), z[:, 3:6].reshape(-1, 3, 1)), -1)


def apply_pose(v, T):
    """v[..., :3] . R + t per sample (R = T[:, :, :3] applied from the right, reference utils_3d.py:374-376)."""
    b = T.shape[0]
    if v.is_cuda and v.dtype == torch.float32 and not (torch.is_grad_enabled() and (v.requires_grad or T.requires_grad)):
        vc = v.detach().reshape(b, -1, v.shape[-1]).contiguous()
        pose = T.detach().to(device=v.device, dtype=torch.float32).reshape(b, 12).contiguous()
        out = torch.empty(b, vc.shape[1], 3, dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            rc = _lib.lib().sr_mesh_pose_apply_f32(_lib.ptr(out), _lib.ptr(vc), _lib.ptr(pose), b, vc.shape[1], vc.shape[2],
                                                   _lib.stream_of(vc))
        _lib.check(rc, "sr_mesh_pose_apply_f32")
        return out
    T = T.to(v.device)
    return torch.matmul(v[..., :3].reshape(b, -1, 3), T[:, :3, :3]) + T[:, :3, 3:].reshape(-1, 1, 3)


def random_pose_draws(batch, p=DEFAULT_POSE_SIGMA):
    """The reference's pose sampling (utils_3d.py:362-367): |p| padded to 7 entries, one normal draw per entry and sample
    on the CPU generator."""
    p = torch.abs(torch.as_tensor(p, dtype=torch.float32).reshape(-1)[:7])
    if p.numel() < 7:
        p = torch.cat((p, torch.zeros(7 - p.numel(), dtype=p.dtype)))
    return torch.normal(mean=0, std=p.unsqueeze(0).expand(batch, -1))


def random_apply_pose3D(p=DEFAULT_POSE_SIGMA, v=None):
    """Random rigid pose + scale of a batch of meshes (reference utils_3d.py:360-378).  v = None returns one 3x4 matrix."""
    batch = len(v) if v is not None and v.dim() >= 3 else 1
    T = pose_matrices(random_pose_draws(batch, p))
    return T[0] if v is None else apply_pose(v, T)


# ------------------------------------------------------------------------------------------------ vertex normals
def _vertex_normals_torch(v, tri):
    """Differentiable formulation on any device: face normals (b - a) x (c - a), summed into their corners, normalised
    with the norm clamped at 1e-8 (reference utils_3d.py:379-404 + layers.py:13-30)."""
    pos = v[..., :3]

    def face_normals(a, b, c):
        # component form with single-rounded products (torch.cross may fuse multiply-adds: faces that cancel exactly in
        # the reference -- a face listed with both windings -- would leave a residual of the other sign)
        p, q = b - a, c - a
        return torch.stack((p[..., 1] * q[..., 2] - p[..., 2] * q[..., 1], p[..., 2] * q[..., 0] - p[..., 0] * q[..., 2],
                            p[..., 0] * q[..., 1] - p[..., 1] * q[..., 0]), -1)
    vn = torch.zeros_like(pos)
    if tri.dim() == 2:
        fn = face_normals(*(pos[:, tri[:, j]] for j in range(3)))
        for j in range(3):
            vn = vn.index_add(1, tri[:, j], fn)
    else:
        idx = tri.unsqueeze(-1).expand(-1, -1, -1, 3)
        fn = face_normals(*(torch.gather(pos, 1, idx[:, :, j]) for j in range(3)))
        for j in range(3):
            vn = vn.scatter_add(1, idx[:, :, j], fn)
    return vn / vn.norm(dim=-1, keepdim=True).clamp_min(1e-8)


def mesh_point_normal(v, tri):
    """Area-weighted, L2-normalised vertex normals [b, n, 3] (v [b, n, >= 3], tri int64 [f, 3] or [b, f, 3]).  float32 CUDA
    vertices that need no gradient take the kernel; everything else the differentiable torch formulation."""
    needs_grad = torch.is_grad_enabled() and v.requires_grad
    if not v.is_cuda or v.dtype != torch.float32 or needs_grad:
        return _vertex_normals_torch(v, tri.to(v.device))
    vv = v[..., :3].detach().contiguous()
    t = tri.to(device=v.device, dtype=torch.int64).contiguous()
    b, n, _ = vv.shape
    out = torch.empty(b, n, 3, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        rc = _lib.lib().sr_mesh_vertex_normals_f32(_lib.ptr(out), _lib.ptr(vv), _lib.ptr(t), b, n, t.shape[-2],
                                                   1 if t.dim() == 2 else 0, 1e-8, _lib.stream_of(v))
    _lib.check(rc, "sr_mesh_vertex_normals_f32")
    return out


# ------------------------------------------------------------------------------------------------ fused front-end
class NormalMaps(list):
    """The normal maps [b, 3, s, s] of a mesh at the generator's resolutions 4, 8, ..., size (coarse to fine), e.g. from
    `normal_pyramid`.  GeneratorWithMap.forward accepts it in place of the (vertices, normals, triangles) tuple and then
    skips its own rasterisation."""


def normal_pyramid(verts, tri, sizes, pose=None, eps=1e-6):
    """(posed vertices [b,n,3], vertex normals [b,n,3], [normal map [b,3,s,s] for s in sizes]) in one C call
    (sr_mesh_normal_pyramid_f32): pose -> vertex normals -> multi-resolution rasterisation, forward only.
    verts [b, n, >= 3] float32 CUDA (e.g. LinearMorphableModel output), tri int64 [f, 3], pose [b, 3, 4] or None.
    Maps equal `rasterize(v, normals, tri, s).permute(0, 3, 1, 2)` of the reference chain (model.py:260-270)."""
    from .op.rasterize import MAX_LEVELS, RasterLevel
    _lib.require_cuda(verts, "normal_pyramid")
    if verts.dtype != torch.float32 or verts.dim() != 3 or tri.dim() != 2:
        raise RuntimeError("normal_pyramid: float32 vertices [b,n,3+] and one shared triangle list [f,3]")
    if not 1 <= len(sizes) <= MAX_LEVELS:
        raise RuntimeError(f"normal_pyramid: 1..{MAX_LEVELS} sizes")
    vc = verts.detach().contiguous()
    b, n, stride = vc.shape
    dev = vc.device
    t = tri.to(device=dev, dtype=torch.int64).contiguous()
    posed = torch.empty(b, n, 3, dtype=torch.float32, device=dev) if (pose is not None or stride != 3) else vc
    if pose is None and stride != 3:                        # extra per-vertex columns: identity pose drops them
        pose = torch.eye(3, 4, device=dev).expand(b, 3, 4)
    pm = pose.detach().to(device=dev, dtype=torch.float32).reshape(b, 12).contiguous() if pose is not None else None
    normals = torch.empty(b, n, 3, dtype=torch.float32, device=dev)
    maps = [torch.empty(b, 3, int(s), int(s), dtype=torch.float32, device=dev) for s in sizes]
    L = _lib.lib()
    csz = (ctypes.c_int64 * len(sizes))(*[int(s) for s in sizes])
    ws = torch.empty(L.sr_rasterize_pyramid_workspace_bytes(b, n, t.shape[0], len(sizes), csz) // 8 + 1, dtype=torch.int64, device=dev)
    arr = (RasterLevel * len(sizes))()
    for i, s in enumerate(sizes):
        arr[i].size, arr[i].out = int(s), _lib.ptr(maps[i])
    with torch.cuda.device(dev):
        rc = L.sr_mesh_normal_pyramid_f32(b, n, t.shape[0], _lib.ptr(vc), stride, _lib.ptr(pm), _lib.ptr(t), _lib.ptr(posed),
                                          _lib.ptr(normals), len(sizes), arr, _lib.ptr(ws), abs(float(eps)), _lib.stream_of(vc))
    _lib.check(rc, "sr_mesh_normal_pyramid_f32")
    return posed, normals, NormalMaps(maps)
