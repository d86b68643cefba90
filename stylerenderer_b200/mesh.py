"""Mesh front-end of the rasteriser: the producers of `rasterize`'s inputs,
with the reference's names and argument meaning.

  LinearMorphableModel   reference face_model.py:4-74   (one nn.Linear -> [b, n, 3]; the GEMM stays cuBLAS)
  euler_mat              reference utils_3d.py:43-80
  random_apply_pose3D    reference utils_3d.py:360-378  (tiny per-sample 3x4 transforms: plain torch)
  mesh_point_normal      reference utils_3d.py:379-404  -> sr_mesh_vertex_normals_f32 (csrc/mesh_ops.cu)
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib


class LinearMorphableModel(nn.Module):                # reference face_model.py:4-74
    def __init__(self, vertices_num, shape_dim=0, expression_dim=0, vertices_mean=None, w_shape_numpy=None,
                 w_expression_numpy=None, sigma_shape=1, sigma_expression=.01, learnable=False):
        super().__init__()
        vertices_num = max(int(vertices_num), 1)
        shape_dim, expression_dim = max(int(shape_dim), 0), max(int(expression_dim), 0)
        k = shape_dim + expression_dim
        v = (np.random.rand(vertices_num * 3).astype(np.float32) * 2 - 1) * np.sqrt(k)
        w = (np.random.rand(k, v.shape[0]).astype(v.dtype) * 2 - 1) * np.sqrt(k)
        if vertices_mean is not None:
            vm = np.array(vertices_mean, np.float32)
            vm = vm.reshape(3, -1).T if vm.shape[0] == 3 else (vm.reshape(-1, vm.shape[-1]) if vm.ndim > 1 else vm.reshape(-1, 3))
            n = min(vertices_num, vm.shape[0])
            v[:3 * n] = vm[:n, :3].reshape(-1)
        for wn, lo, dim in ((w_shape_numpy, 0, shape_dim), (w_expression_numpy, shape_dim, expression_dim)):
            if wn is None or dim == 0:
                continue
            wn = np.array(wn, np.float32)
            wn = wn.reshape((-1, wn.shape[-1]))
            if wn.shape[0] == w.shape[1] and wn.shape[1] >= dim:
                wn = wn.T
            d, n = min(dim, wn.shape[0]), min(vertices_num, wn.shape[1] // 3)
            w[lo:lo + d, :3 * n] = wn[:d, :3 * n]
        ss = [] if sigma_shape is None else np.reshape(sigma_shape, -1)
        se = [] if sigma_expression is None else np.reshape(sigma_expression, -1)
        self.dim = [shape_dim, expression_dim, vertices_num * 3]
        self.fc = nn.Linear(k, vertices_num * 3, bias=True)

        def sig(vals, i):
            return abs(vals[i]) if len(vals) > i else (abs(vals[-1]) if len(vals) > 0 else 1)
        self.sigma = nn.Parameter(torch.Tensor([sig(ss, i) for i in range(shape_dim)] + [sig(se, i) for i in range(expression_dim)]),
                                  requires_grad=False)
        with torch.no_grad():
            self.fc.weight.copy_(torch.from_numpy(w.T).float())
            self.fc.bias.copy_(torch.from_numpy(v).float())
        if not learnable:
            self.fc.weight.requires_grad = False
            self.fc.bias.requires_grad = False

    def random_input(self, batch_size=1):
        return torch.normal(mean=0, std=self.sigma.unsqueeze(0).expand(batch_size, -1))

    def forward(self, x):
        return torch.reshape(self.fc(x), (-1, self.dim[2] // 3, 3))

    def regulation(self, x):
        return ((x / self.sigma[np.newaxis, :]) ** 2).sum()


def euler_mat(angle, _type="yxz"):                    # reference utils_3d.py:43-80
    reshape = angle.dim() == 1
    if reshape:
        angle = angle.view(1, -1)
    c, s = torch.cos(angle), torch.sin(angle)
    one = torch.ones(len(c), 1, dtype=c.dtype, device=c.device)
    zero = torch.zeros(len(c), 1, dtype=c.dtype, device=c.device)
    T = None
    for i in range(3):
        ci, si = c[:, i:i + 1], s[:, i:i + 1]
        a = _type[i].lower()
        if a == "x":
            R = torch.cat((one, zero, zero, zero, ci, -si, zero, si, ci), -1).view(-1, 3, 3)
        elif a == "y":
            R = torch.cat((ci, zero, si, zero, one, zero, -si, zero, ci), -1).view(-1, 3, 3)
        elif a == "z":
            R = torch.cat((ci, -si, zero, si, ci, zero, zero, zero, one), -1).view(-1, 3, 3)
        else:
            continue
        T = R if T is None else torch.matmul(R, T)
    return T.view(3, 3) if reshape else T


def random_apply_pose3D(p=[.5, .1, .05, .1, .1, .1, .15], v=None):   # reference utils_3d.py:360-378
    """p = [yaw, pitch, roll, tx, ty, tz, scale] standard deviations; draws on the CPU generator like the reference."""
    batch = len(v) if v is not None and v.dim() >= 3 else 1
    if not isinstance(p, torch.Tensor):
        p = torch.Tensor(p)
    p = torch.abs(p.reshape(-1)[:7])
    if len(p) < 7:
        p = torch.cat((p, torch.zeros(7 - len(p), dtype=p.dtype, device=p.device)))
    z = torch.normal(mean=0, std=p.unsqueeze(0).expand(batch, -1))
    T = torch.cat((torch.exp(z[:, -1]).view(-1, 1, 1) * euler_mat(z[:, :3], "yxz"), z[:, 3:6].view(-1, 3, 1)), -1)
    if v is None:
        return T[0]
    if v.is_cuda:
        T = T.to(v.device)
    return torch.matmul(v[..., :3].view(batch, -1, 3), T[:, :3, :3]) + T[:, :3, 3:].view(-1, 1, 3)


def mesh_point_normal(v, tri):                        # reference utils_3d.py:379-404
    """Area-weighted, L2-normalised vertex normals [b, n, 3] of a triangle mesh (v [b, n, >=3], tri int64 [f, 3]).
    Forward only (the reference calls it under no_grad, train.py:249-251)."""
    _lib.require_cuda(v, "mesh_point_normal")
    if v.dtype != torch.float32:
        raise RuntimeError("mesh_point_normal: float32 vertices only")
    vv = v[..., :3].detach().contiguous()
    t = tri.to(device=v.device, dtype=torch.int64).contiguous()
    b, n, _ = vv.shape
    shared = t.dim() == 2
    nf = t.shape[-2]
    out = torch.empty(b, n, 3, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        rc = _lib.lib().sr_mesh_vertex_normals_f32(_lib.ptr(out), _lib.ptr(vv), _lib.ptr(t), b, n, nf, 1 if shared else 0,
                                                   1e-8, _lib.stream_of(v))
    _lib.check(rc, "sr_mesh_vertex_normals_f32")
    return out
