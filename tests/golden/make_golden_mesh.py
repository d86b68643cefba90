#!/usr/bin/env python
"""Generate tests/golden/mesh_golden.pt by running the UNMODIFIED reference mesh front-end (utils_3d.py, face_model.py
imported from /root/reference).  Authoring container only; the fixture is committed."""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, grid_mesh, seeded  # noqa: E402


def main():
    sys.path.insert(0, REF)
    import layers  # noqa
    layers.math = math
    import face_model  # noqa
    import utils_3d  # noqa
    out = {}
    v, tri = grid_mesh(24, 3, 811)
    out["normals_grid24"] = dict(v=v, tri=tri, n=utils_3d.mesh_point_normal(v, tri))
    # a soup with a degenerate (zero-area) face, an isolated vertex and repeated corners
    v2 = seeded((2, 9, 3), 812)
    tri2 = torch.tensor([[0, 1, 2], [2, 1, 3], [3, 3, 4], [5, 6, 7], [7, 6, 5], [0, 2, 4]])
    out["normals_soup"] = dict(v=v2, tri=tri2, n=utils_3d.mesh_point_normal(v2, tri2))
    torch.manual_seed(813)
    posed = utils_3d.random_apply_pose3D([.5, .1, .05, .1, .1, .1, .15], v)
    out["pose"] = dict(seed=813, v=v, posed=posed)
    ang = seeded((4, 3), 814)
    out["euler"] = dict(angle=ang, yxz=utils_3d.euler_mat(ang, "yxz"), xyz=utils_3d.euler_mat(ang, "xyz"))
    np.random.seed(815)
    torch.manual_seed(815)
    mean = seeded((50, 3), 816).numpy()
    lm = face_model.LinearMorphableModel(50, 6, 4, vertices_mean=mean)
    x = lm.random_input(3)
    out["morphable"] = dict(seed=815, mean=torch.from_numpy(mean), x=x, verts=lm(x).detach(), sigma=lm.sigma.detach().clone(),
                            weight=lm.fc.weight.detach().clone(), bias=lm.fc.bias.detach().clone())
    path = os.path.join(HERE, "mesh_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
