#!/usr/bin/env python
"""GeneratorWithMap(256) forward+backward (the GAR generator: 7 rasterisations + style-map nets + StyledMapConv blocks),
batch 32, BFM-size mesh -- the generator half of BASELINE.json configs[3].  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--conv-backend", default="tcgen05")
    args = ap.parse_args()
    from stylerenderer_b200 import _lib, layers, mesh, model as M
    layers.set_conv_backend(args.conv_backend)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    G = M.GeneratorWithMap(256, 512, 8, channel_multiplier=2).to(dev)
    B = args.batch
    v, tex, tri = bench.synthetic_mesh(B)
    v, tri = v.to(dev), tri.to(dev)
    normals = mesh.mesh_point_normal(v, tri)
    z = torch.randn(B, 512, device=dev)
    cot = torch.randn(B, 3, 256, 256, device=dev)

    def step():
        for p in G.parameters():
            p.grad = None
        zz = z.detach().requires_grad_(True)
        img, _, _ = G([zz], (v, normals, tri))
        (img * cot).sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    print(json.dumps({"metric": "GeneratorWithMap fwd+bwd images/sec @256px", "value": round(B / ms * 1e3, 1), "ms_per_step": round(ms, 2),
                      "batch": B, "conv_backend": args.conv_backend, "execution": "eager",
                      "gpu_launches": (_lib.launch_count() - n0) // args.steps}))


if __name__ == "__main__":
    main()
