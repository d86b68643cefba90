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
    ap.add_argument("--no-graph", action="store_true")
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

    def timed(fn):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / args.steps

    for _ in range(3):
        step()
    n0 = _lib.launch_count()
    ms_eager = timed(step)
    launches = (_lib.launch_count() - n0) // args.steps
    # the same step captured once as a CUDA graph and replayed (the eager step is bound by ~10 ms of host enqueue)
    ms, execution = ms_eager, "eager"
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        ms, execution = timed(graph.replay), "cuda_graph_replay"
    assert all(p.grad is None or bool(torch.isfinite(p.grad).all()) for p in G.parameters()), "non-finite gradient"
    print(json.dumps({"metric": "GeneratorWithMap fwd+bwd images/sec @256px", "value": round(B / ms * 1e3, 1), "ms_per_step": round(ms, 2),
                      "eager_ms_per_step": round(ms_eager, 2), "batch": B, "conv_backend": args.conv_backend,
                      "execution": execution, "gpu_launches": launches}))


if __name__ == "__main__":
    main()
