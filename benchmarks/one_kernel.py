#!/usr/bin/env python
"""Run ONE kernel a few times (for `ncu --set full` captures): python benchmarks/one_kernel.py <name> [iters]
names: conv512_64 | conv128_256 | wgrad512_64 | wgrad128_256 | up256_128 | prologue | blur_nchw | blur_nhwc | bias_act | raster"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
from stylerenderer_b200 import op, tc_conv as tc  # noqa: E402
B, dev = 32, "cuda"
if name in ("conv512_64", "conv128_256", "wgrad512_64", "wgrad128_256"):
    cin, cout, r = (512, 512, 64) if "512" in name else (128, 128, 256)
    x = tc.modulate(torch.randn(B, r, r, cin, device=dev))
    w = torch.randn(cout, cin, 3, 3, device=dev)
    wm = tc.weight_prep(w, 0.02, 0)
    d = torch.rand(B, cout, device=dev) + 0.5
    bias, noise, nw = torch.randn(cout, device=dev), torch.randn(B, r, r, device=dev), torch.tensor([0.1], device=dev)
    out = torch.empty(B, r, r, cout, device=dev)
    g = tc.modulate(torch.randn(B, r, r, cout, device=dev))
    for _ in range(iters):
        if name.startswith("wgrad"):
            tc.wgrad3x3(g, x)
        else:
            tc.conv3x3(x, wm, out=out, epilogue=1, rowscale=d, bias=bias, noise=noise, noise_weight=nw)
elif name == "up256_128":
    cin, cout, r = 256, 128, 128
    x = tc.modulate(torch.randn(B, r, r, cin, device=dev))
    wm = tc.weight_prep(torch.randn(cout, cin, 3, 3, device=dev), 0.02, 0)
    d = torch.rand(B, cout, device=dev) + 0.5
    out = torch.empty(B, 2 * r + 1, 2 * r + 1, cout, device=dev)
    for _ in range(iters):
        tc.conv_transpose3x3_s2(x, wm, out=out, rowscale=d)
elif name == "prologue":
    c, r = 128, 256
    y, gxs = torch.randn(B, r, r, c, device=dev), torch.randn(B, r, r, c, device=dev)
    noise, nw = torch.randn(B, 1, r, r, device=dev), torch.tensor([0.1], device=dev)
    bias, d, sn = torch.randn(c, device=dev), torch.rand(B, c, device=dev) + 0.5, torch.rand(B, c, device=dev) + 0.5
    for _ in range(iters):
        tc.bwd_prologue2(y, noise, nw, bias, d, 0.2, 2 ** 0.5, True, gxs=gxs, s_next=sn)
elif name == "blur_nchw":
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = k[None] * k[:, None] / 16
    x = torch.randn(B * 128, 257, 257, 1, device=dev)
    for _ in range(iters):
        op.upfirdn2d_raw(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
elif name == "blur_nhwc":
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = k[None] * k[:, None] / 16
    x = torch.randn(B, 257, 257, 128, device=dev)
    for _ in range(iters):
        op.upfirdn2d_raw(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
elif name == "bias_act":
    x, b = torch.randn(B, 128, 256, 256, device=dev), torch.randn(128, device=dev)
    for _ in range(iters):
        op.fused_bias_act(x, b, None, 3, 0, 0.2, 2 ** 0.5)
elif name == "raster":
    from make_golden import grid_mesh, seeded
    v, tri = grid_mesh(189, 64, 4242, jitter=0.002)
    tex = torch.nn.functional.normalize(seeded((64, 189 * 189, 3), 4243), dim=-1)
    v, tri, tex = v.cuda().requires_grad_(True), tri.cuda(), tex.cuda().requires_grad_(True)
    for _ in range(iters):
        out = op.rasterize(v, tex, tri, 256)
        out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("done", name)
