#!/usr/bin/env python
"""Time one conv configuration (profiling aid): python benchmarks/conv_probe.py CIN COUT RES [up|dgrad_up]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stylerenderer_b200 import tc_conv as tc
from conv_bench import time_ms

cin, cout, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "plain"
B, dev = 32, "cuda"
x = tc.modulate(torch.randn(B, r, r, cin, device=dev))
w = torch.randn(cout, cin, 3, 3, device=dev)
d = torch.rand(B, cout, device=dev) + 0.5
flop = 2 * 9 * cin * cout * r * r * B
if kind == "plain":
    wm = tc.weight_prep(w, 0.02, 0)
    bias = torch.randn(cout, device=dev); noise = torch.randn(B, r, r, device=dev); nw = torch.tensor([0.1], device=dev)
    out = torch.empty(B, r, r, cout, device=dev); out2 = torch.empty_like(out)
    s2 = torch.rand(B, cout, device=dev)
    ms = time_ms(lambda: tc.conv3x3(x, wm, out=out, epilogue=1, rowscale=d, bias=bias, noise=noise, noise_weight=nw))
    ms2 = time_ms(lambda: tc.conv3x3(x, wm, out=out, epilogue=1, rowscale=d, bias=bias, noise=noise, noise_weight=nw, out2=out2, scale2=s2))
    ms0 = time_ms(lambda: tc.conv3x3(x, wm, out=out, rowscale=d))
    print(f"plain {cin}->{cout} @{r}: styled {ms:.4f} ms ({flop/ms/1e9:.0f} TF)  styled+out2 {ms2:.4f} ms  rowscale-only {ms0:.4f} ms "
          f"env HALO={os.environ.get('SR_CONV_HALO')} DEBUG={os.environ.get('SR_CONV_DEBUG')} SPLITX={os.environ.get('SR_HALO_SPLITX')}")
elif kind == "up":
    wm = tc.weight_prep(w, 0.02, 0)
    out = torch.empty(B, 2 * r + 1, 2 * r + 1, cout, device=dev)
    ms = time_ms(lambda: tc.conv_transpose3x3_s2(x, wm, out=out, rowscale=d))
    print(f"up {cin}->{cout} @{r}: {ms:.4f} ms ({flop/ms/1e9:.0f} TF) env HALO={os.environ.get('SR_CONV_HALO')} DEBUG={os.environ.get('SR_CONV_DEBUG')} SPLITX={os.environ.get('SR_HALO_SPLITX')}")
elif kind == "gather":
    g = tc.modulate(torch.randn(B, 2 * r + 1, 2 * r + 1, cout, device=dev))
    wg = tc.weight_prep(w, 0.02, 2)
    out = torch.empty(B, r, r, cin, device=dev)
    sc = torch.rand(B, cin, device=dev) + 0.5
    ms = time_ms(lambda: tc.conv3x3_s2_gather(g, wg, (r, r), out=out, rowscale=sc))
    print(f"gather(dgrad of up) {cin}<-{cout} @{r}: {ms:.4f} ms ({flop/ms/1e9:.0f} TF) env HALO={os.environ.get('SR_CONV_HALO')} DEBUG={os.environ.get('SR_CONV_DEBUG')}")
