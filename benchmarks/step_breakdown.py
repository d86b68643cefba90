#!/usr/bin/env python
"""Per-call device time of every stylerenderer_b200 C-ABI call inside one generator fwd+bwd step (B=32, 256 px).
Profiling aid: python benchmarks/step_breakdown.py [--batch 32]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def describe(name, a):
    try:
        if name == "sr_conv_igemm_multi_tf32":
            s = list(a[0])[0]
            kind = "up" if s.out_stride == 2 else ("s2gather" if s.in_stride == 2 else "plain")
            return (f"{kind} cin {s.cin} cout {s.cout} in {s.in_h}x{s.in_w} phases {a[1]} epi {s.epilogue} "
                    f"out2 {int(bool(s.out2))} rgb {int(bool(s.rgb_weight))}")
        if name == "sr_conv_wgrad_tf32":
            s = a[0]._obj
            return f"cin {s.cin} cout {s.cout} grid {s.grid_h}x{s.grid_w} g_stride {s.g_stride}"
        if name.startswith("sr_blur_nhwc"):
            return " ".join(str(int(v)) for v in a if isinstance(v, int) and 0 < v < 100000)[:60]
    except Exception as ex:                                     # noqa: BLE001
        return repr(ex)
    return ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    from stylerenderer_b200 import _lib, layers
    layers.set_conv_backend("tcgen05")
    dev = torch.device("cuda", 0)
    G = bench.build_generator(dev)
    z = torch.randn(args.batch, 512, device=dev)
    cot = torch.randn(args.batch, 3, 256, 256, device=dev)
    for _ in range(3):
        bench.generator_step(G, z, cot)
    names = ["sr_fused_bias_act_f32", "sr_fused_lrelu_backward_f32", "sr_upfirdn2d_f32"] + list(_lib.CONV_EXPORTS)
    with bench.KernelTimer(_lib, names) as kt:
        bench.generator_step(G, z, cot)
        torch.cuda.synchronize()
    tot = 0.0
    for n, a, s, e in kt.records:
        ms = s.elapsed_time(e)
        tot += ms
        print(f"{ms:8.4f} ms  {n:34s} {describe(n, a)}")
    print(f"{tot:8.4f} ms  total of {len(kt.records)} calls")


if __name__ == "__main__":
    main()
