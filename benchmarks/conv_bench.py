#!/usr/bin/env python
"""tcgen05 modulated-conv kernels vs cuDNN (channels_last, TF32) at the layer shapes of G(256), B=32.
Prints one JSON line per (layer, op): ms, TFLOP/s, fraction of the measured dense bf16 peak / 2 (TF32 runs at
half the bf16 rate on sm_100; MEASURED_PEAKS.json has no TF32 entry, so the denominator is stated explicitly)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def peak_tf32():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = json.load(open(p))["bf16_tflops"] if os.path.exists(p) else 1590.0
    return bf16 / 2


_flush = None


def time_ms(fn, iters=8, warmup=2):
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    tot = 0.0
    for _ in range(iters):
        _flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from stylerenderer_b200 import tc_conv as tc
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    B, dev, peak, rows = args.batch, "cuda", peak_tf32(), []

    def emit(**r):
        r["TFLOPs"] = round(r["flop"] / r["ms"] / 1e9, 1)
        r["frac_tf32_peak"] = round(r["TFLOPs"] / peak, 3)
        if r.get("cudnn_ms"):
            r["speedup_vs_cudnn"] = round(r["cudnn_ms"] / r["ms"], 2)
        rows.append(r)
        print(json.dumps(r), flush=True)

    for cin, cout, r in [(512, 512, 4), (512, 512, 8), (512, 512, 16), (512, 512, 32), (512, 512, 64), (256, 256, 128),
                         (128, 128, 256)]:
        x = tc.modulate(torch.randn(B, r, r, cin, device=dev))
        w = torch.randn(cout, cin, 3, 3, device=dev)
        wm, wd = tc.weight_prep(w, 0.02, 0), tc.weight_prep(w, 0.02, 1)
        d = torch.rand(B, cout, device=dev) + 0.5
        bias = torch.randn(cout, device=dev)
        noise = torch.randn(B, r, r, device=dev)
        nw = torch.tensor([0.1], device=dev)
        out = torch.empty(B, r, r, cout, device=dev)
        flop = 2 * 9 * cin * cout * r * r * B
        xcl = x.permute(0, 3, 1, 2)                     # channels_last view for cuDNN
        wcl = w.contiguous(memory_format=torch.channels_last)
        ms = time_ms(lambda: tc.conv3x3(x, wm, out=out, epilogue=1, rowscale=d, bias=bias, noise=noise, noise_weight=nw))
        cms = time_ms(lambda: F.conv2d(xcl, wcl, padding=1))
        emit(op="conv3x3_fwd+styled_epilogue", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4), cudnn_ms=round(cms, 4))
        g = tc.modulate(torch.randn(B, r, r, cout, device=dev))
        dx = torch.empty(B, r, r, cin, device=dev)
        s = torch.rand(B, cin, device=dev) + 0.5
        ms = time_ms(lambda: tc.conv3x3(g, wd, out=dx, rowscale=s))
        emit(op="conv3x3_dgrad", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4))
        ms = time_ms(lambda: tc.wgrad3x3(g, x))
        gcl = g.permute(0, 3, 1, 2)
        cms = time_ms(lambda: torch.ops.aten.convolution_backward(gcl, xcl, wcl, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                                  [False, True, False]))
        emit(op="conv3x3_wgrad", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4), cudnn_ms=round(cms, 4))
        del x, g, out, dx
    for cin, cout, r in [(512, 512, 4), (512, 512, 8), (512, 512, 16), (512, 512, 32), (512, 256, 64), (256, 128, 128)]:
        x = tc.modulate(torch.randn(B, r, r, cin, device=dev))
        w = torch.randn(cout, cin, 3, 3, device=dev)
        wm = tc.weight_prep(w, 0.02, 0)
        d = torch.rand(B, cout, device=dev) + 0.5
        out = torch.empty(B, 2 * r + 1, 2 * r + 1, cout, device=dev)
        flop = 2 * 9 * cin * cout * r * r * B
        xcl = x.permute(0, 3, 1, 2)
        wt = w.transpose(0, 1).contiguous(memory_format=torch.channels_last)
        ms = time_ms(lambda: tc.conv_transpose3x3_s2(x, wm, out=out, rowscale=d))
        cms = time_ms(lambda: F.conv_transpose2d(xcl, wt, stride=2))
        emit(op="conv_transpose_s2_fwd(4 phases)", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4), cudnn_ms=round(cms, 4))
        g = tc.modulate(torch.randn(B, 2 * r + 1, 2 * r + 1, cout, device=dev))
        wg = tc.weight_prep(w, 0.02, 2)
        ms = time_ms(lambda: tc.conv3x3_s2_gather(g, wg, (r, r)))
        emit(op="conv_transpose_s2_dgrad", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4))
        ms = time_ms(lambda: tc.wgrad_transpose3x3_s2(g, x))
        emit(op="conv_transpose_s2_wgrad", cin=cin, cout=cout, res=r, flop=flop, ms=round(ms, 4))
        del x, g, out
    if args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
