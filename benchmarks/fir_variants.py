#!/usr/bin/env python
"""A/B timing of the channels-last FIR kernels (stream = TMA row-streaming ring, the default; sep / window = the per-thread
kernels behind SR_FIR_STREAM=0 [+ SR_FIR_SEP=0], see csrc/upfirdn2d.cu) on the generator's two largest
up-sampling blocks: forward tail (fir + noise + bias + lrelu + tf32 second output) and backward tail (fir^T * d -> tf32).
CUDA events, inputs larger than L2; prints one JSON line per (variant, shape)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    from stylerenderer_b200 import tc_conv as tc
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    k1 = torch.tensor([1., 3., 3., 1.], device=dev)
    taps = torch.outer(k1, k1) / 64 * 4
    B = 32
    for (r, c) in [(256, 128), (128, 256)]:
        t = torch.randn(B, r + 1, r + 1, c, device=dev)
        g = torch.randn(B, r, r, c, device=dev)
        noise, nw = torch.randn(B, 1, r, r, device=dev), torch.tensor([0.3], device=dev)
        bias, d = torch.randn(c, device=dev), torch.rand(B, c, device=dev) + 0.5
        by_f = 4 * B * c * ((r + 1) ** 2 + 2 * r * r)
        by_b = 4 * B * c * (r * r + (r + 1) ** 2)
        for variant in ("stream", "sep", "window"):
            os.environ.pop("SR_FIR_STREAM", None); os.environ.pop("SR_FIR_SEP", None)
            if variant != "stream":
                os.environ["SR_FIR_STREAM"] = "0"
            if variant == "window":
                os.environ["SR_FIR_SEP"] = "0"
            ms_f = timed(lambda: tc.blur_styled(t, taps, (1, 1), noise, nw, bias, 0.2, 2 ** 0.5, scale2=d))
            ms_b = timed(lambda: tc.blur_scaledot(g, taps, (2, 2), d))
            print(json.dumps({"kernel": variant, "res": r, "channels": c, "fwd_tail_ms": round(ms_f, 4),
                              "fwd_GBps": round(by_f / ms_f / 1e6, 1), "fwd_frac": round(by_f / ms_f / 1e6 / peak, 3),
                              "bwd_tail_ms": round(ms_b, 4), "bwd_GBps": round(by_b / ms_b / 1e6, 1),
                              "bwd_frac": round(by_b / ms_b / 1e6 / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
