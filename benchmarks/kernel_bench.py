#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the sizes G(256)/B=32 and config 3 use.

For every kernel: CUDA-event time per launch (L2 flushed between launches by writing a 512 MB buffer), achieved
algorithmic GB/s, fraction of the measured HBM peak -- and, when oracle/_ref was built (reference CUDA kernels
recompiled for sm_100a), the reference kernel timed the same way on the same inputs.
Prints one JSON object per line; `--out file` also writes them to a file (copied under profiles/)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


_flush = None


def time_ms(fn, iters=10, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    tot = 0.0
    for _ in range(iters):
        _flush.fill_(1)                                  # evict L2 (126 MB)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    from stylerenderer_b200 import op
    from stylerenderer_b200.op.fused_act import _lrelu_backward
    from oracle import build_ref
    ref = {}
    for n in ("ref_fused", "ref_upfirdn2d", "ref_rasterize"):
        try:
            ref[n] = build_ref.load(n)
        except Exception:
            ref[n] = None
    hbm = peaks()
    rows = []

    def emit(name, shape, by, ms, ref_ms=None, **kw):
        r = {"kernel": name, "shape": shape, "alg_bytes": by, "ms": round(ms, 4), "GBps": round(by / ms / 1e6, 1),
             "frac_hbm": round(by / ms / 1e6 / hbm, 3), "ref_ms": None if ref_ms is None else round(ref_ms, 4),
             "speedup_vs_ref_kernel": None if ref_ms is None else round(ref_ms / ms, 2)}
        r.update(kw)
        rows.append(r)
        print(json.dumps(r), flush=True)

    B = args.batch
    dev = "cuda"
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k4 = k4[None] * k4[:, None] / 64
    # ---- fused bias act fwd / bwd at every StyledConv output of G(256)
    for c, r in [(512, 4), (512, 8), (512, 16), (512, 32), (512, 64), (256, 128), (128, 256)]:
        x = torch.randn(B, c, r, r, device=dev)
        b = torch.randn(c, device=dev)
        n = x.numel()
        y = op.fused_leaky_relu(x, b)
        ms = time_ms(lambda: op.fused_bias_act(x, b, None, 3, 0, 0.2, 2 ** 0.5))
        rms = None
        if ref["ref_fused"]:
            e = x.new_empty(0)
            rms = time_ms(lambda: ref["ref_fused"].fused_bias_act(x, b, e, 3, 0, 0.2, 2 ** 0.5))
        emit("fused_bias_act_fwd", [B, c, r, r], 8 * n + 4 * c, ms, rms)
        g = torch.randn_like(x)
        ms = time_ms(lambda: _lrelu_backward(g, y, 0.2, 2 ** 0.5, True))
        rms = None
        if ref["ref_fused"]:
            e = x.new_empty(0)
            rms = time_ms(lambda: ref["ref_fused"].fused_bias_act(g, e, y, 3, 1, 0.2, 2 ** 0.5).sum((0, 2, 3)))
        emit("fused_lrelu_bwd+dbias", [B, c, r, r], 12 * n + 4 * c, ms, rms)
        del x, g, y
    # ---- upfirdn2d: blur after up-conv, skip upsample (x2), and their backward shapes
    for c, r in [(512, 8), (512, 16), (512, 32), (512, 64), (256, 128), (128, 256)]:
        x = torch.randn(B * c, r + 1, r + 1, 1, device=dev)
        by = 4 * B * c * ((r + 1) ** 2 + r * r)
        ms = time_ms(lambda: op.upfirdn2d_raw(x, k4 * 4, 1, 1, 1, 1, 1, 1, 1, 1))
        rms = time_ms(lambda: ref["ref_upfirdn2d"].upfirdn2d(x, k4 * 4, 1, 1, 1, 1, 1, 1, 1, 1)) if ref["ref_upfirdn2d"] else None
        emit("upfirdn2d_blur_fwd", [B * c, r + 1, r + 1], by, ms, rms)
        gy = torch.randn(B * c, r, r, 1, device=dev)
        ms = time_ms(lambda: op.upfirdn2d_raw(gy, k4 * 4, 1, 1, 1, 1, 2, 2, 2, 2))
        rms = time_ms(lambda: ref["ref_upfirdn2d"].upfirdn2d(gy, k4 * 4, 1, 1, 1, 1, 2, 2, 2, 2)) if ref["ref_upfirdn2d"] else None
        emit("upfirdn2d_blur_bwd", [B * c, r, r], by, ms, rms)
        del x, gy
    for r in (8, 32, 128):
        x = torch.randn(B * 3, r, r, 1, device=dev)
        by = 4 * B * 3 * (r * r + 4 * r * r)
        ms = time_ms(lambda: op.upfirdn2d_raw(x, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1))
        rms = time_ms(lambda: ref["ref_upfirdn2d"].upfirdn2d(x, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1)) if ref["ref_upfirdn2d"] else None
        emit("upfirdn2d_up2_fwd", [B * 3, r, r], by, ms, rms)
        gy = torch.randn(B * 3, 2 * r, 2 * r, 1, device=dev)
        ms = time_ms(lambda: op.upfirdn2d_raw(gy, k4 * 4, 1, 1, 2, 2, 1, 2, 1, 2))
        rms = time_ms(lambda: ref["ref_upfirdn2d"].upfirdn2d(gy, k4 * 4, 1, 1, 2, 2, 1, 2, 1, 2)) if ref["ref_upfirdn2d"] else None
        emit("upfirdn2d_down2_bwd_of_up2", [B * 3, 2 * r, 2 * r], by, ms, rms)
    # ---- rasterizer, config 3: 189x189 grid mesh, b = 64, 256x256
    from make_golden import grid_mesh, seeded
    b = 64
    v, tri = grid_mesh(189, b, 4242, jitter=0.002)
    tex = torch.nn.functional.normalize(seeded((b, 189 * 189, 3), 4243), dim=-1)
    v, tri, tex = v.cuda(), tri.cuda(), tex.cuda()
    n, f, h, c = 189 * 189, tri.shape[0], 256, 3
    by_f = b * (12 * n + h * h * (24 + 12) + 4 * n * c + 4 * h * h * c) + 24 * f
    ms = time_ms(lambda: op.rasterize(v, tex, tri, h))
    rms = None
    if ref["ref_rasterize"]:
        def ref_fwd():
            ind, coeff = ref["ref_rasterize"].forward(v, tri, h, 0, False, 1e-6)
            t = torch.index_select(tex.view(-1, c), 0, ind.view(-1))
            return torch.sum(t.view(b, h, h, 3, c) * coeff.unsqueeze(-1), -2)
        rms = time_ms(ref_fwd, iters=3, warmup=1)
    emit("rasterize_fwd(ids+bary+interp)", [b, n, f, h, h], by_f, ms, rms, images_per_s=round(b / ms * 1e3, 1))
    vg, tg = v.clone().requires_grad_(True), tex.clone().requires_grad_(True)
    out = op.rasterize(vg, tg, tri, h)
    go = torch.randn_like(out)
    by_b = b * (4 * h * h * c + 24 * h * h + 12 * h * h + 12 * n + 4 * n * c + 12 * n + 4 * n * c)
    ms = time_ms(lambda: torch.autograd.grad(out, (vg, tg), go, retain_graph=True))
    emit("rasterize_bwd(fused scatter)", [b, n, f, h, h], by_b, ms, None, images_per_s=round(b / ms * 1e3, 1))
    if args.out:
        with open(args.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
