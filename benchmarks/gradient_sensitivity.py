#!/usr/bin/env python
"""How far is the reference's own fp32 implementation from its fp64 evaluation?  (CPU only, no GPU, ~20 s.)

oracle/torch_ref.Generator -- the CPU restatement of the reference modules, pinned to the real reference by
tests/test_oracle_pinning.py -- is run in float32 and in float64 on the same inputs; the image and every gradient are
compared (max-norm relative).  The image agrees to ~1e-6; gradients THROUGH the leaky-ReLUs do not: elements whose
pre-activation lies within the forward's rounding error of zero flip their mask, and reductions over few terms (noise
weights, biases, dz) move by up to a percent.  This is the envelope the network-level gradient tests use
(tests/parity_util.py::ENVELOPE); the record is profiles/r2_gradient_sensitivity.md."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
from make_golden import det_fill, seeded  # noqa: E402
from oracle import torch_ref as T  # noqa: E402


def run(G, z, cot, dtype):
    G = G.to(dtype)
    zz = z.to(dtype).clone().requires_grad_(True)
    img, _ = G([zz], randomize_noise=False)
    named = sorted(G.named_parameters())
    gr = torch.autograd.grad(img, [zz] + [p for _, p in named], cot.to(dtype), allow_unused=True)
    return img.detach().double(), [g.detach().double() if g is not None else None for g in gr], ["z"] + [n for n, _ in named]


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def report(title, G, z, cot):
    i32, g32, names = run(G, z, cot, torch.float32)
    i64, g64, _ = run(G, z, cot, torch.float64)
    errs = sorted(((rel(a, b), n) for a, b, n in zip(g32, g64, names) if a is not None and float(b.abs().max()) > 0), reverse=True)
    vals = [e for e, _ in errs]
    print(f"## {title}\nimage: {rel(i32, i64):.2e}; {len(vals)} gradient tensors: median {statistics.median(vals):.2e}, "
          f"{100 * sum(v <= 1e-3 for v in vals) / len(vals):.0f}% within 1e-3, max {vals[0]:.2e}; dz {dict((n, e) for e, n in errs)['z']:.2e}")
    for e, n in errs[:8]:
        print(f"  {e:.2e}  {n}")


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    G = det_fill(T.Generator(64, 64, 2), 1600).eval()
    report("Generator(64, 64, 2), det_fill parameters, batch 2 (tests/golden/make_golden_tc.py case)", G, seeded((2, 64), 1601),
           seeded((2, 3, 64, 64), 1602))
    torch.manual_seed(0)
    G = T.Generator(256, 512, 8, channel_multiplier=2)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("noise.weight") or n.endswith("activate.bias"):
                p.normal_(0, 0.1)
    report("Generator(256, 512, 8), default init, batch 2 (BASELINE.json configs[1]; test_headline_generator256_vs_oracle)", G.eval(),
           seeded((2, 512), 1740), seeded((2, 3, 256, 256), 1741))


if __name__ == "__main__":
    main()
