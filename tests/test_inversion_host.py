"""Host-side pieces of the latent-inversion harness (benchmarks/inversion.py, BASELINE.json configs[4]) that do not need
a GPU: the LPIPS-shaped perceptual distance (formula of reference lpips/networks_basic.py:64-92) and its use as an
optimisation target.  The generator half of the loop is covered by the GPU tests (frozen-weight backward)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))


def test_perceptual_distance_properties():
    import inversion
    P = inversion.PerceptualStack(seed=3)
    assert all(not p.requires_grad for p in P.parameters())
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1, torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    fa, fb = P.features(a), P.features(b)
    assert [f.shape[1] for f in fa] == [64, 128, 256, 512, 512]
    assert [f.shape[2] for f in fa] == [32, 16, 8, 4, 2]
    for f in fa:                                                       # unit-normalised over channels where non-zero
        n = (f * f).sum(1)
        assert float(((n - 1).abs() * (n > 1e-6)).max()) < 1e-4
    assert float(P.distance(fa, fa).abs().max()) == 0.0
    d_ab, d_ba = P.distance(fa, fb), P.distance(fb, fa)
    assert d_ab.shape == (2,) and bool((d_ab > 0).all())
    torch.testing.assert_close(d_ab, d_ba)
    # same seed -> same network (every rank of the harness builds it independently)
    P2 = inversion.PerceptualStack(seed=3)
    torch.testing.assert_close(P2.distance(P2.features(a), P2.features(b)), d_ab)


def test_perceptual_distance_drives_an_optimisation():
    import inversion
    P = inversion.PerceptualStack(seed=1)
    g = torch.Generator().manual_seed(1)
    target = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    tf = P.features(target)
    x = torch.zeros(1, 3, 32, 32, requires_grad=True)
    opt = torch.optim.Adam([x], lr=0.05)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = (P.distance(P.features(x), tf) + 0.1 * ((x - target) ** 2).mean((1, 2, 3))).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.7 * losses[0], losses
