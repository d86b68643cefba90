"""CPU restatements of the INDEXING logic of two kernels (no GPU, no arithmetic parity claims -- those are the `-m gpu`
tests): there is no GPU in the authoring container, so loop structure / ring rotation / block-to-level look-ups are checked
here first ("keep a CPU restatement of each kernel's indexing logic for fast local checks").

* `fir_nhwc_ring_kernel` (csrc/upfirdn2d.cu): input rows stream through a ring of the 4 open output rows; rank-1 taps use
  the separable form through the largest tap.  The model walks exactly the kernel's loops for one channel.
* `raster_resolve_pyramid_kernel` / `raster_backward_pyramid_kernel` (csrc/rasterize.cu): the concatenated 256-pixel blocks
  of all levels and the CTA-uniform `while (blk >= blk_off[l + 1]) ++l` look-up.
"""
import numpy as np
import pytest

K = 4


def fir_reference(x, taps, pad):
    """upfirdn2d with up = down = 1 (reference op/upfirdn2d.py:159-200): zero padding, TRUE convolution (flipped taps)."""
    ih, iw = x.shape
    xp = np.pad(x, ((pad, pad), (pad, pad)))
    oh, ow = ih + 2 * pad - K + 1, iw + 2 * pad - K + 1
    kf = taps[::-1, ::-1]
    return np.array([[(xp[y:y + K, c:c + K] * kf).sum() for c in range(ow)] for y in range(oh)])


def fir_ring_model(x, taps, pad, rows_per_strip, depth):
    """One channel of fir_nhwc_ring_kernel<MODE 0, D = depth>: thread = (strip ys, column pair xp)."""
    ih, iw = x.shape
    oh, ow = ih + 2 * pad - K + 1, iw + 2 * pad - K + 1
    out = np.full((oh, ow), np.nan)
    tk = np.array([[taps[K - 1 - a][K - 1 - b] for b in range(K)] for a in range(K)])
    pa, pb = np.unravel_index(np.argmax(np.abs(tk)), tk.shape)          # first largest tap, like the strict '>' scan
    kv, kh = tk[:, pb].copy(), tk[pa, :] / tk[pa, pb]
    best = np.abs(tk).max()
    sep = best > 0 and bool((np.abs(np.outer(kv, kh) - tk) <= 1e-6 * best).all())

    def load_row(iy, ix0):
        v = np.zeros(K + 1)
        if 0 <= iy < ih:
            for b in range(K + 1):
                if 0 <= ix0 + b < iw:
                    v[b] = x[iy, ix0 + b]
        return v
    for ys in range((oh + rows_per_strip - 1) // rows_per_strip):
        for xp in range((ow + 1) // 2):
            ox0, oy0 = xp * 2, ys * rows_per_strip
            ix0, oy1 = ox0 - pad, min(oh, ys * rows_per_strip + rows_per_strip)
            acc = np.zeros((K, 2))
            nsteps = oy1 - oy0 + K - 1
            buf = [load_row(oy0 - pad + u, ix0) if u < nsteps else None for u in range(depth)]
            for r0 in range(0, nsteps, depth):
                for u in range(depth):
                    r = r0 + u
                    if r >= nsteps:
                        continue
                    v, oy = buf[u], oy0 + r - (K - 1)
                    for j in range(2):
                        if sep:
                            h = sum(v[j + b] * kh[b] for b in range(K))
                            for a in range(K):
                                acc[K - 1 - a][j] += h * kv[a]
                        else:
                            for a in range(K):
                                for b in range(K):
                                    acc[K - 1 - a][j] += v[j + b] * tk[a][b]
                    if oy >= oy0:
                        out[oy, ox0] = acc[0][0]
                        if ox0 + 1 < ow:
                            out[oy, ox0 + 1] = acc[0][1]
                    acc[:-1] = acc[1:].copy()
                    acc[-1] = 0
                    if r + depth < nsteps:
                        buf[u] = load_row(oy0 - pad + r + depth, ix0)
    return out, sep


@pytest.mark.parametrize("depth", [1, 2, 3, 4])
@pytest.mark.parametrize("geom", [(9, 9, 1, 4), (8, 8, 2, 4), (17, 13, 1, 8), (33, 33, 2, 16), (5, 7, 1, 4), (4, 3, 1, 4)])
def test_fir_ring_kernel_indexing(geom, depth):
    ih, iw, pad, rows = geom
    rng = np.random.default_rng(ih * 100 + iw)
    x = rng.standard_normal((ih, iw))
    k1 = np.array([1., 3., 3., 1.])
    cases = [(np.outer(k1, k1) / 64 * 4, True),                         # the model's FIR (reference layers.py:7-12)
             (np.outer([1., 2., 3., 4.], [4., 3., -2., 1.]), True),     # asymmetric rank-1: flips and factor order matter
             (rng.standard_normal((K, K)), False)]                      # general taps: 2-D form
    for taps, expect_sep in cases:
        got, sep = fir_ring_model(x, taps, pad, rows, depth)
        assert sep == expect_sep
        assert not np.isnan(got).any(), "every output pixel is written exactly by one thread"
        np.testing.assert_allclose(got, fir_reference(x, taps, pad), rtol=0, atol=1e-12)


@pytest.mark.parametrize("b,sizes", [(1, [4]), (3, [4, 8, 16, 32, 64, 128, 256]), (2, [1, 7, 33, 128]), (5, [16, 16, 3]),
                                     (32, [4, 8, 16, 32, 64, 128, 256, 5])])
def test_raster_pyramid_block_table(b, sizes):
    threads = 256
    key_off, blk_off, keys = [], [0], 0
    for s in sizes:                                                     # build_pyramid()
        npix = b * s * s
        key_off.append(keys)
        blk_off.append(blk_off[-1] + (npix + threads - 1) // threads)
        keys += npix
    seen = np.zeros(keys, dtype=np.int32)
    grid = min(blk_off[-1], 148 * 8)
    for cta in range(grid):                                             # raster_resolve_pyramid_kernel's block walk
        for blk in range(cta, blk_off[-1], grid):
            lv = 0
            while blk >= blk_off[lv + 1]:
                lv += 1
            npix = b * sizes[lv] * sizes[lv]
            pix0 = (blk - blk_off[lv]) * threads
            count = min(threads, npix - pix0)
            assert count >= 1
            seen[key_off[lv] + pix0: key_off[lv] + pix0 + count] += 1
    assert (seen == 1).all(), "every pixel of every level belongs to exactly one block"


def test_scale_bc_family_is_closed_under_differentiation(monkeypatch):
    """fused.ScaleBC / DotP / ScaleDot (the modulation / demodulation multiplies of ModulatedConv2d in the regulariser
    iterations) express each other's derivatives: first and second derivatives against finite differences, with the CUDA
    primitive (sr_scale_dot_nhwc_f32) replaced by its torch definition so that the autograd logic runs on the CPU."""
    import torch
    from torch.autograd import gradcheck, gradgradcheck
    from stylerenderer_b200 import fused

    def scale_dot(a, other, scale, round_out, want_out=True):
        b, _, _, c = a.shape
        out = (a * scale.view(b, 1, 1, c) if scale is not None else a.clone()) if want_out else None
        dot = (a * other).sum((1, 2)) if other is not None else None
        return out, dot
    monkeypatch.setattr(fused.tc, "scale_dot", scale_dot)
    torch.manual_seed(3)
    a = torch.randn(2, 3, 2, 4, dtype=torch.float64, requires_grad=True)
    o = torch.randn(2, 3, 2, 4, dtype=torch.float64, requires_grad=True)
    s = torch.randn(2, 4, dtype=torch.float64, requires_grad=True)
    assert gradcheck(fused.ScaleBC.apply, (a, s)) and gradgradcheck(fused.ScaleBC.apply, (a, s))
    assert gradcheck(fused.DotP.apply, (a, o)) and gradgradcheck(fused.DotP.apply, (a, o))
    assert gradcheck(fused.ScaleDot.apply, (a, o, s)) and gradgradcheck(fused.ScaleDot.apply, (a, o, s))
    # the shape of the path-length regulariser: a gradient-norm penalty through x -> ScaleBC -> nonlinearity
    x = torch.randn(2, 3, 2, 4, dtype=torch.float64, requires_grad=True)

    def penalty(fn):
        y = torch.tanh(fn(x, s)).sum()
        gx, = torch.autograd.grad(y, x, create_graph=True)
        return torch.autograd.grad(gx.pow(2).sum(), [x, s])
    got = penalty(fused.ScaleBC.apply)
    want = penalty(lambda xx, ss: xx * ss.view(2, 1, 1, 4))
    for g, w in zip(got, want):
        torch.testing.assert_close(g, w)
