"""CPU restatements of the INDEXING logic of two kernels (no GPU, no arithmetic parity claims -- those are the `-m gpu`
tests): there is no GPU in the authoring container, so loop structure / ring rotation / block-to-level look-ups are checked
here first ("keep a CPU restatement of each kernel's indexing logic for fast local checks").

* `fir_nhwc_stream_kernel` / `fir_planes_vec_kernel` (csrc/upfirdn2d.cu): work items, ring stages (TMA boxes with zero fill /
  16-byte aligned bulk-copy chunks with a 0..3-float lead), the compile-time ring of open output rows, the per-row shift
  selection and the interior / side column split; rank-1 taps use the separable form through the largest tap.  The models
  walk exactly the kernels' loops.
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


def rank1_taps(taps):
    """fir_rank1_taps (csrc/upfirdn2d.cu): flipped taps, factorisation through the first largest tap, separability test."""
    tk = np.array([[taps[K - 1 - a][K - 1 - b] for b in range(K)] for a in range(K)])
    pa, pb = np.unravel_index(np.argmax(np.abs(tk)), tk.shape)          # first largest tap, like the strict '>' scan
    kv, kh = tk[:, pb].copy(), tk[pa, :] / tk[pa, pb]
    best = np.abs(tk).max()
    sep = best > 0 and bool((np.abs(np.outer(kv, kh) - tk) <= 1e-6 * best).all())
    return tk, kv, kh, sep


def fir_stream_model(x, taps, pad, strip_w=32, rows=4):
    """One channel of fir_nhwc_stream_kernel<MODE 0>: work item = (column strip xt, row segment seg); the producer's TMA box
    = `rows` input rows x (strip_w + 3) columns with zero fill outside the plane; consumer = one output column; ring slot of
    the output row completed by step u of a stage is (u + 1) & 3."""
    ih, iw = x.shape
    oh, ow = ih + 2 * pad - K + 1, iw + 2 * pad - K + 1
    out = np.full((oh, ow), np.nan)
    writes = np.zeros((oh, ow), dtype=int)
    tk, kv, kh, sep = rank1_taps(taps)
    nseg = max(1, (oh + 32) // 64)
    seg_rows = (oh + nseg - 1) // nseg
    xtiles = (ow + strip_w - 1) // strip_w

    def box(cx, cy):                                                    # TMA box with out-of-bounds zero fill
        t = np.zeros((rows, strip_w + 3))
        for r in range(rows):
            for c in range(strip_w + 3):
                if 0 <= cy + r < ih and 0 <= cx + c < iw:
                    t[r, c] = x[cy + r, cx + c]
        return t
    for xt in range(xtiles):
        for seg in range(nseg):
            oy_start = seg * seg_rows
            oy_end = min(oh, oy_start + seg_rows)
            if oy_end <= oy_start:
                continue
            nstages = (oy_end - oy_start + 3 + rows - 1) // rows
            acc = np.zeros((strip_w, K))
            for k in range(nstages):
                stage = box(xt * strip_w - pad, oy_start - pad + rows * k)
                for u in range(rows):
                    for col in range(strip_w):
                        ox = xt * strip_w + col
                        cur = stage[u, col:col + K]
                        if sep:
                            h = sum(cur[bb] * kh[bb] for bb in range(K))
                            for a_ in range(K):
                                acc[col, (u - a_) & 3] += h * kv[a_]
                        else:
                            for a_ in range(K):
                                for bb in range(K):
                                    acc[col, (u - a_) & 3] += cur[bb] * tk[a_][bb]
                        oy = oy_start + k * rows + u - (K - 1)
                        if (k > 0 or u == K - 1) and oy < oy_end and ox < ow:
                            out[oy, ox] = acc[col, (u + 1) & 3]
                            writes[oy, ox] += 1
                        acc[col, (u + 1) & 3] = 0.0
    return out, writes, sep


def fir_planes_vec_model(planes, taps, pad0, pad1, rows=4):
    """fir_planes_vec_kernel on a stack of planes stored as ONE flat array (the reference layout [major, H, W]): a ring stage
    is a 16-byte aligned chunk (lead 0..3 floats before the first row), interior quads [e0, e1) read three aligned float4 and
    select by the row's shift a = base & 3, edge / tail columns are the scalar side job."""
    major, ih, iw = planes.shape
    flat = planes.reshape(-1)
    assert flat.size % 4 == 0
    oh, ow = ih + pad0 + pad1 - K + 1, iw + pad0 + pad1 - K + 1
    out = np.full((major, oh, ow), np.nan)
    writes = np.zeros((major, oh, ow), dtype=int)
    tk, kv, kh, sep = rank1_taps(taps)
    nq = ow // 4
    e0 = (pad0 + 3) // 4 if pad0 > 0 else 0
    e1 = (iw + pad0 - 7) // 4 + 1 if iw + pad0 - 7 >= 0 else 0
    e1 = min(e1, nq)
    e0 = min(e0, e1)
    side_cols = list(range(4 * e0)) + list(range(4 * e1, ow))
    assert len(side_cols) <= 32
    nseg = max(1, (oh + 64) // 128)
    seg_rows = (oh + nseg - 1) // nseg
    plane_in = ih * iw
    for m in range(major):
        lead_plane = (m * plane_in) & 3
        for seg in range(nseg):
            oy_start = seg * seg_rows
            oy_end = min(oh, oy_start + seg_rows)
            if oy_end <= oy_start:
                continue
            nstages = (oy_end - oy_start + 3 + rows - 1) // rows
            acc = np.zeros((ow, K))
            for k in range(nstages):
                iy0 = oy_start - pad0 + rows * k
                r0, r1 = max(iy0, 0), min(iy0 + rows, ih)
                stage = None
                if r1 > r0:                                             # producer: aligned chunk of the valid rows
                    idx = m * plane_in + r0 * iw
                    lead = idx & 3
                    nbytes = ((lead + (r1 - r0) * iw) * 4 + 15) & ~15
                    assert idx - lead >= 0 and idx - lead + nbytes // 4 <= flat.size, "the over-fetch stays inside the tensor"
                    stage = flat[idx - lead: idx - lead + nbytes // 4]
                    assert lead == (lead_plane + r0 * iw) & 3, "consumers recompute the producer's lead"
                base = ((lead_plane + r0 * iw) & 3) + (iy0 - r0) * iw - pad0
                for u in range(rows):
                    iy = iy0 + u
                    if 0 <= iy < ih:
                        a = base & 3
                        for t in range(e0, e1):                         # interior quads: three aligned float4, no masking
                            q0 = (base - a) + 4 * t
                            assert q0 >= 0 and q0 % 4 == 0 and q0 + 12 <= stage.size + 4, (q0, stage.size)
                            f = np.zeros(12)
                            n = min(12, stage.size - q0)
                            f[:n] = stage[q0:q0 + n]
                            c = f[a:a + 7]
                            for j in range(4):
                                ox = 4 * t + j
                                vals = c[j:j + K]
                                if sep:
                                    h = sum(vals[bb] * kh[bb] for bb in range(K))
                                    for a_ in range(K):
                                        acc[ox, (u - a_) & 3] += h * kv[a_]
                                else:
                                    for a_ in range(K):
                                        for bb in range(K):
                                            acc[ox, (u - a_) & 3] += vals[bb] * tk[a_][bb]
                        for ox in side_cols:                            # side job: predicated scalar loads
                            six = ox - pad0
                            vals = [stage[base + ox + bb] if 0 <= six + bb < iw else 0.0 for bb in range(K)]
                            if sep:
                                h = sum(vals[bb] * kh[bb] for bb in range(K))
                                for a_ in range(K):
                                    acc[ox, (u - a_) & 3] += h * kv[a_]
                            else:
                                for a_ in range(K):
                                    for bb in range(K):
                                        acc[ox, (u - a_) & 3] += vals[bb] * tk[a_][bb]
                    base += iw
                    oy = oy_start + k * rows + u - 3
                    if (k > 0 or u == 3) and oy < oy_end:
                        out[m, oy, :] = acc[:, (u + 1) & 3]
                        writes[m, oy, :] += 1
                    acc[:, (u + 1) & 3] = 0.0
    return out, writes, sep


def _tap_cases(rng):
    k1 = np.array([1., 3., 3., 1.])
    return [(np.outer(k1, k1) / 64 * 4, True),                          # the model's FIR (reference layers.py:7-12)
            (np.outer([1., 2., 3., 4.], [4., 3., -2., 1.]), True),      # asymmetric rank-1: flips and factor order matter
            (rng.standard_normal((K, K)), False)]                       # general taps: 2-D form


@pytest.mark.parametrize("geom", [(33, 33, 1), (32, 32, 2), (70, 41, 1), (129, 36, 2), (40, 65, 1), (34, 100, 2)])
def test_fir_stream_kernel_indexing(geom):
    """Work items, stage boxes and the compile-time ring of fir_nhwc_stream_kernel: every output written exactly once, equal
    to the FIR, for ragged column strips (41, 65, 100 wide), several row segments (129 rows) and both pads of the model."""
    ih, iw, pad = geom
    rng = np.random.default_rng(ih * 100 + iw)
    x = rng.standard_normal((ih, iw))
    for taps, expect_sep in _tap_cases(rng):
        got, writes, sep = fir_stream_model(x, taps, pad)
        assert sep == expect_sep
        assert (writes == 1).all(), "every output pixel is written exactly once"
        np.testing.assert_allclose(got, fir_reference(x, taps, pad), rtol=0, atol=1e-12)


@pytest.mark.parametrize("geom", [(4, 17, 257, 1, 1), (4, 16, 256, 2, 2), (2, 150, 70, 2, 1), (4, 20, 65, 1, 1), (8, 18, 129, 3, 3),
                                  (4, 16, 96, -1, 2), (4, 33, 64, 0, 3)])
def test_fir_planes_vec_kernel_indexing(geom):
    """Chunk alignment (lead), per-row shift selection, interior / side column split and the ring of fir_planes_vec_kernel on
    unaligned 257- / 65- / 129-wide planes, asymmetric and negative pads, several segments: every output written once, equal to
    the FIR, and no read before the chunk or past the tensor."""
    major, ih, iw, pad0, pad1 = geom
    rng = np.random.default_rng(ih * 1000 + iw)
    planes = rng.standard_normal((major, ih, iw))
    for taps, expect_sep in _tap_cases(rng):
        got, writes, sep = fir_planes_vec_model(planes, taps, pad0, pad1)
        assert sep == expect_sep
        assert (writes == 1).all()
        for m in range(major):
            xp = np.pad(planes[m], ((max(pad0, 0), max(pad1, 0)), (max(pad0, 0), max(pad1, 0))))
            xp = xp[max(-pad0, 0): xp.shape[0] - max(-pad1, 0), max(-pad0, 0): xp.shape[1] - max(-pad1, 0)]
            kf = taps[::-1, ::-1]
            oh, ow = xp.shape[0] - K + 1, xp.shape[1] - K + 1
            want = np.array([[(xp[y:y + K, c:c + K] * kf).sum() for c in range(ow)] for y in range(oh)])
            np.testing.assert_allclose(got[m], want, rtol=0, atol=1e-12)


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
