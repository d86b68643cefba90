"""GPU parity tests: the sm_100a kernels (through the C ABI, via stylerenderer_b200.op) against the oracle
(oracle/sr_oracle.c, oracle/torch_ref.py), the committed golden fixtures and size-independent properties.

Bars (BASELINE.json north_star): bit-exact for the rasteriser's integer index buffer (and, here, also its
coefficients and dcoeff), bit-exact for fused_bias_act (same op order as the reference kernel), <= 1e-3 rel
for everything floating-point (tolerances are written at each assert and are far tighter)."""
import math

import pytest
import torch

from make_golden import grid_mesh, seeded
from oracle import cpu as O
from oracle import torch_ref as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def op():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from stylerenderer_b200 import op as _op
    return _op


def cuda(t):
    return t.cuda() if t is not None else None


# ----------------------------------------------------------------------------------- fused_bias_act
FBA_SHAPES = [(2, 5, 7, 9), (4, 16), (2, 8, 4, 4), (3, 12, 8, 8), (32, 512), (1, 3, 1, 1), (2, 6, 5, 4), (5, 7)]


@pytest.mark.parametrize("shape", FBA_SHAPES)
def test_fused_bias_act_all_modes_bit_exact(op, shape):
    x, b, ref = seeded(shape, 1), seeded((shape[1],), 2), seeded(shape, 3)
    empty = x.new_empty(0)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0), (1, 1), (1, 2)]:
        for bb, rr in [(b, empty), (empty, ref), (b, ref), (empty, empty)]:
            want = O.fused_bias_act(x, bb, rr, act, grad, 0.2, 2 ** 0.5)
            got = op.fused_bias_act(cuda(x), cuda(bb), cuda(rr), act, grad, 0.2, 2 ** 0.5).cpu()
            assert torch.equal(got, want), (shape, act, grad)


def test_fused_bias_act_channels_last_and_noncontiguous(op):
    x, b = seeded((3, 8, 6, 5), 4), seeded((8,), 5)
    want = O.fused_leaky_relu(x, b, 0.1, 1.5)
    xcl = cuda(x).contiguous(memory_format=torch.channels_last)
    got = op.fused_leaky_relu(xcl, cuda(b), 0.1, 1.5)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got.cpu(), want)
    xs = cuda(seeded((3, 8, 6, 10), 6))[..., ::2]                # strided view -> .contiguous() inside
    assert torch.equal(op.fused_leaky_relu(xs, cuda(b)).cpu(), O.fused_leaky_relu(xs.cpu(), b))


def test_fused_leaky_relu_golden(op, golden):
    for name, g in golden["fused_leaky_relu"].items():
        got = op.fused_leaky_relu(cuda(g["x"]), cuda(g["b"])).cpu()
        torch.testing.assert_close(got, g["y"], rtol=1e-6, atol=1e-7, msg=name)


@pytest.mark.parametrize("shape", FBA_SHAPES + [(4, 16, 32, 32), (2, 12, 9, 9)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_leaky_relu_backward(op, shape, channels_last):
    if channels_last and len(shape) != 4:
        pytest.skip("channels_last needs 4-D")
    x, b, gy = seeded(shape, 7), seeded((shape[1],), 8), seeded(shape, 9)
    y = O.fused_leaky_relu(x, b)
    gx_want, gb_want = O.fused_leaky_relu_backward(gy, y)
    xc = cuda(x).requires_grad_(True)
    bc = cuda(b).requires_grad_(True)
    xin = xc.contiguous(memory_format=torch.channels_last) if channels_last else xc
    out = op.fused_leaky_relu(xin, bc)
    gx, gb = torch.autograd.grad(out, (xc, bc), cuda(gy))
    assert torch.equal(gx.cpu(), gx_want)                           # dx is elementwise: bit exact
    torch.testing.assert_close(gb.cpu().double(), gb_want, rtol=1e-4, atol=1e-4)   # fp32 sum of up to 4096 terms


def test_fused_leaky_relu_double_backward(op):
    """R1 / path-length regularisers differentiate through the backward (reference op/fused_act.py:43-49)."""
    x, b = seeded((2, 6, 5, 5), 10), seeded((6,), 11)
    def run(f, xx, bb):
        xx = xx.clone().requires_grad_(True); bb = bb.clone().requires_grad_(True)
        y = f(xx, bb)
        gx, gb = torch.autograd.grad(y, (xx, bb), torch.ones_like(y) * 0.5 + y.detach() * 0, create_graph=True)
        loss = (gx * gx).sum() + (gb * gb).sum() + (y * y).sum()
        return torch.autograd.grad(loss, (xx, bb))
    ref = run(lambda xx, bb: torch.nn.functional.leaky_relu(xx + bb.view(1, -1, 1, 1), 0.2) * 2 ** 0.5, x, b)
    got = run(lambda xx, bb: op.fused_leaky_relu(xx, bb), cuda(x), cuda(b))
    for r, g in zip(ref, got):
        torch.testing.assert_close(g.cpu(), r, rtol=1e-5, atol=1e-5)


def test_fused_bias_act_full_size_property(op):
    """Config-2 size (largest activation of G(256) at B=8): bit-equal to the same expression in stock torch ops."""
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(8, 128, 256, 256, device="cuda", generator=g)
    b = torch.randn(128, device="cuda", generator=g)
    y = op.fused_leaky_relu(x, b)
    want = torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert torch.equal(y, want)
    gy = torch.randn_like(x)
    from stylerenderer_b200.op.fused_act import _lrelu_backward
    dx, db = _lrelu_backward(gy, y, 0.2, 2 ** 0.5, True)
    dx_want = torch.where(y > 0, gy, gy * 0.2) * 2 ** 0.5
    assert torch.equal(dx, dx_want)
    torch.testing.assert_close(db, dx_want.sum((0, 2, 3)), rtol=1e-4, atol=1e-2)


# ----------------------------------------------------------------------------------- upfirdn2d
def test_upfirdn2d_golden(op, golden):
    for name, g in golden["upfirdn2d"].items():
        y = op.upfirdn2d(cuda(g["x"]), cuda(g["k"]), up=g["up"], down=g["down"], pad=g["pad"]).cpu()
        assert y.shape == g["y"].shape, name
        torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=1e-6, msg=name)


UPFIR_CASES = [
    # (n, c, h, w, kh, kw, up, down, pad0, pad1)
    (2, 3, 9, 9, 4, 4, 1, 1, 1, 1), (2, 3, 17, 17, 4, 4, 1, 1, 1, 1), (1, 2, 33, 65, 4, 4, 1, 1, 2, 2),
    (3, 5, 4, 4, 4, 4, 1, 1, 2, 1), (2, 300, 5, 5, 4, 4, 1, 1, 1, 1), (1, 1, 129, 70, 4, 4, 1, 1, 2, 1),
    (2, 3, 4, 4, 4, 4, 2, 1, 2, 1), (2, 3, 16, 16, 4, 4, 2, 1, 2, 1), (1, 2, 31, 45, 4, 4, 2, 1, 2, 1),
    (1, 2, 8, 8, 4, 4, 2, 1, 1, 2), (1, 2, 9, 7, 4, 4, 2, 1, 3, 0), (1, 3, 64, 64, 4, 4, 2, 1, 2, 1),
    (2, 3, 8, 8, 4, 4, 1, 2, 1, 1), (2, 3, 32, 32, 4, 4, 1, 2, 1, 1), (1, 2, 33, 47, 4, 4, 1, 2, 2, 2),
    (1, 2, 128, 128, 4, 4, 1, 2, 1, 1), (1, 2, 16, 16, 4, 4, 1, 2, 0, 3),
    (1, 2, 8, 8, 3, 3, 1, 1, 1, 1), (1, 2, 8, 9, 2, 2, 2, 1, 1, 0), (1, 2, 10, 10, 4, 4, 1, 1, -1, 2),
    (1, 2, 8, 8, 4, 4, 2, 2, 1, 2), (1, 2, 12, 12, 6, 6, 3, 2, 3, 2), (1, 1, 7, 5, 5, 3, 1, 1, 2, 2),
    # wide rows: the direct-from-global blur kernel (aligned quads + phase select), general (non-separable) taps
    (1, 2, 40, 300, 4, 4, 1, 1, 1, 1), (2, 1, 19, 131, 4, 4, 1, 1, 2, 2), (1, 1, 6, 1030, 4, 4, 1, 1, 1, 1),
    (3, 2, 33, 257, 4, 4, 1, 1, 3, 3), (1, 1, 70, 129, 4, 4, 1, 1, 0, 0),
]


@pytest.mark.parametrize("case", UPFIR_CASES)
def test_upfirdn2d_vs_oracle(op, case):
    n, c, h, w, kh, kw, up, down, p0, p1 = case
    x = seeded((n, c, h, w), 20)
    k = seeded((kh, kw), 21)
    want = O.upfirdn2d(x, k, up, down, (p0, p1))
    got = op.upfirdn2d(cuda(x), cuda(k), up=up, down=down, pad=(p0, p1)).cpu()
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape,pad", [((2, 3, 9, 9), (1, 1)), ((3, 5, 65, 65), (1, 1)), ((1, 2, 257, 257), (1, 1)),
                                       ((2, 2, 64, 100), (2, 2)), ((1, 3, 31, 200), (2, 1)), ((40, 7, 5, 5), (1, 1))])
def test_upfirdn2d_separable_taps(op, shape, pad):
    """Rank-1 FIRs (the only kind the model builds: make_kernel, reference layers.py:7-12) take the separable strip path
    of the NCHW tile kernel; asymmetric factors so that a wrong flip or a row/column mix-up cannot cancel."""
    x = seeded(shape, 33)
    for kv, kh in ((torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])), (seeded((4,), 34), seeded((4,), 35))):
        k = (kv[:, None] * kh[None, :])
        k = k / k.abs().sum()
        want = O.upfirdn2d(x, k, 1, 1, pad)
        got = op.upfirdn2d(cuda(x), cuda(k), pad=pad).cpu()
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kernel", ["default", "vec", "scalar"])
@pytest.mark.parametrize("rank1", [True, False])
@pytest.mark.parametrize("shape,pad", [((4, 1, 257, 257), (1, 1)), ((2, 2, 256, 256), (2, 2)), ((1, 4, 100, 513), (1, 1)),
                                       ((2, 2, 130, 70), (2, 1)), ((1, 4, 64, 96), (-1, 2)), ((4, 1, 300, 64), (0, 3)),
                                       ((1, 8, 17, 385), (3, 3)), ((2, 2, 16, 66), (1, 1))])
def test_upfirdn2d_planes_stream_kernel(op, shape, pad, rank1, kernel, monkeypatch):
    """The row-streaming bulk-copy blurs of the reference layout ([N*C, H, W] planes, reference op/upfirdn2d.py:99; planes
    >= 16 rows x 64..512 columns: fir_planes_vec_kernel = a thread owns 4 adjacent columns, aligned LDS.128 + selection by
    the row's misalignment, scalar side job for the edge columns; fir_planes_stream_kernel = a lane owns columns 32 apart)
    against the oracle: unaligned 257-wide rows (every chunk lead and row shift 0..3), several row segments, ragged
    widths, asymmetric and negative pads, separable and general taps; "default" = the measured per-shape choice; and
    near-equal values against the tile kernels they replace (SR_FIR_PLANES_STREAM=0)."""
    if kernel == "vec":
        monkeypatch.setenv("SR_FIR_PLANES_VEC", "1")
    elif kernel == "scalar":
        monkeypatch.setenv("SR_FIR_PLANES_VEC", "0")
        monkeypatch.setenv("SR_FIR_PLANES_J", "2")
    x = seeded(shape, 40)
    if rank1:
        k = torch.outer(torch.tensor([1., 3., 3., 1.]), seeded((4,), 41))
        k = k / k.abs().sum()
    else:
        k = seeded((4, 4), 42)
    want = O.upfirdn2d(x, k, 1, 1, pad)
    got = op.upfirdn2d(cuda(x), cuda(k), pad=pad)
    assert got.shape == want.shape
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)
    monkeypatch.setenv("SR_FIR_PLANES_STREAM", "0")
    old = op.upfirdn2d(cuda(x), cuda(k), pad=pad)
    torch.testing.assert_close(got, old, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape,pad", [((2, 8, 9, 9), (1, 1)), ((1, 128, 33, 33), (1, 1)), ((3, 12, 16, 16), (2, 2)),
                                       ((2, 64, 5, 7), (2, 1)), ((1, 4, 64, 64), (2, 2)),
                                       # TMA-staged tile kernel (C % 32 == 0, output >= 32 x 32), ragged edges and pads
                                       ((2, 64, 70, 45), (1, 1)), ((1, 32, 40, 100), (2, 2)), ((3, 96, 33, 47), (2, 1)),
                                       ((1, 256, 65, 65), (1, 1))])
def test_upfirdn2d_channels_last(op, shape, pad):
    """channels_last tensors take the NHWC kernel (no layout copy) and keep their memory format, fwd and bwd."""
    x = seeded(shape, 30)
    k = seeded((4, 4), 31)
    want = O.upfirdn2d(x, k, 1, 1, pad)
    xc = cuda(x).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = op.upfirdn2d(xc, cuda(k), pad=pad)
    assert got.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(got.detach().cpu(), want, rtol=1e-5, atol=1e-5)
    gy = seeded(want.shape, 32)
    gx, = torch.autograd.grad(got, xc, cuda(gy).contiguous(memory_format=torch.channels_last))
    xr = x.clone().requires_grad_(True)
    gx_want, = torch.autograd.grad(T.upfirdn2d(xr, k, 1, 1, pad), xr, gy)
    torch.testing.assert_close(gx.cpu(), gx_want, rtol=1e-5, atol=1e-5)


def test_upfirdn2d_minor_and_mixed_factors(op):
    """The pybind-level entry with minor > 1 and different x / y factors (reference op/upfirdn2d.cpp:24-26)."""
    x = seeded((3, 6, 7, 4), 22)
    k = seeded((4, 3), 23)
    want = O.upfirdn2d_raw(x, k, 2, 1, 1, 2, 1, 2, 2, 0)
    got = op.upfirdn2d_raw(cuda(x), cuda(k), 2, 1, 1, 2, 1, 2, 2, 0).cpu()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("up,down,pad", [(1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (2, 2))])
def test_upfirdn2d_gradients(op, up, down, pad):
    """First and second derivative against autograd through the native-PyTorch restatement."""
    x = seeded((2, 3, 10, 10), 24)
    k = T.fir_taps([1, 3, 3, 1]) * (up ** 2)
    def run(f, xx, kk):
        xx = xx.clone().requires_grad_(True)
        y = f(xx, kk)
        gy = ((torch.arange(y.numel(), dtype=torch.float32, device=y.device).view_as(y) % 7 - 3) / 3).requires_grad_(True)
        gx, = torch.autograd.grad(y, xx, gy, create_graph=True)
        # the op is linear in x, so the second-order path runs through the cotangent (as in R1 / path regularisers)
        ggy, = torch.autograd.grad((gx * gx).sum(), gy)
        return y.detach(), gx.detach(), ggy
    ref = run(lambda a, b: T.upfirdn2d(a, b, up, down, pad), x, k)
    got = run(lambda a, b: op.upfirdn2d(a, b, up=up, down=down, pad=pad), cuda(x), cuda(k))
    for r, g in zip(ref, got):
        torch.testing.assert_close(g.cpu(), r, rtol=1e-5, atol=1e-5)


def test_upfirdn2d_full_size_properties(op):
    """Largest blur of G(256) at B=8 (planes 8*128, 257^2 -> 256^2): vs a depthwise cuDNN conv, plus linearity."""
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(8, 128, 257, 257, device="cuda", generator=g)
    k = (T.fir_taps([1, 3, 3, 1]) * 4).cuda()
    y = op.upfirdn2d(x, k, pad=(1, 1))
    assert y.shape == (8, 128, 256, 256)
    torch.backends.cudnn.allow_tf32 = False              # the comparison conv must be true fp32
    want = torch.nn.functional.conv2d(x.view(-1, 1, 257, 257), torch.flip(k, [0, 1]).view(1, 1, 4, 4), padding=1)
    want = want.view(8, 128, 256, 256)
    torch.testing.assert_close(y, want, rtol=1e-4, atol=1e-4)
    x2 = torch.randn_like(x)
    lin = op.upfirdn2d(2 * x + x2, k, pad=(1, 1))
    torch.testing.assert_close(lin, 2 * y + op.upfirdn2d(x2, k, pad=(1, 1)), rtol=1e-4, atol=1e-4)
    # adjointness <A x, g> == <x, A^T g> (the backward operator is the transpose)
    xs = x[:1].clone().requires_grad_(True)
    ys = op.upfirdn2d(xs, k, pad=(1, 1))
    gy = torch.randn_like(ys)
    gx, = torch.autograd.grad(ys, xs, gy)
    lhs, rhs = (ys.detach() * gy).sum().double(), (xs.detach() * gx).sum().double()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)


# ----------------------------------------------------------------------------------- rasterize
def _raster_both(op, v, tri, h, perspective=False, eps=1e-6):
    ind_w, coeff_w, _ = O.rasterize_forward(v, tri, h, 0, perspective, eps)
    ind_g, coeff_g = op.rasterize_forward(cuda(v), cuda(tri), h, 0, perspective, eps)
    return ind_w, coeff_w, ind_g.cpu(), coeff_g.cpu()


def test_rasterize_known_answer(op, golden):
    g = golden["rasterize"]["selftest"]                              # reference op/rasterize.py:83-107
    for dtype in (torch.float64, torch.float32):
        out = op.rasterize(cuda(g["v"].to(dtype)), cuda(g["t"].to(dtype)), cuda(g["f"]), 5).cpu()
        torch.testing.assert_close(out.double(), g["out"], rtol=0, atol=1e-12 if dtype == torch.float64 else 1e-6)


@pytest.mark.parametrize("name", ["grid24_h32_f32", "grid24_h8_f32", "grid16_h16_f64"])
def test_rasterize_golden(op, golden, name):
    g = golden["rasterize"][name]
    v, tri = grid_mesh(g["n"], g["b"], g["seed"], dtype=g["tex"].dtype)
    vc, tc = cuda(v).requires_grad_(True), cuda(g["tex"]).requires_grad_(True)
    out, ind, coeff = op.rasterize(vc, tc, cuda(tri), g["h"], return_buffers=True)
    assert torch.equal(ind.cpu(), g["ind"].long())
    assert torch.equal(coeff.cpu(), g["coeff"])
    f32 = v.dtype == torch.float32
    torch.testing.assert_close(out.detach().cpu(), g["out"], rtol=1e-5 if f32 else 1e-12, atol=1e-6 if f32 else 1e-13)
    gv, gt = torch.autograd.grad(out, (vc, tc), cuda(g["go"]))
    torch.testing.assert_close(gv.cpu(), g["gv"], rtol=1e-3, atol=1e-4 * float(g["gv"].abs().max()))
    torch.testing.assert_close(gt.cpu(), g["gt"], rtol=1e-3, atol=1e-4 * float(g["gt"].abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("h", [1, 4, 16, 64, 100, 256])
@pytest.mark.parametrize("perspective", [False, True])
def test_rasterize_bit_exact_vs_oracle(op, dtype, h, perspective):
    v, tri = grid_mesh(40, 3, 1000 + h, dtype=dtype)
    if perspective:
        v[..., 2] -= 3
    ind_w, coeff_w, ind_g, coeff_g = _raster_both(op, v, tri, h, perspective)
    assert torch.equal(ind_g, ind_w)
    assert torch.equal(coeff_g, coeff_w)
    dc_w = O.rasterize_backward(v, ind_w, perspective, 1e-6)
    dc_g = op.rasterize_backward(cuda(v), cuda(ind_w), perspective, 1e-6).cpu()
    assert torch.equal(dc_g, dc_w)


def test_rasterize_edge_cases(op):
    g = torch.Generator().manual_seed(5)
    # soup: zero-area triangles, repeated vertices, out-of-range ids, off-screen and screen-filling triangles
    v = torch.rand(2, 50, 3, generator=g) * 2.6 - 1.3
    v[:, 10] = v[:, 11]
    v[:, 12, :2] = v[:, 13, :2]
    tri = torch.randint(0, 50, (400, 3), generator=g)
    tri[5] = torch.tensor([10, 11, 20]); tri[6] = torch.tensor([7, 7, 7]); tri[7] = torch.tensor([0, 60, 1])
    tri[8] = torch.tensor([-1, 2, 3]); tri[9] = torch.tensor([12, 13, 12])
    for h in (1, 7, 33, 128):                                      # large boxes -> the warp-cooperative path
        ind_w, coeff_w, ind_g, coeff_g = _raster_both(op, v, tri, h)
        assert torch.equal(ind_g, ind_w) and torch.equal(coeff_g, coeff_w), h
    # equal depths: duplicated triangles -> the first one in list order must win (strict z test)
    vq = torch.tensor([[[-1., -1, 0], [-1, 1, 0], [1, 1, 0], [1, -1, 0], [-1, -1, 0], [-1, 1, 0], [1, 1, 0]]])
    tq = torch.tensor([[6, 5, 4], [2, 1, 0], [3, 2, 0], [2, 1, 0]])
    ind_w, coeff_w, ind_g, coeff_g = _raster_both(op, vq, tq, 16)
    assert torch.equal(ind_g, ind_w) and torch.equal(coeff_g, coeff_w)
    assert int((ind_w[..., 0] == 6).sum()) > 0
    # per-batch triangle lists and the unbatched [n,3]/[f,3] form
    trib = torch.stack([tri, tri.flip(0)])
    ind_w, coeff_w, _ = O.rasterize_forward(v, trib, 16)
    ind_g, coeff_g = op.rasterize_forward(cuda(v), cuda(trib), 16)
    assert torch.equal(ind_g.cpu(), ind_w) and torch.equal(coeff_g.cpu(), coeff_w)
    ind_w, coeff_w, _ = O.rasterize_forward(v[0], tri, 16)
    ind_g, coeff_g = op.rasterize_forward(cuda(v[0]), cuda(tri), 16)
    assert ind_g.shape == ind_w.shape and torch.equal(ind_g.cpu(), ind_w) and torch.equal(coeff_g.cpu(), coeff_w)
    # empty mesh -> background everywhere
    ind_g, coeff_g = op.rasterize_forward(cuda(v), torch.zeros(0, 3, dtype=torch.int64, device="cuda"), 8)
    assert int(ind_g.abs().sum()) == 0 and float(coeff_g.abs().sum()) == 0
    with pytest.raises(RuntimeError):
        op.rasterize_forward(cuda(v), cuda(tri), 8, 16)            # non-square: refused (reference quirk 6)
    with pytest.raises(RuntimeError):
        op.rasterize_forward(v, tri, 8)                            # CPU tensors: no fallback


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("scalar_tex", [False, True])
def test_rasterize_autograd_vs_oracle(op, dtype, scalar_tex):
    v, tri = grid_mesh(30, 2, 77, dtype=dtype)
    tex = seeded((2, 900) if scalar_tex else (2, 900, 3), 78, dtype)
    out_w, ind_w, coeff_w = O.rasterize(v, tex, tri, 48)
    go = seeded(out_w.shape, 79, dtype)
    gv_w, gt_w = O.rasterize_grads(v, tex, ind_w, coeff_w, go)
    vc, tc = cuda(v).requires_grad_(True), cuda(tex).requires_grad_(True)
    out = op.rasterize(vc, tc, cuda(tri), 48)
    f32 = dtype == torch.float32
    torch.testing.assert_close(out.detach().cpu(), out_w, rtol=1e-5 if f32 else 1e-12, atol=1e-6 if f32 else 1e-13)
    gv, gt = torch.autograd.grad(out, (vc, tc), cuda(go))
    tol = dict(rtol=1e-3, atol=1e-4 * float(gv_w.abs().max())) if f32 else dict(rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(gv.cpu(), gv_w, **tol)
    tol = dict(rtol=1e-3, atol=1e-4 * float(gt_w.abs().max())) if f32 else dict(rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(gt.cpu(), gt_w, **tol)


def test_rasterize_gradcheck_f64(op):
    """The reference's own self-test ends with gradcheck in float64 (reference op/rasterize.py:104-107)."""
    v = torch.tensor([[[-1, -1, 0], [-1, 1, 0], [1, 0, 0]]], dtype=torch.float64, device="cuda")
    t = torch.tensor([[[1, 0], [0, 1], [0, 0]]], dtype=torch.float64, device="cuda")
    f = torch.tensor([[2, 1, 0]], device="cuda")
    x = torch.cat((v, t), -1).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda x_: op.rasterize(x_[:, :, :3], x_[:, :, 3:], f, 5), x, eps=1e-6, atol=1e-6,
                                    nondet_tol=1e-9)


def test_rasterize_config3_full_size(op):
    """BASELINE.json configs[2]: BFM-size mesh (35 721 verts / 70 688 tris) -> 256x256, bit-exact on a 4-image slice
    against the oracle, and internal consistency of the whole batch of 64."""
    v, tri = grid_mesh(189, 64, 4242, jitter=0.002)
    tex = torch.nn.functional.normalize(seeded((64, 189 * 189, 3), 4243), dim=-1)
    vc, tc, fc = cuda(v), cuda(tex), cuda(tri)
    out, ind, coeff = op.rasterize(vc, tc, fc, 256, return_buffers=True)
    ind_w, coeff_w, _ = O.rasterize_forward(v[:4], tri, 256, 0, False, 1e-6)
    assert torch.equal(ind[:4].cpu(), ind_w) and torch.equal(coeff[:4].cpu(), coeff_w)
    covered = (coeff.sum(-1) > 0)
    assert 0.3 < float(covered.float().mean()) < 1.0
    s = coeff.sum(-1)[covered]
    assert float((s - 1).abs().max()) < 1e-5                       # barycentrics sum to one
    assert bool((ind[covered] // (189 * 189) == torch.arange(64, device="cuda").view(64, 1, 1).expand(64, 256, 256)[covered].unsqueeze(-1)).all())
    # depth check: z = sum coeff * v_z is the max over all candidate triangles -> re-rendering is idempotent
    out2, ind2, coeff2 = op.rasterize(vc, tc, fc, 256, return_buffers=True)
    assert torch.equal(ind, ind2) and torch.equal(coeff, coeff2) and torch.equal(out, out2)


# ----------------------------------------------------------------------------------- rasterize pyramid
PYRAMID_SIZES = [4, 8, 16, 32, 64, 128, 256]


@pytest.mark.parametrize("perspective", [False, True])
def test_rasterize_pyramid_bit_exact_vs_oracle_and_single_size(op, perspective):
    """sr_rasterize_pyramid_forward_f32: every level of the GeneratorWithMap pyramid (reference model.py:260-270) must be
    bit-identical to the oracle's rasterize_cpu restatement at that size -- ids, coefficients -- and to the single-size
    kernel including the interpolated attributes."""
    from stylerenderer_b200.op.rasterize import RasterizePyramid
    v, tri = grid_mesh(40, 3, 3100)
    if perspective:
        v[..., 2] -= 3
    tex = seeded((3, 1600, 3), 3101)
    vc, tc, fc = cuda(v), cuda(tex), cuda(tri)
    outs = op.rasterize_pyramid(vc, tc, fc, PYRAMID_SIZES, perspective)
    assert [tuple(o.shape) for o in outs] == [(3, s, s, 3) for s in PYRAMID_SIZES]
    for s, o in zip(PYRAMID_SIZES, outs):
        out1, ind1, coeff1 = op.rasterize(vc, tc, fc, s, 0, perspective, return_buffers=True)
        assert torch.equal(o, out1), s
        ind_w, coeff_w, _ = O.rasterize_forward(v, tri, s, 0, perspective, 1e-6)
        assert torch.equal(ind1.cpu(), ind_w) and torch.equal(coeff1.cpu(), coeff_w), s
    # the buffers the pyramid itself wrote (saved for its backward), not only the interpolated maps
    vr = vc.clone().requires_grad_(True)
    outs = RasterizePyramid.apply(vr, tc, fc, tuple(PYRAMID_SIZES), perspective, 1e-6)
    saved = outs[0].grad_fn.saved_tensors
    inds, coeffs = saved[2:2 + len(PYRAMID_SIZES)], saved[2 + len(PYRAMID_SIZES):]
    for s, ind, coeff in zip(PYRAMID_SIZES, inds, coeffs):
        ind_w, coeff_w, _ = O.rasterize_forward(v, tri, s, 0, perspective, 1e-6)
        assert torch.equal(ind.cpu(), ind_w) and torch.equal(coeff.cpu(), coeff_w), s


def test_rasterize_pyramid_edge_cases(op):
    g = torch.Generator().manual_seed(6)
    v = torch.rand(2, 50, 3, generator=g) * 2.6 - 1.3                # soup with big, degenerate and invalid triangles
    v[:, 10] = v[:, 11]
    tri = torch.randint(0, 50, (300, 3), generator=g)
    tri[5] = torch.tensor([10, 11, 20]); tri[6] = torch.tensor([7, 7, 7]); tri[7] = torch.tensor([0, 60, 1])
    tex = seeded((2, 50), 3)                                          # scalar attribute [b,n]
    sizes = [1, 7, 33, 128]                                           # any square sizes, not only powers of two
    outs = op.rasterize_pyramid(cuda(v), cuda(tex), cuda(tri), sizes)
    for s, o in zip(sizes, outs):
        assert o.shape == (2, s, s)
        assert torch.equal(o, op.rasterize(cuda(v), cuda(tex), cuda(tri), s)), s
    # per-batch triangle lists, the unbatched form, an empty mesh, a single level
    trib = torch.stack([tri, tri.flip(0)])
    for vv, ff, tt in [(v, trib, tex), (v[0], tri, tex[0])]:
        outs = op.rasterize_pyramid(cuda(vv), cuda(tt), cuda(ff), [8, 16])
        for s, o in zip([8, 16], outs):
            assert torch.equal(o, op.rasterize(cuda(vv), cuda(tt), cuda(ff), s))
    empty = op.rasterize_pyramid(cuda(v), cuda(tex), torch.zeros(0, 3, dtype=torch.int64, device="cuda"), [4, 8])
    assert all(float(o.abs().sum()) == 0 for o in empty)
    one, = op.rasterize_pyramid(cuda(v), cuda(tex), cuda(tri), [16])
    assert torch.equal(one, op.rasterize(cuda(v), cuda(tex), cuda(tri), 16))
    with pytest.raises(RuntimeError):
        op.rasterize_pyramid(cuda(v), cuda(tex), cuda(tri), list(range(1, 10)))      # more than SR_RASTER_MAX_LEVELS
    with pytest.raises(RuntimeError):
        op.rasterize_pyramid(v, tex, tri, [8])                                       # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        op.rasterize_pyramid(cuda(v).double(), cuda(tex).double(), cuda(tri), [8])   # float32 only


def test_rasterize_pyramid_backward_is_the_sum_of_the_levels(op):
    """One scatter launch for all levels = the sum of the per-size backward passes (oracle: O.rasterize_grads per level);
    levels without a gradient take no part."""
    v, tri = grid_mesh(30, 2, 87)
    tex = seeded((2, 900, 3), 88)
    sizes = [4, 16, 48, 64]
    gos = [seeded((2, s, s, 3), 90 + i) for i, s in enumerate(sizes)]
    gv_w, gt_w = torch.zeros_like(v), torch.zeros_like(tex)
    for s, go in zip(sizes, gos):
        if s == 16:
            continue                                                   # this level's output is unused below
        out_w, ind_w, coeff_w = O.rasterize(v, tex, tri, s)
        a, b_ = O.rasterize_grads(v, tex, ind_w, coeff_w, go)
        gv_w += a
        gt_w += b_
    vc, tc = cuda(v).requires_grad_(True), cuda(tex).requires_grad_(True)
    outs = op.rasterize_pyramid(vc, tc, cuda(tri), sizes)
    used = [(o, go) for s, o, go in zip(sizes, outs, gos) if s != 16]
    gv, gt = torch.autograd.grad([o for o, _ in used], (vc, tc), [cuda(go) for _, go in used])
    torch.testing.assert_close(gv.cpu(), gv_w, rtol=1e-3, atol=1e-4 * float(gv_w.abs().max()))
    torch.testing.assert_close(gt.cpu(), gt_w, rtol=1e-3, atol=1e-4 * float(gt_w.abs().max()))
    # only the vertex gradient / only the attribute gradient
    outs = op.rasterize_pyramid(vc, tc.detach(), cuda(tri), sizes)
    gv2, = torch.autograd.grad(outs, (vc,), [cuda(go) for go in gos])
    outs = op.rasterize_pyramid(vc.detach(), tc, cuda(tri), sizes)
    gt2, = torch.autograd.grad(outs, (tc,), [cuda(go) for go in gos])
    assert gv2.shape == v.shape and gt2.shape == tex.shape and float(gv2.abs().max()) > 0 and float(gt2.abs().max()) > 0


def test_rasterize_pyramid_config3_mesh(op):
    """BFM-size mesh (35 721 verts / 70 688 tris), batch 8, the seven GeneratorWithMap sizes: identical to seven
    single-size calls, in 3 launches (per-vertex pre-pass, triangle pass, resolve) instead of 21."""
    from stylerenderer_b200 import _lib
    v, tri = grid_mesh(189, 8, 4242, jitter=0.002)
    tex = torch.nn.functional.normalize(seeded((8, 189 * 189, 3), 4243), dim=-1)
    vc, tc, fc = cuda(v), cuda(tex), cuda(tri)
    n0 = _lib.launch_count()
    outs = op.rasterize_pyramid(vc, tc, fc, PYRAMID_SIZES)
    n1 = _lib.launch_count()
    singles = [op.rasterize(vc, tc, fc, s) for s in PYRAMID_SIZES]
    n2 = _lib.launch_count()
    for s, o, o1 in zip(PYRAMID_SIZES, outs, singles):
        assert torch.equal(o, o1), s
    assert n1 - n0 == 3 and n2 - n1 == 3 * len(PYRAMID_SIZES), (n1 - n0, n2 - n1)
