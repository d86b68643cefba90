"""Pin the oracle (oracle/sr_oracle.c + oracle/torch_ref.py) BEFORE trusting it:
  1. against the reference's only known-answer test (op/rasterize.py:83-107, values in SURVEY.md section 4),
  2. against fixtures produced by running the unmodified reference (tests/golden/make_golden.py),
  3. bit-exactly against the reference's own CPU code compiled into oracle/_ref (when present).
CPU only."""
import numpy as np
import pytest
import torch

from make_golden import det_fill, grid_mesh, seeded
from oracle import cpu as O
from oracle import torch_ref as T

SELFTEST_CH0 = np.array([[.05, 0, 0, 0, 0], [.25, .15, .05, 0, 0], [.45, .35, .25, .15, .05],
                         [.65, .55, .45, 0, 0], [.85, 0, 0, 0, 0]])


def test_rasterize_known_answer():
    v = torch.tensor([[[-1, -1, 0], [-1, 1, 0], [1, 0, 0]]], dtype=torch.float64)
    f = torch.tensor([[2, 1, 0]])
    t = torch.tensor([[[1, 0], [0, 1], [0, 0]]], dtype=torch.float64)
    out, ind, coeff = O.rasterize(v, t, f, 5)
    np.testing.assert_allclose(out[0, :, :, 0].numpy(), SELFTEST_CH0, atol=1e-12)
    np.testing.assert_allclose(out[0, :, :, 1].numpy(), SELFTEST_CH0[::-1], atol=1e-12)
    assert int((ind.sum(-1) > 0).sum()) == 13          # quirk 9: [2,1,0] renders 13 px ...
    _, ind2, _ = O.rasterize(v, t, torch.tensor([[0, 1, 2]]), 5)
    assert int(ind2.abs().sum()) == 0                   # ... and the opposite winding none
    out32, _, _ = O.rasterize(v.float(), t.float(), f, 5)
    np.testing.assert_allclose(out32[0, :, :, 0].numpy(), SELFTEST_CH0, atol=1e-6)


def test_rasterize_selftest_grads(golden):
    g = golden["rasterize"]["selftest"]
    out, ind, coeff = O.rasterize(g["v"], g["t"], g["f"], 5)
    assert torch.equal(out, g["out"])
    gv, gt = O.rasterize_grads(g["v"], g["t"], ind, coeff, g["go"])
    # the reference scatter-add runs through a float32 sparse.mm even for float64 inputs (op/rasterize.py:63,76)
    torch.testing.assert_close(gv, g["gv"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gt, g["gt"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["grid24_h32_f32", "grid24_h8_f32", "grid16_h16_f64"])
def test_rasterize_golden(golden, name):
    g = golden["rasterize"][name]
    v, tri = grid_mesh(g["n"], g["b"], g["seed"], dtype=g["tex"].dtype)
    out, ind, coeff = O.rasterize(v, g["tex"], tri, g["h"])
    assert torch.equal(ind, g["ind"].long())             # integer buffers: bit exact
    assert torch.equal(coeff, g["coeff"])
    tol = dict(rtol=1e-5, atol=1e-6) if v.dtype == torch.float32 else dict(rtol=1e-12, atol=1e-13)
    torch.testing.assert_close(out, g["out"], **tol)
    gv, gt = O.rasterize_grads(v, g["tex"], ind, coeff, g["go"])
    gtol = dict(rtol=1e-4, atol=1e-4) if v.dtype == torch.float32 else dict(rtol=1e-4, atol=1e-5)  # float32 sparse.mm inside the reference
    torch.testing.assert_close(gv, g["gv"], **gtol)
    torch.testing.assert_close(gt, g["gt"], **gtol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("h", [4, 16, 64, 100])
@pytest.mark.parametrize("perspective", [False, True])
def test_rasterize_vs_compiled_reference(ref_ext, dtype, h, perspective):
    R = ref_ext("ref_rasterize")
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    v, tri = grid_mesh(40, 3, 1000 + h, dtype=dtype)
    if perspective:
        v[..., 2] -= 3
    ind, coeff, _ = O.rasterize_forward(v, tri, h, 0, perspective, 1e-6)
    ri, rc = R.forward(v, tri, h, 0, perspective, 1e-6)
    assert torch.equal(ind, ri) and torch.equal(coeff, rc)
    assert torch.equal(O.rasterize_backward(v, ind, perspective, 1e-6), R.backward(v, ri, perspective, 1e-6))


def test_rasterize_degenerate_and_edge_cases(ref_ext):
    R = ref_ext("ref_rasterize")
    if R is None:
        pytest.skip("oracle/_ref not built")
    g = torch.Generator().manual_seed(5)
    # random soup incl. zero-area triangles, repeated vertices, out-of-range ids, off-screen triangles
    v = (torch.rand(2, 50, 3, generator=g) * 2.6 - 1.3)
    v[:, 10] = v[:, 11]
    v[:, 12, :2] = v[:, 13, :2]
    tri = torch.randint(0, 50, (400, 3), generator=g)
    tri[5] = torch.tensor([10, 11, 20]); tri[6] = torch.tensor([7, 7, 7]); tri[7] = torch.tensor([0, 60, 1])
    tri[8] = torch.tensor([-1, 2, 3]); tri[9] = torch.tensor([12, 13, 12])
    for h in (1, 7, 33):
        ind, coeff, _ = O.rasterize_forward(v, tri, h, 0, False, 1e-6)
        ri, rc = R.forward(v, tri, h, 0, False, 1e-6)
        assert torch.equal(ind, ri) and torch.equal(coeff, rc)
    # per-batch triangles and the unbatched [n,3] / [f,3] form
    trib = torch.stack([tri, tri.flip(0)])
    ind, coeff, _ = O.rasterize_forward(v, trib, 16)
    ri, rc = R.forward(v, trib, 16, 0, False, 1e-9)
    assert torch.equal(ind, ri) and torch.equal(coeff, rc)
    ind, coeff, _ = O.rasterize_forward(v[0], tri, 16)
    ri, rc = R.forward(v[0], tri, 16, 0, False, 1e-9)
    assert ind.shape == ri.shape and torch.equal(ind, ri) and torch.equal(coeff, rc)


def test_upfirdn2d_golden(golden):
    for name, g in golden["upfirdn2d"].items():
        y = O.upfirdn2d(g["x"], g["k"], g["up"], g["down"], g["pad"])
        assert y.shape == g["y"].shape, name
        torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=1e-6, msg=name)
        torch.testing.assert_close(T.upfirdn2d(g["x"], g["k"], g["up"], g["down"], g["pad"]), g["y"],
                                   rtol=1e-6, atol=1e-7, msg=name)


def test_fused_bias_act_golden_and_reference(golden, ref_ext):
    for name, g in golden["fused_leaky_relu"].items():
        y = O.fused_leaky_relu(g["x"], g["b"])
        torch.testing.assert_close(y, g["y"], rtol=1e-6, atol=1e-7, msg=name)
        assert torch.equal(T.fused_leaky_relu(g["x"], g["b"]), g["y"])
    F = ref_ext("ref_fused")
    if F is None:
        pytest.skip("oracle/_ref not built")
    x = seeded((3, 6, 5, 4), 1); b = seeded((6,), 2); ref = seeded((3, 6, 5, 4), 3)
    empty = x.new_empty(0)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0), (1, 1), (1, 2)]:
        for bb, rr in [(b, empty), (empty, ref), (b, ref)]:
            want = F.fused_bias_act(x, bb, rr, act, grad, 0.2, 2 ** 0.5)
            got = O.fused_bias_act(x, bb, rr, act, grad, 0.2, 2 ** 0.5)
            assert torch.equal(got, want), (act, grad)


# ------------------------------------------------------------------ module-level restatement
def _grads(mod, args, wrt, gy):
    y = mod(*args)
    params = [p for _, p in sorted(mod.named_parameters())]
    gr = torch.autograd.grad(y, wrt + params, gy, allow_unused=True)
    return y.detach(), gr[:len(wrt)], dict(zip([n for n, _ in sorted(mod.named_parameters())], gr[len(wrt):]))


def _check_param_grads(got, want, tol):
    assert set(got) == set(want)
    for k in want:
        if want[k] is None:
            assert got[k] is None or float(got[k].abs().max()) == 0
        else:
            torch.testing.assert_close(got[k], want[k], msg=k, **tol)


TOL = dict(rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", ["modconv_plain", "modconv_up", "modconv_1x1_nodemod"])
def test_torch_ref_modconv(golden, name):
    g = golden["modules"][name]
    kw = dict(g["kw"])
    m = det_fill(T.ModulatedConv2d(kw.pop("in_channel"), kw.pop("out_channel"), kw.pop("kernel_size"),
                                   kw.pop("style_dim"), **kw), 500)
    x = g["x"].clone().requires_grad_(True); s = g["style"].clone().requires_grad_(True)
    y, (gx, gs), gp = _grads(m, (x, s), [x, s], g["gy"])
    torch.testing.assert_close(y, g["y"], **TOL)
    torch.testing.assert_close(gx, g["gx"], **TOL)
    torch.testing.assert_close(gs, g["gs"], **TOL)
    _check_param_grads(gp, g["gp"], TOL)


def test_torch_ref_styled_blocks(golden):
    mods = golden["modules"]
    for name, up in [("styledconv_plain", False), ("styledconv_up", True)]:
        g = mods[name]
        m = det_fill(T.StyledConv(8, 12, 3, 32, upsample=up), 501)
        x = g["x"].clone().requires_grad_(True); s = g["style"].clone().requires_grad_(True)
        y, (gx, gs), gp = _grads(m, (x, s, g["noise"]), [x, s], g["gy"])
        torch.testing.assert_close(y, g["y"], **TOL)
        torch.testing.assert_close(gx, g["gx"], **TOL)
        torch.testing.assert_close(gs, g["gs"], **TOL)
        _check_param_grads(gp, g["gp"], TOL)
    g = mods["styledmapconv"]
    m = det_fill(T.StyledMapConv(8, 12, 3, 32), 502)
    x = g["x"].clone().requires_grad_(True); s = g["style"].clone().requires_grad_(True)
    sm = g["stylemap"].clone().requires_grad_(True)
    y, (gx, gs, gm), gp = _grads(m, (x, s, sm, g["noise"]), [x, s, sm], g["gy"])
    torch.testing.assert_close(y, g["y"], **TOL)
    torch.testing.assert_close(gm, g["gm"], **TOL)
    _check_param_grads(gp, g["gp"], TOL)
    g = mods["torgb"]
    m = det_fill(T.ToRGB(8, 32), 503)
    x = g["x"].clone().requires_grad_(True); s = g["style"].clone().requires_grad_(True)
    sk = g["skip"].clone().requires_grad_(True)
    y, (gx, gs, gk), gp = _grads(m, (x, s, sk), [x, s, sk], g["gy"])
    torch.testing.assert_close(y, g["y"], **TOL)
    torch.testing.assert_close(gk, g["gk"], **TOL)
    _check_param_grads(gp, g["gp"], TOL)


def test_torch_ref_networks(golden):
    nets = golden["networks"]
    g = nets["generator32"]
    G = det_fill(T.Generator(32, 64, 2), 600).eval()
    assert len(G.state_dict()) == g["n_keys"]           # incl. the duplicated ToRGB list (quirk 4)
    z = g["z"].clone().requires_grad_(True)
    img, _ = G([z], randomize_noise=False)
    torch.testing.assert_close(img, g["img"], rtol=1e-3, atol=1e-4)
    gz, gw = torch.autograd.grad(img, (z, G.convs[3].conv.weight), g["gimg"])
    torch.testing.assert_close(gz, g["gz"], rtol=2e-3, atol=1e-4 * float(g["gz"].abs().max()))
    torch.testing.assert_close(gw[0, :4, :4], g["gw_convs3_slice"], rtol=2e-3,
                               atol=1e-4 * float(g["gw_convs3_slice"].abs().max()))
    g = nets["generatorwithmap16"]
    v, tri = grid_mesh(24, 2, 611)
    GM = det_fill(T.GeneratorWithMap(16, 64, 2, rasterize=lambda v_, t_, f_, h, w: O.rasterize(v_, t_, f_, h, w)[0]),
                  610).eval()
    assert len(GM.state_dict()) == g["n_keys"]
    img, _, normals = GM([g["z"]], (v, g["tex"], tri), return_normals=True, randomize_noise=False)
    torch.testing.assert_close(normals[-1], g["normal16"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(img, g["img"], rtol=1e-3, atol=1e-4)
    g = nets["discriminator16"]
    D = det_fill(T.Discriminator(16), 620).eval()
    torch.testing.assert_close(D(g["x"]), g["y"], rtol=1e-3, atol=1e-4)
