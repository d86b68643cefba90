"""GPU parity of the module layer (stylerenderer_b200.layers / .model) against fixtures produced by the unmodified
reference on CPU (tests/golden/make_golden.py) -- bar: 1e-3 relative in fp32 (BASELINE.json north_star)."""
import pytest
import torch

from make_golden import det_fill, grid_mesh

pytestmark = pytest.mark.gpu

REL = 1e-3


def close(got, want, what=""):
    """max-norm relative error <= 1e-3 (the north_star's "within 1e-3 rel fp32")."""
    got = got.detach().cpu()
    err = float((got - want).abs().max())
    scale = max(float(want.abs().max()), 1e-12)
    assert got.shape == want.shape, what
    assert err <= REL * scale, f"{what}: max abs err {err:.3e} vs max |ref| {scale:.3e} (rel {err / scale:.2e})"


@pytest.fixture(scope="module", autouse=True)
def fp32_math():
    """Parity runs pin true-fp32 library math; the tcgen05 path states its own precision."""
    assert torch.cuda.is_available()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def grads(mod, args, wrt, gy):
    y = mod(*args)
    names = [n for n, _ in sorted(mod.named_parameters())]
    params = [p for _, p in sorted(mod.named_parameters())]
    gr = torch.autograd.grad(y, wrt + params, gy.to(y.device), allow_unused=True)
    return y.detach(), gr[:len(wrt)], dict(zip(names, gr[len(wrt):]))


def check_param_grads(got, want):
    assert set(got) == set(want)
    for k, w in want.items():
        if w is None:
            assert got[k] is None or float(got[k].abs().max()) == 0, k
        else:
            close(got[k], w, k)


@pytest.mark.parametrize("backend", ["cudnn", "tcgen05"])
@pytest.mark.parametrize("name", ["modconv_plain", "modconv_up", "modconv_1x1_nodemod"])
def test_modulated_conv(golden, name, backend):
    from stylerenderer_b200 import layers as L
    if backend == "tcgen05" and not getattr(L, "HAVE_TCGEN05", False):
        pytest.skip("tcgen05 modulated conv not built yet")
    L.set_conv_backend(backend)
    try:
        g = golden["modules"][name]
        m = det_fill(L.ModulatedConv2d(**g["kw"]), 500).cuda()
        x = g["x"].cuda().requires_grad_(True)
        s = g["style"].cuda().requires_grad_(True)
        y, (gx, gs), gp = grads(m, (x, s), [x, s], g["gy"])
        close(y, g["y"], "y"); close(gx, g["gx"], "gx"); close(gs, g["gs"], "gs")
        check_param_grads(gp, g["gp"])
    finally:
        L.set_conv_backend("cudnn")


def test_styled_blocks(golden):
    from stylerenderer_b200 import model as M
    mods = golden["modules"]
    for name, up in [("styledconv_plain", False), ("styledconv_up", True)]:
        g = mods[name]
        m = det_fill(M.StyledConv(8, 12, 3, 32, upsample=up), 501).cuda()
        x = g["x"].cuda().requires_grad_(True); s = g["style"].cuda().requires_grad_(True)
        y, (gx, gs), gp = grads(m, (x, s, g["noise"].cuda()), [x, s], g["gy"])
        close(y, g["y"], name); close(gx, g["gx"], name); close(gs, g["gs"], name)
        check_param_grads(gp, g["gp"])
    g = mods["styledmapconv"]
    m = det_fill(M.StyledMapConv(8, 12, 3, 32), 502).cuda()
    x = g["x"].cuda().requires_grad_(True); s = g["style"].cuda().requires_grad_(True)
    sm = g["stylemap"].cuda().requires_grad_(True)
    y, (gx, gs, gm), gp = grads(m, (x, s, sm, g["noise"].cuda()), [x, s, sm], g["gy"])
    close(y, g["y"]); close(gx, g["gx"]); close(gm, g["gm"])
    check_param_grads(gp, g["gp"])
    g = mods["torgb"]
    m = det_fill(M.ToRGB(8, 32), 503).cuda()
    x = g["x"].cuda().requires_grad_(True); s = g["style"].cuda().requires_grad_(True)
    sk = g["skip"].cuda().requires_grad_(True)
    y, (gx, gs, gk), gp = grads(m, (x, s, sk), [x, s, sk], g["gy"])
    close(y, g["y"]); close(gx, g["gx"]); close(gk, g["gk"])
    check_param_grads(gp, g["gp"])


def test_networks(golden):
    from stylerenderer_b200 import model as M
    nets = golden["networks"]
    g = nets["generator32"]
    G = det_fill(M.Generator(32, 64, 2), 600).cuda().eval()
    assert len(G.state_dict()) == g["n_keys"]
    z = g["z"].cuda().requires_grad_(True)
    img, _ = G([z], randomize_noise=False)
    close(img, g["img"], "generator image")
    gz, gw = torch.autograd.grad(img, (z, G.convs[3].conv.weight), g["gimg"].cuda())
    # dz is ill-conditioned: the reference's own fp32 result is 3.3e-4 (max-norm rel) off its fp64 evaluation
    err = float((gz.cpu() - g["gz"]).abs().max() / g["gz"].abs().max())
    assert err < 3e-3, f"dz rel err {err:.2e}"
    # cuDNN's fp32 backward algorithms (non-fused Winograd) are ~1e-3 accurate here: measured 1.3e-3 vs fp64 while
    # the reference's CPU fp32 is 7e-5 (scratch/debug_modules.py); the hand-written path is checked separately
    err = float((gw[0, :4, :4].cpu() - g["gw_convs3_slice"]).abs().max() / g["gw_convs3_slice"].abs().max())
    assert err < 5e-3, f"dW slice rel err {err:.2e}"
    assert abs(float(gw.norm()) - float(g["gw_convs3_norm"])) <= REL * float(g["gw_convs3_norm"])
    g = nets["generatorwithmap16"]
    v, tri = grid_mesh(24, 2, 611)
    GM = det_fill(M.GeneratorWithMap(16, 64, 2), 610).cuda().eval()
    assert len(GM.state_dict()) == g["n_keys"]
    img, _, normals = GM([g["z"].cuda()], (v.cuda(), g["tex"].cuda(), tri.cuda()), return_normals=True,
                         randomize_noise=False)
    close(normals[-1], g["normal16"], "rasterized normals")
    close(img, g["img"], "GAR image")
    g = nets["discriminator16"]
    D = det_fill(M.Discriminator(16), 620).cuda().eval()
    close(D(g["x"].cuda()), g["y"], "discriminator logits")


# ------------------------------------------------------------------ tcgen05 StyledConv block vs the oracle
@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 128, 8), (3, 128, 256, 16), (1, 256, 128, 32)])
def test_styled_conv_tcgen05_vs_oracle(up, shape):
    """The fused tcgen05 StyledConv (fwd + all gradients) against the CPU restatement of the reference, at channel
    counts the tensor-core path supports.  TF32 multiplicands: the bar is the north_star's 1e-3 (max-norm relative)."""
    from oracle import torch_ref as T
    from stylerenderer_b200 import layers as L, model as M
    from make_golden import seeded
    b, cin, cout, r = shape
    ref = det_fill(T.StyledConv(cin, cout, 3, 64, upsample=up), 700)
    mod = det_fill(M.StyledConv(cin, cout, 3, 64, upsample=up), 700).cuda()
    x, style = seeded((b, cin, r, r), 701), seeded((b, 64), 702)
    ro = 2 * r if up else r
    noise, gy = seeded((b, 1, ro, ro), 703), seeded((b, cout, ro, ro), 704)
    xr, sr = x.clone().requires_grad_(True), style.clone().requires_grad_(True)
    want_y, want_g, want_p = grads(ref, (xr, sr, noise), [xr, sr], gy)
    L.set_conv_backend("tcgen05")
    try:
        xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        sc = style.cuda().requires_grad_(True)
        got_y, got_g, got_p = grads(mod, (xc, sc, noise.cuda()), [xc, sc], gy)
    finally:
        L.set_conv_backend("cudnn")
    close(got_y, want_y, "y")
    # Gradients in the shipped tf32 mode.  A tf32 forward moves pre-activations by ~3e-4, which flips the leaky-ReLU mask of
    # the few elements that sit that close to zero, and one flip is a percent-level change of a max-norm at these sizes.
    # The oracle is therefore evaluated once more with the KERNEL'S OWN mask in its leaky-ReLU backward (the forward is
    # untouched): what remains is the arithmetic of the tf32 backward kernels on generic inputs, held to 1e-3.
    mask = (got_y.detach().cpu() > 0)
    flips = int((mask != (want_y > 0)).sum())

    class MaskedLReLU(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return torch.nn.functional.leaky_relu(t, 0.2) * 2 ** 0.5

        @staticmethod
        def backward(ctx, g):
            return torch.where(mask, g, g * 0.2) * 2 ** 0.5

    old = T.fused_leaky_relu
    T.fused_leaky_relu = lambda t, bias, negative_slope=0.2, scale=2 ** 0.5: MaskedLReLU.apply(t + bias.view(1, -1, 1, 1))
    try:
        xr, sr = x.clone().requires_grad_(True), style.clone().requires_grad_(True)
        _, want_g, want_p = grads(ref, (xr, sr, noise), [xr, sr], gy)
    finally:
        T.fused_leaky_relu = old
    print(f"StyledConv {shape} up={up}: {flips} of {mask.numel()} mask elements differ between the tf32 forward and the oracle")
    close(got_g[0], want_g[0], "gx"); close(got_g[1], want_g[1], "gs")
    check_param_grads(got_p, {k_: (v_.detach() if v_ is not None else None) for k_, v_ in want_p.items()})


def _exact_inputs(b, cin, cout, r, up):
    """Operands whose products and partial sums are exactly representable: the tf32 forward is then bit-identical to
    an fp32/fp64 one (no leaky-ReLU mask flips) and the backward kernels can be checked to 1e-3."""
    g = torch.Generator().manual_seed(900 + cin + cout + r)
    x = torch.randint(-8, 9, (b, cin, r, r), generator=g).float() / 4
    w = torch.randint(-2, 3, (1, cout, cin, 3, 3), generator=g).float() / 2
    s = torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (b, cin), generator=g)]
    d = torch.rand(b, cout, generator=g) + 0.5
    ro = 2 * r if up else r
    noise = torch.randn(b, 1, ro, ro, generator=g)
    nw = torch.tensor([0.37])
    bias = torch.randn(cout, generator=g) * 0.2
    gy = torch.randn(b, cout, ro, ro, generator=g)
    return x, w, s, d, noise, nw, bias, gy


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 128, 8), (1, 256, 128, 16), (3, 128, 256, 4), (2, 128, 128, 32)])
def test_styled_conv_function_exact_forward(up, shape):
    import torch.nn.functional as F
    from stylerenderer_b200 import fused
    b, cin, cout, r = shape
    scale, alpha, gain = 2.0 ** -5, 0.2, 2 ** 0.5
    x, w, s, d, noise, nw, bias, gy = _exact_inputs(b, cin, cout, r, up)
    taps = torch.tensor([1., 3., 3., 1.])
    taps = (taps[None] * taps[:, None]) / 16                              # reference make_kernel * factor**2
    # fp64 reference of the same block (activation-scaling form of reference layers.py:293-323 + model.py:26-32)
    leaves = [t.double().requires_grad_(True) for t in (x, w, s, d, nw, bias)]
    xd, wd, sd, dd, nwd, bd = leaves
    xm = xd * sd.view(b, cin, 1, 1)
    if up:
        t = F.conv_transpose2d(xm, (wd[0] * scale).transpose(0, 1), stride=2) * dd.view(b, cout, 1, 1)
        t = F.conv2d(F.pad(t, [1, 1, 1, 1]).reshape(1, b * cout, 2 * r + 3, 2 * r + 3),
                     taps.double().flip(0, 1).view(1, 1, 4, 4).repeat(b * cout, 1, 1, 1), groups=b * cout).view(b, cout, 2 * r, 2 * r)
    else:
        t = F.conv2d(xm, wd[0] * scale, padding=1) * dd.view(b, cout, 1, 1)
    want_y = F.leaky_relu(t + nwd * noise.double() + bd.view(1, -1, 1, 1), alpha) * gain
    want = torch.autograd.grad(want_y, leaves, gy.double())
    cu = [t.cuda().requires_grad_(True) for t in (x, w, s, d, nw, bias)]
    xc = cu[0].detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got_y = fused.StyledConvTC.apply(xc, cu[1], cu[2], cu[3], noise.cuda(), cu[4], cu[5], scale, up, taps.cuda(), alpha, gain)
    got = torch.autograd.grad(got_y, [xc] + cu[1:], gy.cuda())
    close(got_y, want_y.detach().float(), "y")
    assert float((got_y.detach().cpu().double() - want_y.detach()).abs().max()) < 1e-4      # forward is (nearly) exact
    for name, g_, w_ in zip(["dx", "dweight", "ds", "dd", "dnoise_w", "dbias"], got, want):
        close(g_, w_.float(), name)


@pytest.mark.parametrize("with_map", [False, True])
@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 128, 8), (1, 128, 256, 16)])
def test_styled_layer_chain_exact_forward(up, shape, with_map):
    """The chained block (pre-modulated input, second output for the next layer, fused ToRGB; with_map: the StyledMapConv
    affine `t * map0 + map1` of reference model.py:50 and the gradients of both maps) against fp64 autograd."""
    import torch.nn.functional as F
    from stylerenderer_b200 import fused
    b, cin, cout, r = shape
    scale, alpha, gain = 2.0 ** -5, 0.2, 2 ** 0.5
    x, w, s, d, noise, nw, bias, gy = _exact_inputs(b, cin, cout, r, up)
    g = torch.Generator().manual_seed(77)
    xs = x * s.view(b, cin, 1, 1)                                          # exact (s is a power of two)
    s_next = torch.rand(b, cout, generator=g) + 0.5
    wb = torch.randn(b, 3, cout, generator=g) * 0.1 if not up else None
    ro = 2 * r if up else r
    g_rgb = torch.randn(b, ro, ro, 3, generator=g)
    taps = torch.tensor([1., 3., 3., 1.])
    taps = (taps[None] * taps[:, None]) / 16
    smap = None
    if with_map:                                         # a channel slice of a wider map tensor, map0 away from zero
        wide = torch.randn(b, 4, ro, ro, generator=g)
        wide[:, 2] = wide[:, 2].abs() + 0.5
        smap = wide[:, 2:]
    leaves = [t.double().requires_grad_(True) for t in ([xs, w, d, nw, bias, s_next] + ([wb] if wb is not None else []))]
    smd = smap.double().requires_grad_(True) if with_map else None
    xd, wd, dd, nwd, bd, snd = leaves[:6]
    if up:
        t = F.conv_transpose2d(xd, (wd[0] * scale).transpose(0, 1), stride=2) * dd.view(b, cout, 1, 1)
        t = F.conv2d(F.pad(t, [1, 1, 1, 1]).reshape(1, b * cout, 2 * r + 3, 2 * r + 3),
                     taps.double().flip(0, 1).view(1, 1, 4, 4).repeat(b * cout, 1, 1, 1), groups=b * cout).view(b, cout, ro, ro)
    else:
        t = F.conv2d(xd, wd[0] * scale, padding=1) * dd.view(b, cout, 1, 1)
    if with_map:
        t = t * smd[:, :1] + smd[:, 1:2]
    y = F.leaky_relu(t + nwd * noise.double() + bd.view(1, -1, 1, 1), alpha) * gain
    main = y * snd.view(b, cout, 1, 1)
    loss = (main * gy.double()).sum()
    if wb is not None:
        rgb_ref = torch.einsum("bchw,bkc->bhwk", y, leaves[6])
        loss = loss + (rgb_ref * g_rgb.double()).sum()
    want = torch.autograd.grad(loss, leaves + ([smd] if with_map else []))
    cu = [t.cuda().requires_grad_(True) for t in ([xs, w, d, nw, bias, s_next] + ([wb] if wb is not None else []))]
    xc = cu[0].detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    smc = None
    if with_map:
        smc = wide.cuda()[:, 2:].requires_grad_(True)   # non-contiguous batch stride (4 planes per image)
    got_main, got_rgb, _ = fused.StyledLayerTC.apply(xc, cu[1], cu[2], noise.cuda(), cu[3], cu[4], cu[5],
                                                  cu[6] if wb is not None else None, scale, up, taps.cuda(), alpha, gain,
                                                  None, None, smc)
    close(got_main, main.detach().float(), "main (tf32-rounded)")
    lo = (got_main * gy.cuda()).sum()
    if wb is not None:
        close(got_rgb, rgb_ref.detach().float(), "rgb")
        lo = lo + (got_rgb * g_rgb.cuda()).sum()
    got = torch.autograd.grad(lo, [xc] + cu[1:] + ([smc] if with_map else []))
    names = ["dxs", "dweight", "dd", "dnoise_w", "dbias", "ds_next"] + (["drgb_weight"] if wb is not None else []) + ["dstylemap"]
    for name, g_, w_ in zip(names, got, want):
        close(g_, w_.float(), name)


def test_fused_pass_kernels():
    """sr_styled_bwd_prologue_f32 / sr_scale_dot_nhwc_f32 / sr_blur_nhwc_styled_f32 against their torch definitions."""
    from stylerenderer_b200 import tc_conv as tc
    from stylerenderer_b200.op import upfirdn2d_raw
    from make_golden import seeded
    for (b, h, w, c) in [(2, 8, 8, 128), (3, 5, 7, 256), (1, 33, 33, 512), (2, 64, 64, 128)]:
        gy, y = seeded((b, h, w, c), 1).cuda(), seeded((b, h, w, c), 2).cuda()
        noise, nw = seeded((b, 1, h, w), 3).cuda(), torch.tensor([0.4], device="cuda")
        bias, d = seeded((c,), 4).cuda(), (seeded((b, c), 5).abs() + 0.5).cuda()
        for nz in (noise, noise[:1]):
            ga, gb, gnw, e = tc.bwd_prologue(gy, y, nz, nw, bias, d, 0.2, 2 ** 0.5, True)
            gp = torch.where(y > 0, gy, gy * 0.2) * 2 ** 0.5
            u = torch.where(y > 0, y / 2 ** 0.5, y / (2 ** 0.5 * 0.2)) - nw * nz.view(-1, h, w, 1) - bias
            torch.testing.assert_close(ga, gp * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-6)      # tf32 rounding
            torch.testing.assert_close(gb, gp.sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
            torch.testing.assert_close(gnw, (gp * nz.view(-1, h, w, 1)).sum().view(1), rtol=1e-4, atol=1e-2)
            torch.testing.assert_close(e, (gp * u).sum((1, 2)), rtol=1e-4, atol=1e-2)
        g2, gb2, gnw2, e2 = tc.bwd_prologue(gy, y, noise, nw, bias, None, 0.2, 2 ** 0.5, False)
        assert e2 is None and torch.equal(g2, gp)
        out, dot = tc.scale_dot(gy, y, d, False)
        assert torch.equal(out, gy * d.view(b, 1, 1, c))
        torch.testing.assert_close(dot, (gy * y).sum((1, 2)), rtol=1e-4, atol=1e-2)
        out, _ = tc.scale_dot(gy, None, d, True)
        torch.testing.assert_close(out, gy * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-6)
        k = seeded((4, 4), 6).cuda()
        t = seeded((b, h + 1, w + 1, c), 7).cuda()
        got = tc.blur_styled(t, k, (1, 1), noise, nw, bias, 0.2, 2 ** 0.5)
        ref = upfirdn2d_raw(t, k, 1, 1, 1, 1, 1, 1, 1, 1) + nw * noise.view(b, h, w, 1) + bias
        ref = torch.where(ref > 0, ref, ref * 0.2) * 2 ** 0.5
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
        got1, got2 = tc.blur_styled(t, k, (1, 1), noise, nw, bias, 0.2, 2 ** 0.5, scale2=d)
        assert torch.equal(got1, got)
        torch.testing.assert_close(got2, ref * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-6)
        other = seeded((b, h + 2, w + 2, c), 8).cuda()
        o, dt = tc.blur_scaledot(t, k, (2, 2), d, other)                      # [b,h+1,w+1,c] -> [b,h+2,w+2,c]
        f = upfirdn2d_raw(t, k, 1, 1, 1, 1, 2, 2, 2, 2)
        torch.testing.assert_close(o, f * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-6)
        torch.testing.assert_close(dt, (f * other).sum((1, 2)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("variant", ["stream", "sep", "window"])
@pytest.mark.parametrize("rank1", [True, False])
def test_fir_nhwc_kernel_variants_vs_oracle(variant, rank1, monkeypatch):
    """Every channels-last FIR kernel ("stream" = the default: TMA row-streaming kernel for planes >= 32 x 32 with C % 32 == 0,
    per-thread kernels below that; SR_FIR_STREAM=0: "sep" = the separable register-ring kernel everywhere, "window" =
    SR_FIR_SEP=0, the 4 x 5 input-window kernel of round 1) in its three modes -- plain, styled tail, scale(+dot) tail -- against the oracle's upfirdn2d
    (oracle/sr_oracle.c, reference op/upfirdn2d.py:159-200), with the model's rank-1 taps (separable form inside the
    kernels, reference layers.py:7-12) and with general, asymmetric taps (2-D form).  Shapes cover several column strips
    with a ragged last strip (65, 35 wide), several row segments (129 rows) and the sub-32 fall-back."""
    from oracle import cpu as O
    from stylerenderer_b200 import tc_conv as tc
    from stylerenderer_b200.op import upfirdn2d_raw
    from make_golden import seeded
    if variant != "stream":
        monkeypatch.setenv("SR_FIR_STREAM", "0")
        if variant == "window":
            monkeypatch.setenv("SR_FIR_SEP", "0")
    k1 = torch.tensor([1., 3., 3., 1.])
    k = (torch.outer(k1, k1) / 64 * 4) if rank1 else seeded((4, 4), 16)
    shapes = [(2, 8, 8, 128), (3, 5, 7, 256), (1, 33, 35, 512), (2, 64, 64, 128), (1, 3, 2, 4)]
    if variant == "stream":
        shapes += [(1, 70, 65, 64), (2, 129, 40, 32), (1, 32, 32, 96)]
    for (b, h, w, c) in shapes:
        t = seeded((b, h + 1, w + 1, c), 17)
        noise, nw = seeded((b, 1, h, w), 18), torch.tensor([0.4])
        bias, d = seeded((c,), 19), seeded((b, c), 20).abs() + 0.5

        def fir(pad):
            return O.upfirdn2d(t.permute(0, 3, 1, 2).contiguous(), k, pad=(pad, pad)).permute(0, 2, 3, 1).contiguous()
        f1, f2 = fir(1), fir(2)                                            # [b,h,w,c] and [b,h+2,w+2,c]
        tol = dict(rtol=1e-5, atol=1e-5)
        got = upfirdn2d_raw(t.cuda(), k.cuda(), 1, 1, 1, 1, 1, 1, 1, 1)    # plain mode
        torch.testing.assert_close(got.cpu(), f1, **tol)
        ref = f1 + nw * noise.view(b, h, w, 1) + bias
        ref = torch.where(ref > 0, ref, ref * 0.2) * 2 ** 0.5
        got1, got2 = tc.blur_styled(t.cuda(), k.cuda(), (1, 1), noise.cuda(), nw.cuda(), bias.cuda(), 0.2, 2 ** 0.5,
                                    scale2=d.cuda())
        torch.testing.assert_close(got1.cpu(), ref, **tol)
        torch.testing.assert_close(got2.cpu(), ref * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-5)     # tf32 rounding
        other = seeded((b, h + 2, w + 2, c), 21)
        o, dt = tc.blur_scaledot(t.cuda(), k.cuda(), (2, 2), d.cuda(), other.cuda())
        torch.testing.assert_close(o.cpu(), f2 * d.view(b, 1, 1, c), rtol=6e-4, atol=1e-5)
        torch.testing.assert_close(dt.cpu(), (f2 * other).sum((1, 2)), rtol=1e-4, atol=1e-2)
        o2, none = tc.blur_scaledot(t.cuda(), k.cuda(), (2, 2), d.cuda())
        assert none is None and torch.equal(o2, o)
        gbc = tc.blur_styled(t.cuda(), k.cuda(), (1, 1), noise[:1].cuda(), nw.cuda(), bias.cuda(), 0.2, 2 ** 0.5)   # broadcast noise
        refb = f1 + nw * noise[:1].view(1, h, w, 1) + bias
        torch.testing.assert_close(gbc.cpu(), torch.where(refb > 0, refb, refb * 0.2) * 2 ** 0.5, **tol)
        # StyledMapConv tail: y = lrelu(fir * map0 + map1 + noise + bias) * gain; `out` keeps the FIR output, out2 sees y
        smap = seeded((b, 2, h, w), 22)
        refm = f1 * smap[:, 0].view(b, h, w, 1) + smap[:, 1].view(b, h, w, 1) + nw * noise.view(b, h, w, 1) + bias
        refm = torch.where(refm > 0, refm, refm * 0.2) * 2 ** 0.5
        gm1, gm2 = tc.blur_styled(t.cuda(), k.cuda(), (1, 1), noise.cuda(), nw.cuda(), bias.cuda(), 0.2, 2 ** 0.5,
                                  scale2=d.cuda(), stylemap=smap.cuda())
        torch.testing.assert_close(gm1.cpu(), f1, **tol)
        torch.testing.assert_close(gm2.cpu(), refm * d.view(b, 1, 1, c), rtol=6e-4, atol=2e-5)
        # bfloat16 GEMM operands (tcgen05 kind::f16): the fp32 output is unchanged, the operand is rounded to bf16
        with tc.precision("bf16"):
            gb1, gb2 = tc.blur_styled(t.cuda(), k.cuda(), (1, 1), noise.cuda(), nw.cuda(), bias.cuda(), 0.2, 2 ** 0.5,
                                      scale2=d.cuda())
            ob, _ = tc.blur_scaledot(t.cuda(), k.cuda(), (2, 2), d.cuda())
        assert gb2.dtype == torch.bfloat16 and ob.dtype == torch.bfloat16
        torch.testing.assert_close(gb1.cpu(), ref, **tol)
        torch.testing.assert_close(gb2.float().cpu(), ref * d.view(b, 1, 1, c), rtol=5e-3, atol=1e-4)
        torch.testing.assert_close(ob.float().cpu(), f2 * d.view(b, 1, 1, c), rtol=5e-3, atol=1e-4)


@pytest.mark.parametrize("up", [False, True])
def test_modulated_conv_tcgen05_exact_forward(up):
    """ModulatedConv2d alone on the tensor cores (the path StyledMapConv and user code take), exact-product operands."""
    import torch.nn.functional as F
    from stylerenderer_b200 import fused
    b, cin, cout, r = 2, 128, 128, 8
    scale = 2.0 ** -5
    x, w, s, d, _, _, _, _ = _exact_inputs(b, cin, cout, r, up)
    ro = 2 * r if up else r
    gy = torch.randn(b, cout, ro, ro, generator=torch.Generator().manual_seed(5))
    taps = torch.tensor([1., 3., 3., 1.])
    taps = (taps[None] * taps[:, None]) / 16
    leaves = [t.double().requires_grad_(True) for t in (x, w, s, d)]
    xd, wd, sd, dd = leaves
    xm = xd * sd.view(b, cin, 1, 1)
    if up:
        t = F.conv_transpose2d(xm, (wd[0] * scale).transpose(0, 1), stride=2) * dd.view(b, cout, 1, 1)
        y = F.conv2d(F.pad(t, [1, 1, 1, 1]).reshape(1, b * cout, 2 * r + 3, 2 * r + 3),
                     taps.double().flip(0, 1).view(1, 1, 4, 4).repeat(b * cout, 1, 1, 1), groups=b * cout).view(b, cout, ro, ro)
    else:
        y = F.conv2d(xm, wd[0] * scale, padding=1) * dd.view(b, cout, 1, 1)
    want = torch.autograd.grad(y, leaves, gy.double())
    cu = [t.cuda().requires_grad_(True) for t in (x, w, s, d)]
    got_y = fused.ModConvTC.apply(cu[0], cu[1], cu[2], cu[3], scale, up, taps.cuda())
    got = torch.autograd.grad(got_y, cu, gy.cuda())
    close(got_y, y.detach().float(), "y")
    for name, g_, w_ in zip(["dx", "dweight", "ds", "dd"], got, want):
        close(g_, w_.float(), name)


def test_generator_with_map_tcgen05_runs_and_matches():
    """GeneratorWithMap with conv_backend=tcgen05 (chained StyledMapConv blocks: style map in the conv epilogue / FIR tail,
    map gradients from the backward prologue) vs the composed path in true fp32: image and every gradient, including
    the ones that flow through the style maps into the rasterised normals (mesh vertices and normals)."""
    from stylerenderer_b200 import layers as L, model as M
    from make_golden import seeded, grid_mesh
    G = det_fill(M.GeneratorWithMap(32, 64, 2), 720).cuda().eval()
    v, tri = grid_mesh(24, 2, 721)
    tex = torch.nn.functional.normalize(seeded((2, 576, 3), 722), dim=-1)
    z = seeded((2, 64), 723).cuda()
    cot = seeded((2, 3, 32, 32), 724).cuda()

    def run():
        zz = z.clone().requires_grad_(True)
        vv, tt = v.cuda().requires_grad_(True), tex.cuda().requires_grad_(True)
        img, _, _ = G([zz], (vv, tt, tri.cuda()), randomize_noise=False)
        ps = [p for _, p in sorted(G.named_parameters()) if p.requires_grad]
        return img.detach(), torch.autograd.grad(img, [zz, vv, tt] + ps, cot, allow_unused=True)
    img_a, gr_a = run()
    L.set_conv_backend("tcgen05")
    try:
        img_b, gr_b = run()
    finally:
        L.set_conv_backend("cudnn")
    close(img_b, img_a.detach().cpu(), "GAR image")
    names = ["z", "verts", "tex"] + [n for n, p in sorted(G.named_parameters()) if p.requires_grad]
    from parity_util import hold_envelope, rel_err
    errs = []
    for n, a, bb in zip(names, gr_a, gr_b):
        if a is None or float(a.abs().max()) == 0:
            continue
        assert bb is not None, n
        errs.append((rel_err(bb, a), n))
    hold_envelope("gwm32_tcgen05_vs_composed[tf32]", errs, "tf32")


def test_generator_tcgen05_backend_matches_cudnn_backend():
    """Generator(64) end to end: tcgen05 backend (channels_last pipeline) vs the composed-op backend in true fp32."""
    from stylerenderer_b200 import layers as L, model as M
    from make_golden import seeded
    G = det_fill(M.Generator(64, 64, 2), 710).cuda().eval()
    z = seeded((2, 64), 711).cuda()
    cot = seeded((2, 3, 64, 64), 712).cuda()

    def run():
        zz = z.clone().requires_grad_(True)
        img, _ = G([zz], randomize_noise=False)
        ps = [p for _, p in sorted(G.named_parameters()) if p.requires_grad]
        gr = torch.autograd.grad(img, [zz] + ps, cot, allow_unused=True)
        return img.detach(), gr
    img_a, gr_a = run()
    L.set_conv_backend("tcgen05")
    try:
        img_b, gr_b = run()
    finally:
        L.set_conv_backend("cudnn")
    close(img_b, img_a.cpu(), "image")
    names = ["z"] + [n for n, p in sorted(G.named_parameters()) if p.requires_grad]
    from parity_util import hold_envelope, rel_err
    errs = [(rel_err(bb, a), n) for n, a, bb in zip(names, gr_a, gr_b) if a is not None and float(a.abs().max()) > 0]
    hold_envelope("generator64_tcgen05_vs_composed[tf32]", errs, "tf32")


def test_generator_config2_full_size_properties():
    """BASELINE.json configs[1] at its real size -- Generator(256, 512, 8), batch 32 -- where the CPU oracle is too slow to
    be the checker: (a) the chained tcgen05 path against this package's composed path in true fp32 (itself pinned to the
    reference by the golden fixtures), (b) linearity of the backward pass in the cotangent, (c) independence of an
    image from the rest of its batch (up to the rounding of the kernel variants a batch size selects)."""
    from stylerenderer_b200 import layers as L, model as M
    from make_golden import seeded
    torch.manual_seed(0)
    G = M.Generator(256, 512, 8, channel_multiplier=2)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("noise.weight") or n.endswith("activate.bias"):
                p.normal_(0, 0.1)
    G = G.cuda().eval()
    z = seeded((32, 512), 740).cuda()
    cots = [seeded((32, 3, 256, 256), 741).cuda(), seeded((32, 3, 256, 256), 742).cuda()]

    def run(zin, cot):
        zz = zin.clone().requires_grad_(True)
        img, _ = G([zz], randomize_noise=False)
        gz = torch.autograd.grad(img, [zz], cot)[0] if cot is not None else None
        return img.detach(), gz
    img_ref, gz_ref = run(z, cots[0])                                       # composed ops, fp32 (allow_tf32 off)
    L.set_conv_backend("tcgen05")
    try:
        img, gz1 = run(z, cots[0])
        _, gz2 = run(z, cots[1])
        _, gz12 = run(z, 2.0 * cots[0] - 3.0 * cots[1])
        img8, _ = run(z[:8], None)
    finally:
        L.set_conv_backend("cudnn")
    err = float((img - img_ref).abs().max() / img_ref.abs().max())
    gz_err = float((gz1 - gz_ref).abs().max() / gz_ref.abs().max())
    print(f"Generator(256) B=32: tcgen05 (tf32) vs fp32 composed path: image max-norm rel err {err:.2e}, dz max-norm rel err {gz_err:.2e}")
    # the checker here is this package's composed path on cuDNN fp32 (itself ~1e-3 from the oracle in places); the oracle
    # comparison of this very configuration at 1e-3 is tests/test_gpu_parity_tc.py::test_headline_generator256_vs_oracle
    assert err <= 1.5e-3, err                                               # measured on B200: 9.5e-4
    assert gz_err <= 5e-2, gz_err                                           # leaky-ReLU mask flips (parity_util envelope)
    lin = 2.0 * gz1 - 3.0 * gz2
    lin_err = float((gz12 - lin).abs().max() / lin.abs().max())
    assert lin_err <= 2e-3, lin_err                                         # tf32 rounding of the GEMM operands only (3.9e-4)
    # a batch of 8 selects other kernel variants than a batch of 32 (halo / im2col / CTA-pair tiles, split counts), whose
    # fp32 accumulation orders differ and can flip the tf32 rounding of a handed-on operand: a tf32-level difference
    dep = float((img8 - img[:8]).abs().max() / img.abs().max())
    print(f"batch 8 vs batch 32, same latents: max-norm rel difference {dep:.2e}; linearity error {lin_err:.2e}")
    assert dep <= 2e-3, dep                                                 # measured 4.8e-4


def test_generator_frozen_weights_latent_gradient_is_unchanged():
    """Latent inversion back-propagates to the latents only: with every parameter frozen the
    chained blocks skip their weight-gradient GEMMs, and the latent gradient must equal the one of the full backward
    (the same kernels on the same operands; only the order of float atomics inside the reductions may differ)."""
    from stylerenderer_b200 import _lib, layers as L, model as M
    from make_golden import seeded
    G = det_fill(M.Generator(64, 64, 2), 730).cuda().eval()
    w = seeded((2, G.n_latent, 64), 731).cuda()
    cot = seeded((2, 3, 64, 64), 732).cuda()

    def run():
        ww = w.clone().requires_grad_(True)
        n0 = _lib.launch_count()
        img, _ = G([ww], input_is_latent=True, randomize_noise=False)
        g, = torch.autograd.grad(img, [ww], cot)
        return img.detach(), g, _lib.launch_count() - n0
    L.set_conv_backend("tcgen05")
    try:
        img_a, g_a, n_a = run()
        for p in G.parameters():
            p.requires_grad_(False)
        img_b, g_b, n_b = run()
    finally:
        L.set_conv_backend("cudnn")
    scale = float(g_a.abs().max())
    assert float((img_a - img_b).abs().max()) <= 1e-5 * float(img_a.abs().max())
    assert float((g_a - g_b).abs().max()) <= 1e-5 * scale, float((g_a - g_b).abs().max()) / scale
    assert n_b < n_a, (n_a, n_b)                       # the wgrad / weight-gradient layout launches are gone
    print("launches per fwd+bwd: all gradients", n_a, "latents only", n_b)


def test_style_scales_all_vs_torch():
    """Batched style path (csrc/style_ops.cu: s, d for every layer in one call) against the per-layer torch formulation
    ModulatedConv2d.style_scales (reference layers.py:232-239, 295-299) in float64, values and all gradients."""
    from stylerenderer_b200 import layers as L, style
    from make_golden import seeded
    torch.manual_seed(5)
    b, n_latent, k = 5, 6, 96
    specs = [(64, 128, 3, True), (128, 40, 3, True), (40, 3, 1, False), (72, 64, 3, True)]       # cin, cout, ksize, demod
    mods = [L.ModulatedConv2d(ci, co, ks, k, demodulate=dm).cuda() for ci, co, ks, dm in specs]
    for j, m in enumerate(mods):
        with torch.no_grad():
            m.modulation.bias.copy_(seeded(m.modulation.bias.shape, 40 + j).cuda() * 0.3 + 1)
    lat_idx = [0, 3, 3, 5]
    latent = seeded((b, n_latent, k), 50).cuda().requires_grad_(True)
    got = style.style_scales_all(latent, mods, lat_idx)
    flat = [t for sd in got for t in sd if t is not None]
    cots = [seeded(t.shape, 60 + i).cuda() for i, t in enumerate(flat)]
    params = [p for m in mods for p in (m.modulation.weight, m.modulation.bias, m.weight)]
    grads = torch.autograd.grad(flat, [latent] + params, cots, allow_unused=True)
    # torch formulation in float64
    lat64 = latent.detach().double().requires_grad_(True)
    p64, want = [], []
    for m, li in zip(mods, lat_idx):
        wm, bm, w = (t.detach().double().requires_grad_(True) for t in (m.modulation.weight, m.modulation.bias, m.weight))
        p64 += [wm, bm, w]
        s = torch.nn.functional.linear(lat64[:, li], wm * m.modulation.scale, bm * m.modulation.lr_mul)
        want.append(s)
        if m.demodulate:
            wsq = (w[0] * m.scale).pow(2).sum([2, 3])
            want.append(torch.rsqrt(torch.nn.functional.linear(s * s, wsq) + m.eps))
    for g, w_ in zip(flat, want):
        torch.testing.assert_close(g.double(), w_.detach(), rtol=2e-5, atol=1e-6)
    want_g = torch.autograd.grad(want, [lat64] + p64, [c.double() for c in cots], allow_unused=True)
    for i, (g, w_) in enumerate(zip(grads, want_g)):
        if w_ is None:
            assert g is None or float(g.abs().max()) == 0, i
            continue
        torch.testing.assert_close(g.double(), w_, rtol=1e-4, atol=1e-5 * float(w_.abs().max()) + 1e-9, msg=lambda m_: f"grad {i}: {m_}")


def test_discriminator_tcgen05_backend_matches_cudnn_backend():
    """Discriminator(64) (channels 512, all ResBlock convs on the tensor-core path) vs the composed cuDNN path in true fp32:
    logits and every gradient (input, weights, both biases of a ConvLayer)."""
    from stylerenderer_b200 import layers as L, model as M
    from make_golden import seeded
    D = det_fill(M.Discriminator(64), 730).cuda().to(memory_format=torch.channels_last)
    x = seeded((4, 3, 64, 64), 731).cuda().contiguous(memory_format=torch.channels_last)

    def run():
        xx = x.clone().requires_grad_(True)
        y = D(xx)
        ps = [p for _, p in sorted(D.named_parameters())]
        return y.detach(), torch.autograd.grad(y.sum(), [xx] + ps)
    y_a, g_a = run()
    L.set_conv_backend("tcgen05")
    try:
        y_b, g_b = run()
    finally:
        L.set_conv_backend("cudnn")
    close(y_b, y_a.cpu(), "discriminator logits")
    names = ["x"] + [n for n, _ in sorted(D.named_parameters())]
    from parity_util import hold_envelope, rel_err
    errs = [(rel_err(bb, a), n) for n, a, bb in zip(names, g_a, g_b) if float(a.abs().max()) > 0]
    hold_envelope("discriminator64_tcgen05_vs_composed[tf32]", errs, "tf32")


@pytest.mark.parametrize("co", [2, 4])
@pytest.mark.parametrize("shape", [(2, 4, 4), (3, 37, 70), (1, 16, 32), (2, 256, 256)])
def test_stylemap_resblock_kernel_vs_composed_fp64(co, shape):
    """sr_stylemap_resblock_forward/backward_f32 (GeneratorWithMap's style-map nets, reference model.py:194-216: one kernel
    per direction) against the composed ResBlock of reference layers.py:379-391 evaluated in float64 on the CPU -- output
    and every parameter gradient (both biases of each layer, the 1x1 skip), ragged tiles included."""
    from stylerenderer_b200 import _lib, fused, layers as L
    from make_golden import seeded
    b, h, w = shape
    torch.manual_seed(co * 100 + h)
    blk = L.ResBlock(3, co, downsample=False)
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.3)                                     # zero biases would hide the bias paths
    x = seeded((b, 3, h, w), 5)
    x[:, :, : h // 2, : w // 3] = 0                                   # background pixels of a rasterised map are exactly 0
    gy = seeded((b, co, h, w), 6)
    ref = L.ResBlock(3, co, downsample=False).double()
    ref.load_state_dict({k: v.double() for k, v in blk.state_dict().items()})
    want = ref(x.double())
    want.backward(gy.double())
    blk = blk.cuda()
    old = L.get_conv_backend()
    L.set_conv_backend("tcgen05")
    try:
        assert fused.stylemap_resblock_supported(blk, x.cuda())
        n0 = _lib.launch_count()
        got = blk(x.cuda())
        got.backward(gy.cuda())
        assert _lib.launch_count() - n0 == 2, "one kernel forward, one backward"
    finally:
        L.set_conv_backend(old)
    scale = want.abs().max().item()
    assert (got.detach().cpu().double() - want.detach()).abs().max().item() <= 2e-6 * scale + 1e-6
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        err = (p.grad.cpu().double() - q.grad).abs().max().item()
        assert err <= 2e-5 * q.grad.abs().max().item() + 1e-5, (n, err, q.grad.abs().max().item())


@pytest.mark.parametrize("cfg", [(3, 3, 3), (3, 4, 3), (4, 3, 3), (3, 2, 1), (8, 3, 3), (1, 8, 1)])
def test_small_conv_pair_any_order(cfg):
    """fused.SmallConvFn / SmallWgradFn (sr_small_conv_f32, sr_small_conv_wgrad_f32) against F.conv2d in float64 on the CPU:
    output, first-order gradients, and the gradients of a gradient-norm penalty (the shape of the path-length regulariser,
    reference train.py:118-134, which differentiates through the backward pass of the style-map nets)."""
    import torch.nn.functional as F
    from stylerenderer_b200 import fused
    from make_golden import seeded
    ci, co, k = cfg
    b, h, w = 2, 37, 45
    x0, w0, gy = seeded((b, ci, h, w), 1), seeded((co, ci, k, k), 2) * 0.3, seeded((b, co, h, w), 3)

    def penalty(conv, x, wt):
        y = conv(x, wt)
        loss1 = (torch.tanh(y) * gy.to(y)).sum()
        gx, = torch.autograd.grad(loss1, x, create_graph=True)
        return y, (gx.pow(2).sum() + loss1)

    xr, wr = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    yr, pr = penalty(lambda a, c: F.conv2d(a, c, padding=k // 2), xr, wr)
    gxr, gwr = torch.autograd.grad(pr, [xr, wr])
    xg, wg = x0.cuda().requires_grad_(True), w0.cuda().requires_grad_(True)
    yg, pg = penalty(fused.SmallConvFn.apply, xg, wg)
    gxg, gwg = torch.autograd.grad(pg, [xg, wg])
    for name, got, want in (("y", yg, yr), ("dx", gxg, gxr), ("dw", gwg, gwr)):
        err = (got.detach().cpu().double() - want.detach()).abs().max().item()
        assert err <= 3e-5 * want.abs().max().item() + 1e-6, (name, err, want.abs().max().item())


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("shape", [(2, 128, 19, 23), (1, 256, 8, 8), (3, 64, 32, 32)])
def test_stem_conv_kernel_vs_composed_fp64(shape, channels_last):
    """sr_stem_conv_forward/backward_f32 (the Discriminator's ConvLayer(3, C, 1), reference model.py:303) against the composed
    layer in float64: output, weight / both bias gradients, input gradient; planar and channels_last images."""
    from stylerenderer_b200 import _lib, fused, layers as L
    from make_golden import seeded
    b, c, h, w = shape
    torch.manual_seed(c + h)
    layer = L.ConvLayer(3, c, 1)
    with torch.no_grad():
        layer[0].bias.normal_(0, 0.3)
        layer[1].bias.normal_(0, 0.3)
    x0, gy = seeded((b, 3, h, w), 7), seeded((b, c, h, w), 8)
    ref = L.ConvLayer(3, c, 1).double()
    ref.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
    xr = x0.double().requires_grad_(True)
    want = ref(xr)
    want.backward(gy.double())
    layer = layer.cuda()
    xg = x0.cuda()
    if channels_last:
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    old = L.get_conv_backend()
    L.set_conv_backend("tcgen05")
    try:
        n0 = _lib.launch_count()
        got = layer(xg)
        got.backward(gy.cuda().contiguous(memory_format=torch.channels_last))
        assert _lib.launch_count() - n0 == 2, "one kernel forward, one backward"
    finally:
        L.set_conv_backend(old)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)

    def close(name, a, bb, rel):
        err = (a.detach().cpu().double() - bb.detach()).abs().max().item()
        assert err <= rel * bb.abs().max().item() + 1e-6, (name, err, bb.abs().max().item())
    close("y", got, want, 2e-6)
    close("dx", xg.grad, xr.grad, 2e-5)
    for (n, p), (_, q) in zip(layer.named_parameters(), ref.named_parameters()):
        close(n, p.grad, q.grad, 3e-5)


@pytest.mark.parametrize("mode", ["tf32x3", "tf32"])
@pytest.mark.parametrize("up", [False, True])
def test_modulated_conv_double_backward_mode_matches_composed(up, mode):
    """ModulatedConv2d under layers.double_backward() on the tensor-core backend (fused.mod_conv_dd: ScaleBC / DotP
    modulation pair + ConvTC contraction, every piece differentiable to any order) against the composed torch ops in
    float64 on the CPU: output, first-order gradients, and the gradients of a path-length-style penalty (reference
    train.py:118-134: norm of the gradient w.r.t. the style, differentiated again w.r.t. the parameters)."""
    from stylerenderer_b200 import layers as L
    from make_golden import seeded
    torch.manual_seed(11)
    b, cin, cout, h = 2, 128, 128, 8
    mod = L.ModulatedConv2d(cin, cout, 3, 64, upsample=up)
    x0, w0 = seeded((b, cin, h, h), 50), seeded((b, 64), 51)
    oh = 2 * h if up else h
    noise = seeded((b, cout, oh, oh), 52) / oh

    def run(m, x, w):
        y = m(x, w)
        g, = torch.autograd.grad((torch.tanh(y) * noise.to(y)).sum(), w, create_graph=True)
        pen = g.pow(2).sum(1).sqrt().sum() + y.square().mean()
        params = [p for p in m.parameters()]
        grads = torch.autograd.grad(pen, params + [x], allow_unused=True)
        return y, pen, grads

    ref = L.ModulatedConv2d(cin, cout, 3, 64, upsample=up).double()
    ref.load_state_dict({k: v.double() for k, v in mod.state_dict().items()})
    yr, pr, gr = run(ref, x0.double().requires_grad_(True), w0.double().requires_grad_(True))
    mod = mod.cuda()
    noise = noise.cuda()
    old = L.get_conv_backend()
    L.set_conv_backend("tcgen05")
    from stylerenderer_b200 import tc_conv as tc
    try:
        with L.double_backward(), tc.precision(mode):
            xg = x0.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
            yg, pg, gg = run(mod, xg, w0.cuda().requires_grad_(True))
    finally:
        L.set_conv_backend(old)
    tol = 1e-3 if mode == "tf32x3" else 3e-2                          # tf32: second-order quantities through tf32 contractions

    def rel(a, b_):
        return (a.detach().cpu().double() - b_.detach()).abs().max().item() / (b_.detach().abs().max().item() + 1e-12)
    assert rel(yg, yr) < min(tol, 2e-3)
    assert abs(float(pg.detach()) - float(pr.detach())) < tol * abs(float(pr.detach()))
    for a, b_ in zip(gg, gr):
        assert (a is None) == (b_ is None)
        if a is not None:
            assert rel(a, b_) < tol, (mode, rel(a, b_))
