"""GPU parity of the tcgen05 implicit-GEMM convolution (csrc/modconv.cu) against float64 torch convolutions of the
same tf32-rounded operands (so the only difference left is fp32 accumulation order)."""
import pytest
import torch
import torch.nn.functional as F

from make_golden import seeded

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc():
    assert torch.cuda.is_available()
    from stylerenderer_b200 import tc_conv
    return tc_conv


@pytest.fixture(params=["tf32", "bf16"], autouse=True)
def operand_mode(request):
    """Every kernel-level test runs in both operand modes: tf32 (fp32 words, kind::tf32) and bf16 (kind::f16).  The
    references are float64 convolutions of the SAME rounded operands (tc.modulate / tc.weight_prep produce them in the
    active mode), so the bounds -- fp32 accumulation order only -- are the same in both."""
    from stylerenderer_b200 import tc_conv
    with tc_conv.precision(request.param):
        yield request.param


def skip_unless_supported(mode, cin):
    if mode == "bf16" and cin % 64:
        pytest.skip("bf16 operands: channels per K block = 64")


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


def relerr(got, want):
    return float((got.double() - want).abs().max() / want.abs().max())


CASES = [(2, 16, 16, 64, 128), (3, 8, 8, 32, 256), (9, 4, 4, 96, 128), (1, 24, 40, 32, 384), (2, 64, 64, 128, 128),
         (5, 5, 5, 32, 128), (2, 9, 9, 64, 256), (1, 17, 33, 32, 128),
         # large enough for the halo CTA-pair kernel (8 x 16 tiles, one activation box per K block)
         (8, 32, 32, 64, 128), (4, 40, 56, 32, 256), (16, 33, 17, 32, 128), (8, 64, 64, 32, 256), (6, 32, 48, 128, 384)]


@pytest.mark.parametrize("case", CASES)
def test_plain_conv_and_dgrad(tc, case, operand_mode):
    b, h, w, cin, cout = case
    skip_unless_supported(operand_mode, cin)
    x = tc.modulate(nhwc(seeded((b, cin, h, w), 1)).cuda())               # tf32-rounded operand
    wt = seeded((cout, cin, 3, 3), 2).cuda()
    wm = tc.weight_prep(wt, 0.05, 0)
    w_rounded = wm.view(cout, 3, 3, cin).permute(0, 3, 1, 2)              # == tf32(wt * 0.05)
    y = tc.conv3x3(x, wm)
    want = F.conv2d(nchw(x).double(), w_rounded.double(), padding=1)
    assert relerr(nchw(y), want) < 2e-5, case
    if cin % 128 == 0 and cout % 32 == 0:
        g = tc.modulate(nhwc(seeded((b, cout, h, w), 3)).cuda())
        wd = tc.weight_prep(wt, 0.05, 1)
        dx = tc.conv3x3(g, wd)
        want = torch.autograd.grad(F.conv2d(nchw(x).double().requires_grad_(True), w_rounded.double(), padding=1).sum() * 0
                                   + 0, [], allow_unused=True) if False else None
        xin = nchw(x).double().requires_grad_(True)
        ref, = torch.autograd.grad(F.conv2d(xin, w_rounded.double(), padding=1), xin, nchw(g).double())
        assert relerr(nchw(dx), ref) < 2e-5, case


@pytest.mark.parametrize("shape", [(3, 16, 16, 64, 128), (8, 32, 32, 64, 128), (5, 48, 40, 32, 256)])
def test_styled_epilogue(tc, shape, operand_mode):
    b, h, w, cin, cout = shape
    skip_unless_supported(operand_mode, cin)
    x = tc.modulate(nhwc(seeded((b, cin, h, w), 4)).cuda())
    wt = seeded((cout, cin, 3, 3), 5).cuda()
    wm = tc.weight_prep(wt, 0.04, 0)
    w_rounded = wm.view(cout, 3, 3, cin).permute(0, 3, 1, 2).double()
    d = (seeded((b, cout), 6).abs() + 0.5).cuda()
    s2 = (seeded((b, cout), 7) + 1).cuda()
    bias = seeded((cout,), 8).cuda()
    noise = seeded((b, 1, h, w), 9).cuda()
    nw = torch.tensor([0.3], device="cuda")
    smap = seeded((b, 4, h, w), 10).cuda()[:, 2:]                          # non-contiguous batch stride
    out = torch.empty(b, h, w, cout, device="cuda")
    out2 = tc.operand_like(out)
    tc.conv3x3(x, wm, out=out, epilogue=1, rowscale=d, out2=out2, scale2=s2, bias=bias, noise=noise.view(b, h, w),
               noise_weight=nw, stylemap=smap)
    conv = F.conv2d(nchw(x).double(), w_rounded, padding=1) * d.double().view(b, cout, 1, 1)
    t = conv * smap[:, :1].double() + smap[:, 1:2].double() + 0.3 * noise.double() + bias.double().view(1, -1, 1, 1)
    want = F.leaky_relu(t, 0.2) * 2 ** 0.5
    assert relerr(nchw(out), want) < 3e-5
    # the second output carries one rounding to the operand type: 2^-11 (tf32) / 2^-8 (bf16)
    assert relerr(nchw(out2), want * s2.double().view(b, cout, 1, 1)) < (6e-4 if operand_mode == "tf32" else 5e-3)
    # broadcast noise plane, no stylemap, no second output
    out3 = torch.empty_like(out)
    tc.conv3x3(x, wm, out=out3, epilogue=1, rowscale=d, bias=bias, noise=noise[0, 0].contiguous(), noise_weight=nw)
    want = F.leaky_relu(conv + 0.3 * noise[0, 0].double() + bias.double().view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert relerr(nchw(out3), want) < 3e-5


@pytest.mark.parametrize("case", [(2, 4, 4, 64, 128), (2, 8, 8, 32, 128), (1, 16, 16, 32, 256), (3, 32, 32, 64, 128),
                                  (1, 7, 12, 32, 128), (8, 32, 32, 128, 128), (6, 24, 40, 128, 256), (16, 16, 16, 128, 128)])
def test_transposed_conv_and_its_dgrad(tc, case, operand_mode):
    b, h, w, cin, cout = case
    skip_unless_supported(operand_mode, cin)
    x = tc.modulate(nhwc(seeded((b, cin, h, w), 11)).cuda())
    wt = seeded((cout, cin, 3, 3), 12).cuda()
    wm = tc.weight_prep(wt, 0.05, 0)
    w_rounded = wm.view(cout, 3, 3, cin).permute(0, 3, 1, 2).double()      # [cout, cin, 3, 3]
    d = (seeded((b, cout), 13).abs() + 0.5).cuda()
    y = tc.conv_transpose3x3_s2(x, wm, rowscale=d)
    assert y.shape == (b, 2 * h + 1, 2 * w + 1, cout)
    xin = nchw(x).double().requires_grad_(True)
    ref = F.conv_transpose2d(xin, w_rounded.transpose(0, 1), stride=2)
    assert relerr(nchw(y), ref.detach() * d.double().view(b, cout, 1, 1)) < 2e-5, case
    if cout % 32 == 0 and cin % 128 == 0:
        g = tc.modulate(nhwc(seeded((b, cout, 2 * h + 1, 2 * w + 1), 14)).cuda())
        wg = tc.weight_prep(wt, 0.05, 2)
        dx = tc.conv3x3_s2_gather(g, wg, (h, w))
        want, = torch.autograd.grad(ref, xin, nchw(g).double())
        assert relerr(nchw(dx), want) < 2e-5, case


def test_transposed_conv_dgrad_wide(tc):
    b, h, w, cin, cout = 2, 16, 16, 128, 128
    wt = seeded((cout, cin, 3, 3), 15).cuda()
    g = tc.modulate(nhwc(seeded((b, cout, 2 * h + 1, 2 * w + 1), 16)).cuda())
    wg = tc.weight_prep(wt, 0.05, 2)
    w_rounded = wg.view(cin, 3, 3, cout).permute(3, 0, 1, 2).double()      # back to [cout, cin, 3, 3]
    dx = tc.conv3x3_s2_gather(g, wg, (h, w))
    xin = torch.zeros(b, cin, h, w, dtype=torch.float64, device="cuda", requires_grad=True)
    want, = torch.autograd.grad(F.conv_transpose2d(xin, w_rounded.transpose(0, 1), stride=2), xin, nchw(g).double())
    assert relerr(nchw(dx), want) < 2e-5
    # plain dgrad at cin = cout = 128
    x = tc.modulate(nhwc(seeded((b, cin, h, w), 17)).cuda())
    gg = tc.modulate(nhwc(seeded((b, cout, h, w), 18)).cuda())
    wd = tc.weight_prep(wt, 0.05, 1)
    w_r = wd.view(cin, 3, 3, cout).flip(1, 2).permute(3, 0, 1, 2).double()
    xin = nchw(x).double().requires_grad_(True)
    want, = torch.autograd.grad(F.conv2d(xin, w_r, padding=1), xin, nchw(gg).double())
    assert relerr(nchw(tc.conv3x3(gg, wd)), want) < 2e-5


@pytest.mark.parametrize("case", [(2, 16, 16, 128, 128), (3, 8, 8, 128, 256), (9, 4, 4, 256, 128), (1, 20, 36, 128, 128),
                                  (2, 64, 64, 128, 128), (5, 5, 5, 128, 128),
                                  (2, 16, 16, 256, 256), (3, 8, 8, 512, 256), (5, 4, 4, 256, 512), (1, 32, 32, 256, 256)])   # CTA-pair kernel
def test_wgrad(tc, case):
    b, h, w, cin, cout = case
    x = tc.modulate(nhwc(seeded((b, cin, h, w), 21)).cuda())
    g = tc.modulate(nhwc(seeded((b, cout, h, w), 22)).cuda())
    dw = tc.wgrad3x3(g, x)                                                   # [cout, 9, cin]
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    want, = torch.autograd.grad(F.conv2d(nchw(x).double(), wt, padding=1), wt, nchw(g).double())
    got = dw.view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    assert relerr(got, want) < 5e-5, case
    # transposed stride-2 conv
    g2 = tc.modulate(nhwc(seeded((b, cout, 2 * h + 1, 2 * w + 1), 23)).cuda())
    dw2 = tc.wgrad_transpose3x3_s2(g2, x)
    wt2 = torch.zeros(cin, cout, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    want2, = torch.autograd.grad(F.conv_transpose2d(nchw(x).double(), wt2, stride=2), wt2, nchw(g2).double())
    got2 = dw2.view(cout, 3, 3, cin).permute(3, 0, 1, 2)                   # -> [cin, cout, 3, 3]
    assert relerr(got2, want2) < 5e-5, case


@pytest.mark.parametrize("shape", [(128, 128), (512, 256), (96, 160), (40, 3)])
def test_weight_prep_dual_matches_single_mode_kernels(tc, shape, operand_mode):
    """The one-pass weight preparation (forward layout, transposed layout, demodulation statistic) equals the per-mode
    kernels bit for bit and the torch definition of Wsq (reference layers.py:297)."""
    cout, cin = shape
    w = seeded((cout, cin, 3, 3), 41).cuda()
    for flip, mode in ((True, 1), (False, 2)):
        fwd, tr, wsq = tc.weight_prep_dual(w, 0.037, flip)
        if operand_mode == "bf16":      # the single-mode path is a torch cast of w * scale; the kernel rounds the same product
            torch.testing.assert_close(fwd.float(), tc.weight_prep(w, 0.037, 0).float(), rtol=8e-3, atol=0)
            torch.testing.assert_close(tr.float(), tc.weight_prep(w, 0.037, mode).float(), rtol=8e-3, atol=0)
            assert fwd.dtype == tr.dtype == torch.bfloat16
        else:
            assert torch.equal(fwd, tc.weight_prep(w, 0.037, 0))
            assert torch.equal(tr, tc.weight_prep(w, 0.037, mode))
        want = (w.double() * 0.037).pow(2).sum([2, 3])
        torch.testing.assert_close(wsq.double(), want, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("kind,shape", [("s1", (2, 128, 128, 24, 40)), ("s1", (3, 256, 128, 16, 16)),
                                        ("s2", (2, 128, 256, 33, 33)), ("s2", (1, 128, 128, 65, 37)),
                                        ("p2", (2, 128, 128, 31, 31)), ("p2", (3, 256, 128, 15, 33))])
def test_plain_conv_layer_tc(kind, shape, operand_mode):
    """EqualConv2d [+ bias + FusedLeakyReLU] of the Discriminator (reference layers.py:204-221, 341-378) on the tensor-core
    kernels: values and every gradient against float64 torch on the same tf32-rounded operands."""
    from stylerenderer_b200 import fused, tc_conv as tcm
    b, cin, cout, h, w = shape
    k, stride, pad = {"s1": (3, 1, 1), "s2": (3, 2, 0), "p2": (1, 2, 0)}[kind]
    # an input that is exact in the operand type (so the forward bound is accumulation order only)
    x = tcm.modulate(nhwc(seeded((b, cin, h, w), 51)).cuda()).float().permute(0, 3, 1, 2).requires_grad_(True)
    wt = seeded((cout, cin, k, k), 52).cuda().requires_grad_(True)
    bias = (seeded((cout,), 53).cuda() * 0.2).requires_grad_(True) if kind != "p2" else None
    scale = 1.0 / (cin * k * k) ** 0.5
    y = fused.PlainConvTC.apply(x, wt, bias, scale, kind, 0.2, 2 ** 0.5)
    gy = seeded(tuple(y.shape), 54).cuda()
    grads = torch.autograd.grad(y, [x, wt] + ([bias] if bias is not None else []), gy)
    w_r = tcm.weight_prep(wt.detach(), scale, 0).view(cout, k, k, cin).permute(0, 3, 1, 2).double().requires_grad_(True)
    x64 = x.detach().double().requires_grad_(True)
    b64 = bias.detach().double().requires_grad_(True) if bias is not None else None
    t = F.conv2d(x64, w_r, stride=stride, padding=pad)
    want = F.leaky_relu(t + b64.view(1, -1, 1, 1), 0.2) * 2 ** 0.5 if bias is not None else t
    assert relerr(y.detach(), want.detach()) < 3e-5, (kind, shape)
    # gradients: the backward operand (activation backward of gy) is rounded to the operand type inside the block
    gtol = 2e-3 if operand_mode == "tf32" else 2.5e-2
    wg = torch.autograd.grad(want, [x64, w_r] + ([b64] if bias is not None else []), gy.double())
    assert relerr(grads[0], wg[0]) < gtol, (kind, "dx")
    assert relerr(grads[1], wg[1] * scale) < gtol, (kind, "dw")       # d/dw = scale * d/d(scale*w)
    if bias is not None:
        # fp32 reduction of the masked gradient; a leaky-ReLU mask element that differs from the float64 evaluation (the
        # forward agrees to ~1e-5 only) moves it by up to a percent
        assert relerr(grads[2], wg[2]) < (1e-4 if operand_mode == "tf32" else 3e-2), (kind, "dbias")


@pytest.mark.parametrize("kind,shape", [("plain", (2, 128, 128, 12, 20)), ("up", (2, 128, 256, 9, 7)),
                                        ("down", (2, 128, 128, 17, 13)), ("down1", (2, 256, 128, 15, 9))])
def test_conv_tc_double_backward(kind, shape, operand_mode):
    """ConvTC / ConvDgradTC / ConvWgradTC (the twice-differentiable tensor-core convolution used on R1 / path-length
    iterations): first and second order gradients against torch's float64 convolution double backward."""
    from stylerenderer_b200 import fused
    b, cin, cout, h, w = shape
    x = (seeded((b, h, w, cin), 61).cuda()).requires_grad_(True)
    k = 1 if kind == "down1" else 3
    wt = (seeded((cout, cin, k, k), 62).cuda() * 0.05).requires_grad_(True)
    y = fused.ConvTC.apply(x, wt, kind)
    gy = seeded(tuple(y.shape), 63).cuda().requires_grad_(True)
    gx, gw = torch.autograd.grad(y, (x, wt), gy, create_graph=True)
    # a scalar that depends on the first-order gradients (like the path-length penalty) and its gradients
    c1, c2 = seeded(tuple(gx.shape), 64).cuda(), seeded(tuple(gw.shape), 65).cuda()
    pen = (gx * c1).sum() + (gw * c2).sum() + (gx ** 2).sum() * 0.01
    got = torch.autograd.grad(pen, (x, wt, gy))

    x64, w64, g64 = (t.detach().double().requires_grad_(True) for t in (x, wt, gy))
    xn = x64.permute(0, 3, 1, 2)
    if kind == "plain":
        yr = F.conv2d(xn, w64, padding=1)
    elif kind == "up":
        yr = F.conv_transpose2d(xn, w64.transpose(0, 1), stride=2)
    else:
        yr = F.conv2d(xn, w64, stride=2)
    tol = 2e-3 if operand_mode == "tf32" else 2.5e-2       # generic (unrounded) inputs: operand rounding 2^-11 / 2^-8
    assert relerr(y.detach().permute(0, 3, 1, 2), yr.detach()) < tol
    gxr, gwr = torch.autograd.grad(yr, (x64, w64), g64.permute(0, 3, 1, 2), create_graph=True)
    assert relerr(gx.detach(), gxr.detach()) < tol and relerr(gw.detach(), gwr.detach()) < tol
    penr = (gxr * c1.double()).sum() + (gwr * c2.double()).sum() + (gxr ** 2).sum() * 0.01
    want = torch.autograd.grad(penr, (x64, w64, g64))
    for name, a, bb in zip(("d/dx", "d/dw", "d/dgy"), got, want):
        assert relerr(a, bb) < 1.5 * tol, (kind, name, relerr(a, bb))
