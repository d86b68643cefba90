"""bf16 operand mode (tc_conv.precision("bf16"); BASELINE.json configs[3]: the GAR train step with bf16 convolutions).

GEMM operands are bfloat16 (8-bit significands), everything else fp32: the bar here is the bf16-appropriate one SURVEY.md
section 7 ("Hard parts") asks for -- outputs within 2e-2 (max-norm relative) of the fp32 oracle / reference fixtures,
gradients finite and inside the leaky-ReLU envelope.  The kernels themselves are checked to accumulation order in
tests/test_gpu_conv.py (every test there runs in both operand modes against float64 convolutions of the same rounded
operands)."""
import pytest
import torch

from make_golden import det_fill, grid_mesh, seeded
from parity_util import REPORT, hold, rel_err, tcgen05

pytestmark = pytest.mark.gpu

OUT_TOL = 2e-2


@pytest.fixture(scope="module")
def golden_tc():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return torch.load(os.path.join(root, "tests", "golden", "reference_golden_tc.pt"), weights_only=False)


def envelope(key, errs, med_tol, max_tol):
    import statistics
    vals = sorted(e for e, _ in errs)
    assert all(v == v and v != float("inf") for v in vals), f"{key}: non-finite gradient"
    med, worst = statistics.median(vals), max(errs)
    REPORT[key + "/gradients"] = {"tensors": len(vals), "median": med, "max": worst[0], "argmax": worst[1]}
    print(f"{key}: {len(vals)} gradient tensors, median {med:.2e}, max {worst[0]:.2e} ({worst[1]})")
    assert med <= med_tol and worst[0] <= max_tol, (key, med, worst)


@pytest.mark.parametrize("name", ["modconv_plain_128", "modconv_plain_128_256", "modconv_up_128", "modconv_up_256_128"])
def test_modulated_conv_bf16_vs_reference_fixture(golden_tc, name):
    """ModulatedConv2d (no activation inside): output and every gradient within the bf16 bound of the reference's numbers."""
    from stylerenderer_b200 import layers as L
    g = golden_tc["modules"][name]
    m = det_fill(L.ModulatedConv2d(**g["kw"]), 1500).cuda()
    x = g["x"].cuda().requires_grad_(True)
    s = g["style"].cuda().requires_grad_(True)
    with tcgen05("bf16") as t:
        y = m(x, s)
        names = [n for n, _ in sorted(m.named_parameters())]
        gr = torch.autograd.grad(y, [x, s] + [p for _, p in sorted(m.named_parameters())], g["gy"].cuda())
    assert t.calls >= 3
    k = f"{name}[bf16]"
    hold(k + "/y", y, g["y"], OUT_TOL)
    hold(k + "/gx", gr[0], g["gx"], OUT_TOL)
    hold(k + "/gs", gr[1], g["gs"], OUT_TOL)
    for n, g_ in zip(names, gr[2:]):
        w = g["gp"][n]
        if isinstance(w, dict):
            idx = tuple(slice(0, s_) for s_ in w["slice"].shape)
            scale = float(w["norm"]) / (g_.numel() ** 0.5)
            e = float((g_[idx].cpu().double() - w["slice"].double()).abs().max()) / max(float(w["slice"].abs().max()), scale)
            assert e <= OUT_TOL, (n, e)
        else:
            hold(f"{k}/{n}", g_, w, OUT_TOL)


def test_generator64_chain_bf16_vs_reference_fixture(golden_tc):
    """Generator(64) on the chained blocks with bfloat16 operands handed from epilogue to GEMM (fused._proxy carries the
    autograd edge): image within 2e-2 of the reference, every gradient finite and in the envelope."""
    from stylerenderer_b200 import model as M
    g = golden_tc["networks"]["generator64"]
    G = det_fill(M.Generator(64, 64, 2), 1600).cuda().eval()
    z = g["z"].cuda().requires_grad_(True)
    with tcgen05("bf16") as t:
        img, _ = G([z], randomize_noise=False)
        names = [n for n, _ in sorted(G.named_parameters())]
        gr = torch.autograd.grad(img, [z] + [p for _, p in sorted(G.named_parameters())], g["gimg"].cuda(), allow_unused=True)
    assert t.calls >= 20
    hold("generator64[bf16]/img", img, g["img"], OUT_TOL)
    from parity_util import param_errors
    errs = [(rel_err(gr[0], g["gz"]), "z")] + param_errors("generator64[bf16]", dict(zip(names, gr[1:])), g["gp"])
    envelope("generator64[bf16]", errs, 1e-1, 1.0)


def test_train_step_networks_bf16_vs_tf32():
    """GeneratorWithMap(64) + Discriminator(64) forward / backward in bf16 operand mode against the shipped tf32 mode of the
    same package (itself held to the reference elsewhere): images / logits within 2e-2, all gradients finite."""
    from stylerenderer_b200 import model as M
    G = det_fill(M.GeneratorWithMap(64, 64, 2), 1900).cuda().eval()
    D = det_fill(M.Discriminator(64), 1901).cuda().to(memory_format=torch.channels_last)
    v, tri = grid_mesh(24, 4, 1902)
    tex = torch.nn.functional.normalize(seeded((4, 576, 3), 1903), dim=-1)
    z = seeded((4, 64), 1904).cuda()

    def run(mode):
        with tcgen05(mode):
            zz = z.clone().requires_grad_(True)
            img, _, _ = G([zz], (v.cuda(), tex.cuda(), tri.cuda()), randomize_noise=False)
            logits = D(img.contiguous(memory_format=torch.channels_last))
            ps = [p for _, p in sorted(G.named_parameters()) if p.requires_grad] + [p for _, p in sorted(D.named_parameters())]
            gr = torch.autograd.grad(logits.sum(), [zz] + ps, allow_unused=True)
        return img.detach(), logits.detach(), gr
    img_a, log_a, gr_a = run("tf32")
    img_b, log_b, gr_b = run("bf16")
    hold("train_nets[bf16 vs tf32]/img", img_b, img_a, OUT_TOL)
    hold("train_nets[bf16 vs tf32]/logits", log_b, log_a, 5e-2)
    errs = [(rel_err(b_, a), str(i)) for i, (a, b_) in enumerate(zip(gr_a, gr_b)) if a is not None and float(a.abs().max()) > 0]
    envelope("train_nets[bf16 vs tf32]", errs, 1.5e-1, 2.0)
