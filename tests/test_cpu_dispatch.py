"""CPU-tensor paths of the drop-in API ("CPU tensors must keep working").

The reference's `op.upfirdn2d` and `op.fused_leaky_relu` dispatch on the device: CPU tensors go to plain torch ops
(`upfirdn2d_native`, reference op/upfirdn2d.py:146-200; the `F.leaky_relu` branch, op/fused_act.py:87-94), and callers rely
on it (reference utils_face.py:515-517; BASELINE.json configs[0] is exactly that path).  The package mirrors the dispatch
with its own plain-torch formulation -- no import of oracle/ -- checked here against the fixtures the real reference
produced (tests/golden/make_golden.py)."""
import pytest
import torch

from stylerenderer_b200 import layers as L, op


@pytest.mark.parametrize("name", ["blur_cfg1", "blur_after_upconv", "skip_upsample", "downsample", "d_blur_22", "asym_up2",
                                  "asym_down2", "k3_plain", "neg_pad", "up2_down2"])
def test_upfirdn2d_cpu_matches_reference(golden, name):
    g = golden["upfirdn2d"][name]                        # blur_cfg1 = BASELINE.json configs[0]
    y = op.upfirdn2d(g["x"], g["k"], up=g["up"], down=g["down"], pad=g["pad"])
    assert y.shape == g["y"].shape
    torch.testing.assert_close(y, g["y"], rtol=1e-6, atol=1e-6)


def test_config0_through_the_module_api(golden):
    """BASELINE.json configs[0]: 4x4 blur on 1x3x64x64 through the public API on the CPU, incl. the Blur module."""
    g = golden["upfirdn2d"]["blur_cfg1"]
    blur = L.Blur([1, 3, 3, 1], pad=(2, 1))
    torch.testing.assert_close(blur(g["x"]), g["y"], rtol=1e-6, atol=1e-6)


def test_upfirdn2d_cpu_is_twice_differentiable():
    x = torch.randn(1, 2, 6, 5, dtype=torch.float64, requires_grad=True)
    k = torch.randn(4, 4, dtype=torch.float64)
    f = lambda t: op.upfirdn2d(t, k, up=2, down=1, pad=(2, 1))          # noqa: E731
    assert torch.autograd.gradcheck(f, (x,))
    assert torch.autograd.gradgradcheck(f, (x,))


@pytest.mark.parametrize("name", ["case0", "case1", "case2"])
def test_fused_leaky_relu_cpu_matches_reference(golden, name):
    g = golden["fused_leaky_relu"][name]
    assert torch.equal(op.fused_leaky_relu(g["x"], g["b"]), g["y"])


def test_fused_leaky_relu_cpu_slope_quirk_and_module():
    """The reference's CPU branch ignores `negative_slope` (op/fused_act.py:91 hard-codes 0.2); reproduced."""
    x, b = torch.randn(2, 3, 4, 4), torch.randn(3)
    assert torch.equal(op.fused_leaky_relu(x, b, negative_slope=0.5), op.fused_leaky_relu(x, b, negative_slope=0.2))
    m = op.FusedLeakyReLU(3)
    with torch.no_grad():
        m.bias.copy_(b)
    y = m(x.requires_grad_(True))
    y.sum().backward()
    assert m.bias.grad is not None and x.grad is not None


def test_tensor_core_backend_leaves_cpu_tensors_on_the_composed_path():
    """With conv_backend = "tcgen05" selected (as a GPU job would), modules fed CPU tensors must take the composed torch ops:
    every fused dispatcher of the package (style-map ResBlock kernel, small-channel conv pair, Discriminator stem, blur ->
    operand, residual combine, tensor-core ConvLayers) gates on `is_cuda`.  Same outputs and gradients as the "cudnn" backend."""
    from stylerenderer_b200 import model as M
    torch.manual_seed(5)
    nets = [L.ResBlock(3, 4, downsample=False), L.ConvLayer(3, 128, 1), L.ResBlock(128, 128), M.Discriminator(16)]
    xs = [torch.randn(2, 3, 9, 9), torch.randn(2, 3, 8, 8), torch.randn(1, 128, 8, 8), torch.randn(2, 3, 16, 16)]
    old = L.get_conv_backend()
    try:
        for net, x in zip(nets, xs):
            res = {}
            for backend in ("cudnn", "tcgen05"):
                L.set_conv_backend(backend)
                net.zero_grad()
                xx = x.clone().requires_grad_(True)
                y = net(xx)
                y.square().mean().backward()
                res[backend] = (y.detach(), xx.grad, [p.grad.clone() for p in net.parameters()])
            torch.testing.assert_close(res["tcgen05"][0], res["cudnn"][0], rtol=0, atol=0)
            torch.testing.assert_close(res["tcgen05"][1], res["cudnn"][1], rtol=0, atol=0)
            for a, b in zip(res["tcgen05"][2], res["cudnn"][2]):
                torch.testing.assert_close(a, b, rtol=0, atol=0)
    finally:
        L.set_conv_backend(old)
