"""CPU: the checkpoint contract -- product modules expose exactly the reference's state_dict keys and shapes
(pinned through oracle/torch_ref.py, whose key count is pinned to the real reference by the golden fixtures)."""
import torch

from oracle import torch_ref as T


def _sig(m):
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_generator_keys(golden):
    from stylerenderer_b200 import model as M
    a, b = _sig(M.Generator(32, 64, 2)), _sig(T.Generator(32, 64, 2))
    assert a == b
    assert len(a) == golden["networks"]["generator32"]["n_keys"]
    for key in ("convs.1.conv.weight", "convs.0.conv.blur.kernel", "convs.0.conv.modulation.bias", "convs.0.noise.weight",
                "convs.0.activate.bias", "to_rgbs.0.upsample.kernel", "to_rgbs.11.bias" if False else "to_rgbs.5.bias",
                "noises.noise_0", "style.1.weight", "input.input", "to_rgb1.conv.weight"):
        assert key in a, key


def test_generator_with_map_and_discriminator_keys(golden):
    from stylerenderer_b200 import model as M
    a, b = _sig(M.GeneratorWithMap(16, 64, 2)), _sig(T.GeneratorWithMap(16, 64, 2))
    assert a == b and len(a) == golden["networks"]["generatorwithmap16"]["n_keys"]
    assert any(k.startswith("norm1.") for k in a) and any(k.startswith("norm_to_style.0.") for k in a)
    assert _sig(M.Discriminator(16)) == _sig(T.Discriminator(16))


def test_reference_checkpoint_loads():
    from stylerenderer_b200 import model as M
    src = T.Generator(16, 32, 2)
    dst = M.Generator(16, 32, 2)
    missing, unexpected = dst.load_state_dict(src.state_dict(), strict=True)
    assert not missing and not unexpected
    assert torch.equal(dst.convs[0].conv.weight, src.convs[0].conv.weight)
