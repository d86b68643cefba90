#!/usr/bin/env python
"""Generate tests/golden/reference_golden_tc.pt: fixtures at channel counts the tcgen05 path takes (multiples of 128),
produced by running the UNMODIFIED reference (imported from /root/reference) on the CPU in fp32.

The small fixtures of make_golden.py use 8 -> 12 channel modules, which the tensor-core kernels do not accept
(tc_conv.supported: cin % 32, cout % 128), so `conv_backend = "tcgen05"` silently fell through to the composed path
there.  These cases make the tensor-core kernels themselves meet reference-generated numbers:
  ModulatedConv2d (reference layers.py:259-323) plain / up-sampling, 128 -> 128 and 128 -> 256 channels;
  StyledConv / StyledMapConv (reference model.py:11-55) plain / up-sampling, 128 channels;
  Generator(64) / GeneratorWithMap(32) / Discriminator(32) (reference model.py:71-336; 512-channel layers).
Parameters are replayed by det_fill (not stored); large weight gradients are stored as a corner slice + their norm.
Authoring container only; the fixture is committed."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import det_fill, grid_mesh, import_reference, seeded  # noqa: E402


def compress(name, g):
    """Larger gradients (>= 2048 elements) -> (corner slice, norm); everything else as is."""
    if g is None:
        return None
    if g.numel() >= 2048:
        idx = tuple(slice(0, min(8, d)) for d in g.shape)
        return {"slice": g[idx].clone(), "norm": g.double().norm().float()}
    return g.clone()


def run(mod, args, wrt, seed=999):
    y = mod(*args)
    gy = seeded(y.shape, seed)
    names = [n for n, _ in sorted(mod.named_parameters())]
    params = [p for _, p in sorted(mod.named_parameters())]
    grads = torch.autograd.grad(y, wrt + params, gy, allow_unused=True)
    gin = [g.detach() if g is not None else None for g in grads[:len(wrt)]]
    gp = {n: compress(n, g.detach() if g is not None else None) for n, g in zip(names, grads[len(wrt):])}
    return y.detach(), gy, gin, gp


def main():
    layers, model, op = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    mods = {}
    b, sd = 2, 64
    style = seeded((b, sd), 1400)
    for name, kw, res in [
        ("modconv_plain_128", dict(in_channel=128, out_channel=128, kernel_size=3, style_dim=sd), 8),
        ("modconv_plain_128_256", dict(in_channel=128, out_channel=256, kernel_size=3, style_dim=sd), 8),
        ("modconv_up_128", dict(in_channel=128, out_channel=128, kernel_size=3, style_dim=sd, upsample=True), 8),
        ("modconv_up_256_128", dict(in_channel=256, out_channel=128, kernel_size=3, style_dim=sd, upsample=True), 4),
    ]:
        m = det_fill(layers.ModulatedConv2d(**kw), 1500)
        x = seeded((b, kw["in_channel"], res, res), 1401).requires_grad_(True)
        s = style.clone().requires_grad_(True)
        y, gy, (gx, gs), gp = run(m, (x, s), [x, s])
        mods[name] = dict(kw=kw, x=x.detach(), style=style, y=y, gy=gy, gx=gx, gs=gs, gp=gp)
    for name, up in [("styledconv_plain_128", False), ("styledconv_up_128", True)]:
        m = det_fill(model.StyledConv(128, 128, 3, sd, upsample=up), 1501)
        x = seeded((b, 128, 8, 8), 1402).requires_grad_(True)
        s = style.clone().requires_grad_(True)
        r = 16 if up else 8
        noise = seeded((b, 1, r, r), 1403)
        y, gy, (gx, gs), gp = run(m, (x, s, noise), [x, s])
        mods[name] = dict(up=up, x=x.detach(), style=style, noise=noise, y=y, gy=gy, gx=gx, gs=gs, gp=gp)
    for name, up in [("styledmapconv_plain_128", False), ("styledmapconv_up_128", True)]:
        m = det_fill(model.StyledMapConv(128, 128, 3, sd, upsample=up), 1502)
        x = seeded((b, 128, 8, 8), 1404).requires_grad_(True)
        s = style.clone().requires_grad_(True)
        r = 16 if up else 8
        smap = seeded((b, 2, r, r), 1405)
        smap[:, :, : r // 2] = 0                       # half of the map exactly 0: background of a rasterised normal map
        smap.requires_grad_(True)
        noise = seeded((b, 1, r, r), 1406)
        y, gy, (gx, gs, gm), gp = run(m, (x, s, smap, noise), [x, s, smap])
        mods[name] = dict(up=up, x=x.detach(), style=style, stylemap=smap.detach(), noise=noise, y=y, gy=gy, gx=gx, gs=gs,
                          gm=gm, gp=gp)
    out["modules"] = mods

    nets = {}
    g = det_fill(model.Generator(64, 64, 2), 1600).eval()
    z = seeded((2, 64), 1601).requires_grad_(True)
    img, _ = g([z], randomize_noise=False)
    gimg = seeded(img.shape, 1602)
    names = [n for n, p in sorted(g.named_parameters())]
    grads = torch.autograd.grad(img, [z] + [p for _, p in sorted(g.named_parameters())], gimg, allow_unused=True)
    nets["generator64"] = dict(z=z.detach(), img=img.detach(), gimg=gimg, gz=grads[0].detach(),
                               gp={n: compress(n, gr.detach() if gr is not None else None) for n, gr in zip(names, grads[1:])},
                               n_keys=len(g.state_dict()))
    gm = det_fill(model.GeneratorWithMap(32, 64, 2), 1610).eval()
    vv, tri = grid_mesh(24, 2, 1611)
    tex = torch.nn.functional.normalize(seeded((2, 24 * 24, 3), 1612), dim=-1)
    z = seeded((2, 64), 1613).requires_grad_(True)
    vv.requires_grad_(True)
    tex.requires_grad_(True)
    img, _, normals = gm([z], (vv, tex, tri), return_normals=True, randomize_noise=False)
    gimg = seeded(img.shape, 1614)
    names = [n for n, p in sorted(gm.named_parameters())]
    grads = torch.autograd.grad(img, [z, vv, tex] + [p for _, p in sorted(gm.named_parameters())], gimg, allow_unused=True)
    nets["generatorwithmap32"] = dict(z=z.detach(), tex=tex.detach(), img=img.detach(), normal32=normals[-1].detach(), gimg=gimg,
                                      gz=grads[0].detach(), gv=grads[1].detach(), gtex=grads[2].detach(),
                                      gp={n: compress(n, gr.detach() if gr is not None else None)
                                          for n, gr in zip(names, grads[3:])},
                                      n_keys=len(gm.state_dict()))
    d = det_fill(model.Discriminator(32), 1620).eval()
    x = seeded((4, 3, 32, 32), 1621).requires_grad_(True)
    y = d(x)
    names = [n for n, p in sorted(d.named_parameters())]
    grads = torch.autograd.grad(y.sum(), [x] + [p for _, p in sorted(d.named_parameters())])
    nets["discriminator32"] = dict(x=x.detach(), y=y.detach(), gx=grads[0].detach(),
                                   gp={n: compress(n, gr.detach()) for n, gr in zip(names, grads[1:])})
    out["networks"] = nets
    path = os.path.join(HERE, "reference_golden_tc.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
