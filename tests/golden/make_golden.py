#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference).

Runs only in the authoring container (the reference tree does not exist on the GPU box); the
fixtures it writes are committed.  Nothing here is imported by the product.

The reference is imported as is, with the two import-time monkey patches SURVEY.md section 4 lists
(`layers.math = math`; `ConvLayer(activate=False)` -> no activation) -- the reference files are
not edited.  Its JIT-built extensions go to a scratch TORCH_EXTENSIONS_DIR.

Parameters are NOT stored (a 512-channel generator is ~100 MB): every module's state_dict is
filled by `det_fill`, a deterministic function of (seed, key name, shape), which the tests replay.
"""
import math
import os
import sys
import zlib

os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/sr_ref_ext")
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")

import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STYLERENDERER_REFERENCE", "/root/reference")


def det_fill(module, seed, scale_bias=0.1):
    """Fill every parameter/buffer deterministically from (seed, name).  Biases / noise weights get
    small non-zero values (they initialise to 0 and would hide bugs)."""
    sd = module.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            if name.endswith("kernel"):      # FIR taps are structural, keep
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            r = torch.randn(t.shape, generator=g, dtype=torch.float32)
            if name.endswith("modulation.bias"):
                r = 1.0 + 0.1 * r
            elif name.endswith("bias") or name.endswith("noise.weight"):
                r = scale_bias * r
            t.copy_(r.to(t.dtype))
    module.load_state_dict(sd)
    return module


def seeded(shape, seed, dtype=torch.float32):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32).to(dtype)


def grid_mesh(n, b, seed, jitter=0.03, dtype=torch.float32):
    """n x n vertex grid over [-0.9,0.9]^2 with a gaussian bump, CCW in y-up NDC."""
    lin = torch.linspace(-0.9, 0.9, n)
    ys, xs = torch.meshgrid(lin, lin, indexing="ij")
    base = torch.stack([xs, ys, 0.5 * torch.exp(-2 * (xs ** 2 + ys ** 2))], -1).view(-1, 3)
    v = base[None] + jitter * seeded((b, n * n, 3), seed)
    idx = torch.arange(n * n).view(n, n)
    a, bb, c, d = (idx[:-1, :-1].reshape(-1), idx[:-1, 1:].reshape(-1), idx[1:, :-1].reshape(-1),
                   idx[1:, 1:].reshape(-1))
    tri = torch.cat([torch.stack([a, bb, c], 1), torch.stack([bb, d, c], 1)], 0)
    return v.to(dtype).contiguous(), tri.contiguous()


def import_reference():
    sys.path.insert(0, REF)
    import layers  # noqa
    import model  # noqa
    import op  # noqa
    layers.math = math
    orig = layers.ConvLayer.__init__

    def patched(self, *a, **k):
        if len(a) >= 7 and a[6] is False:
            a = a[:6] + ("none",) + a[7:]
        if k.get("activate", "lrelu") is False:
            k["activate"] = "none"
        orig(self, *a, **k)

    layers.ConvLayer.__init__ = patched
    return layers, model, op


def main():
    layers, model, op = import_reference()
    out = {}

    # ---- upfirdn2d_native (op/upfirdn2d.py:159-200); case 0 is BASELINE.json configs[0]
    k4 = layers.make_kernel([1, 3, 3, 1])
    kasym = seeded((4, 4), 77)
    k3 = seeded((3, 3), 78)
    cases = [
        ("blur_cfg1", (1, 3, 64, 64), k4, 1, 1, (2, 1)),
        ("blur_after_upconv", (2, 5, 17, 17), k4 * 4, 1, 1, (1, 1)),
        ("skip_upsample", (2, 3, 16, 16), k4 * 4, 2, 1, (2, 1)),
        ("downsample", (2, 3, 16, 16), k4, 1, 2, (1, 1)),
        ("d_blur_22", (1, 4, 12, 12), k4, 1, 1, (2, 2)),
        ("asym_up2", (1, 2, 9, 7), kasym, 2, 1, (2, 1)),
        ("asym_down2", (1, 2, 11, 13), kasym, 1, 2, (2, 2)),
        ("k3_plain", (1, 2, 8, 8), k3, 1, 1, (1, 1)),
        ("neg_pad", (1, 2, 10, 10), kasym, 1, 1, (-1, 2)),
        ("up2_down2", (1, 2, 8, 8), kasym, 2, 2, (1, 2)),
    ]
    up = {}
    for i, (name, shape, k, u, d, pad) in enumerate(cases):
        x = seeded(shape, 100 + i)
        up[name] = dict(x=x, k=k.clone(), up=u, down=d, pad=pad, y=op.upfirdn2d(x, k, up=u, down=d, pad=pad))
    out["upfirdn2d"] = up

    # ---- fused_leaky_relu CPU branch (op/fused_act.py:87-94) and the C++ CPU kernel via oracle/_ref
    fl = {}
    for i, shape in enumerate([(2, 5, 7, 9), (4, 16), (1, 3, 1, 1)]):
        x = seeded(shape, 200 + i)
        b = seeded((shape[1],), 210 + i)
        fl["case%d" % i] = dict(x=x, b=b, y=op.fused_leaky_relu(x, b))
    out["fused_leaky_relu"] = fl

    # ---- rasterizer: the reference's only known-answer test (op/rasterize.py:83-107) + seeded meshes
    v = torch.tensor([[[-1, -1, 0], [-1, 1, 0], [1, 0, 0]]], dtype=torch.float64)
    f = torch.tensor([[2, 1, 0]])
    t = torch.tensor([[[1, 0], [0, 1], [0, 0]]], dtype=torch.float64)
    v.requires_grad_(True)
    t.requires_grad_(True)
    o = op.rasterize(v, t, f, 5)
    go = seeded(o.shape, 5, torch.float64)
    gv, gt = torch.autograd.grad(o, (v, t), go)
    ras = {"selftest": dict(v=v.detach(), t=t.detach(), f=f, h=5, out=o.detach(), go=go, gv=gv, gt=gt)}
    for name, (n, b, h, dtype) in {"grid24_h32_f32": (24, 2, 32, torch.float32),
                                   "grid24_h8_f32": (24, 2, 8, torch.float32),
                                   "grid16_h16_f64": (16, 1, 16, torch.float64)}.items():
        vv, tri = grid_mesh(n, b, 300 + h, dtype=dtype)
        tex = seeded((b, n * n, 3), 310 + h, dtype)
        vv.requires_grad_(True)
        tex.requires_grad_(True)
        o = op.rasterize(vv, tex, tri, h)
        go = seeded(o.shape, 320 + h, dtype)
        gv, gt = torch.autograd.grad(o, (vv, tex), go)
        ind, coeff = op.rasterize.__globals__["rasterize_op"].forward(vv.detach(), tri, h, 0, False, 1e-6)
        ras[name] = dict(n=n, b=b, h=h, seed=300 + h, tex=tex.detach(), go=go, out=o.detach(),
                         ind=ind.to(torch.int32), coeff=coeff, gv=gv, gt=gt)
    out["rasterize"] = ras

    # ---- module level (layers.py / model.py), parameters via det_fill
    mods = {}
    style = seeded((3, 32), 400)

    def run(mod, args, wrt):
        y = mod(*args)
        gy = seeded(y.shape, 999)
        params = [p for _, p in sorted(mod.named_parameters())]
        grads = torch.autograd.grad(y, wrt + params, gy, allow_unused=True)
        return y.detach(), gy, [g.detach() if g is not None else None for g in grads[:len(wrt)]], \
            {n: (g.detach() if g is not None else None)
             for (n, _), g in zip(sorted(mod.named_parameters()), grads[len(wrt):])}

    for name, ctor, kw, res in [
        ("modconv_plain", layers.ModulatedConv2d, dict(in_channel=8, out_channel=12, kernel_size=3, style_dim=32), 9),
        ("modconv_up", layers.ModulatedConv2d, dict(in_channel=8, out_channel=6, kernel_size=3, style_dim=32,
                                                    upsample=True), 8),
        ("modconv_1x1_nodemod", layers.ModulatedConv2d, dict(in_channel=8, out_channel=3, kernel_size=1, style_dim=32,
                                                             demodulate=False), 8),
    ]:
        m = det_fill(ctor(**kw), 500)
        x = seeded((3, 8, res, res), 401).requires_grad_(True)
        s = style.clone().requires_grad_(True)
        y, gy, (gx, gs), gp = run(m, (x, s), [x, s])
        mods[name] = dict(kw=kw, x=x.detach(), style=style, y=y, gy=gy, gx=gx, gs=gs, gp=gp)
    for name, up in [("styledconv_plain", False), ("styledconv_up", True)]:
        m = det_fill(model.StyledConv(8, 12, 3, 32, upsample=up), 501)
        x = seeded((3, 8, 8, 8), 402).requires_grad_(True)
        s = style.clone().requires_grad_(True)
        r = 16 if up else 8
        noise = seeded((3, 1, r, r), 403)
        y, gy, (gx, gs), gp = run(m, (x, s, noise), [x, s])
        mods[name] = dict(x=x.detach(), style=style, noise=noise, y=y, gy=gy, gx=gx, gs=gs, gp=gp)
    m = det_fill(model.StyledMapConv(8, 12, 3, 32), 502)
    x = seeded((3, 8, 8, 8), 404).requires_grad_(True)
    s = style.clone().requires_grad_(True)
    smap = seeded((3, 2, 8, 8), 405).requires_grad_(True)
    noise = seeded((3, 1, 8, 8), 406)
    y, gy, (gx, gs, gm), gp = run(m, (x, s, smap, noise), [x, s, smap])
    mods["styledmapconv"] = dict(x=x.detach(), style=style, stylemap=smap.detach(), noise=noise, y=y, gy=gy, gx=gx,
                                 gs=gs, gm=gm, gp=gp)
    m = det_fill(model.ToRGB(8, 32), 503)
    x = seeded((3, 8, 8, 8), 407).requires_grad_(True)
    s = style.clone().requires_grad_(True)
    skip = seeded((3, 3, 4, 4), 408).requires_grad_(True)
    y, gy, (gx, gs, gk), gp = run(m, (x, s, skip), [x, s, skip])
    mods["torgb"] = dict(x=x.detach(), style=style, skip=skip.detach(), y=y, gy=gy, gx=gx, gs=gs, gk=gk, gp=gp)
    out["modules"] = mods

    # ---- networks: Generator(32), GeneratorWithMap(16), Discriminator(16); params via det_fill
    nets = {}
    g = det_fill(model.Generator(32, 64, 2), 600).eval()
    z = seeded((2, 64), 601).requires_grad_(True)
    img, _ = g([z], randomize_noise=False)
    gimg = seeded(img.shape, 602)
    gz, gw = torch.autograd.grad(img, (z, g.convs[3].conv.weight), gimg)
    nets["generator32"] = dict(z=z.detach(), img=img.detach(), gimg=gimg, gz=gz, gw_convs3_norm=gw.norm(),
                               gw_convs3_slice=gw[0, :4, :4].clone(), n_keys=len(g.state_dict()))
    gm = det_fill(model.GeneratorWithMap(16, 64, 2), 610).eval()
    vv, tri = grid_mesh(24, 2, 611)
    tex = torch.nn.functional.normalize(seeded((2, 24 * 24, 3), 612), dim=-1)
    z = seeded((2, 64), 613)
    img, _, normals = gm([z], (vv, tex, tri), return_normals=True, randomize_noise=False)
    nets["generatorwithmap16"] = dict(z=z, tex=tex, img=img.detach(), normal16=normals[-1].detach(),
                                      n_keys=len(gm.state_dict()))
    d = det_fill(model.Discriminator(16), 620).eval()
    x = seeded((4, 3, 16, 16), 621)
    nets["discriminator16"] = dict(x=x, y=d(x).detach())
    out["networks"] = nets

    torch.save(out, os.path.join(HERE, "reference_golden.pt"))
    print("wrote", os.path.join(HERE, "reference_golden.pt"),
          os.path.getsize(os.path.join(HERE, "reference_golden.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
