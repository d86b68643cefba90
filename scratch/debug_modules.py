import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import det_fill, seeded
from oracle import torch_ref as T
from stylerenderer_b200 import layers as L, model as M
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
golden = torch.load(os.path.join(ROOT, "tests/golden/reference_golden.pt"), weights_only=False)

def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

def grads(mod, args, wrt, gy):
    y = mod(*args)
    names = [n for n, _ in sorted(mod.named_parameters())]
    params = [p for _, p in sorted(mod.named_parameters())]
    gr = torch.autograd.grad(y, wrt + params, gy.to(y.device), allow_unused=True)
    return y.detach(), gr[:len(wrt)], dict(zip(names, gr[len(wrt):]))

g = golden["networks"]["generator32"]
G = det_fill(M.Generator(32, 64, 2), 600).cuda().eval()
z = g["z"].cuda().requires_grad_(True)
img, _ = G([z], randomize_noise=False)
gz, gw = torch.autograd.grad(img, (z, G.convs[3].conv.weight), g["gimg"].cuda())
print("networks: img", rel(img, g["img"]), "gz", rel(gz, g["gz"]), "gw slice", rel(gw[0, :4, :4], g["gw_convs3_slice"]),
      "gw norm", float(gw.norm()), float(g["gw_convs3_norm"]))
# fp64 reference of the same thing
Gd = det_fill(T.Generator(32, 64, 2), 600).double().eval()
zd = g["z"].double().requires_grad_(True)
imgd, _ = Gd([zd], randomize_noise=False)
gzd, gwd = torch.autograd.grad(imgd, (zd, Gd.convs[3].conv.weight), g["gimg"].double())
print("  vs fp64: gw slice mine", rel(gw[0, :4, :4], gwd[0, :4, :4]), "golden(fp32 ref)", rel(g["gw_convs3_slice"], gwd[0, :4, :4]),
      "full gw mine", rel(gw, gwd))

for up in (False, True):
    for shape in [(2, 128, 128, 8), (3, 128, 256, 16)]:
        b, cin, cout, r = shape
        ref = det_fill(T.StyledConv(cin, cout, 3, 64, upsample=up), 700).double()
        mod = det_fill(M.StyledConv(cin, cout, 3, 64, upsample=up), 700).cuda()
        x, style = seeded((b, cin, r, r), 701), seeded((b, 64), 702)
        ro = 2 * r if up else r
        noise, gy = seeded((b, 1, ro, ro), 703), seeded((b, cout, ro, ro), 704)
        xr, sr = x.double().requires_grad_(True), style.double().requires_grad_(True)
        wy, wg, wp = grads(ref, (xr, sr, noise.double()), [xr, sr], gy.double())
        for backend in ("cudnn", "tcgen05"):
            L.set_conv_backend(backend)
            xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
            sc = style.cuda().requires_grad_(True)
            gy_, gg, gp = grads(mod, (xc, sc, noise.cuda()), [xc, sc], gy)
            L.set_conv_backend("cudnn")
            print(f"up={up} {shape} {backend}: y {rel(gy_, wy):.2e} dx {rel(gg[0], wg[0]):.2e} dstyle {rel(gg[1], wg[1]):.2e} " +
                  " ".join(f"{k}:{rel(gp[k], wp[k]):.2e}" for k in wp if wp[k] is not None))
