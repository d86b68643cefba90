#!/usr/bin/env python
"""BASELINE.json configs[3]: full GAR train step (G + D + rasterize + R1 + path-length regulariser) at 256x256 with DDP.

A benchmark harness, not a training product: it restates the per-iteration work of the reference's train loop
(reference train.py:239-358: D step, R1 every 16, G step, path regulariser every 4 with path_batch_shrink=2, EMA)
on synthetic data -- random "real" images, a synthetic 3DMM (grid mesh + random low-frequency basis) -- with
`stylerenderer_b200.model.GeneratorWithMap` / `Discriminator`, the B200 rasteriser and ops, Adam with the lazy-regularisation
betas (train.py:529-536) and `torch.nn.parallel.DistributedDataParallel(broadcast_buffers=False)` over NCCL when launched
under torchrun (reference distributed.py:98-105).  The only collective of the path is DDP's gradient all-reduce.

  python benchmarks/train_step.py [--batch 16] [--iters 16] [--size 256]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/train_step.py ...

Prints one JSON line on rank 0: G images/s (global batch x iterations / time), ms per iteration averaged over a full
16-iteration regulariser cycle.
"""
import argparse
import copy
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch import autograd  # noqa: E402


# ---------------------------------------------------------------- synthetic 3DMM front-end (reference face_model.py / utils_3d.py)
class SyntheticMorphableModel(torch.nn.Module):
    """`LinearMorphableModel` (reference face_model.py:4-74) with a synthetic mean shape and a low-frequency random basis."""

    def __init__(self, n=189, dims=(80, 64), seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        lin = torch.linspace(-0.8, 0.8, n)
        ys, xs = torch.meshgrid(lin, lin, indexing="ij")
        mean = torch.stack([xs, ys, 0.5 * torch.exp(-2 * (xs ** 2 + ys ** 2))], -1).view(-1)
        k = sum(dims)
        fx, fy = torch.rand(k, 3, generator=g) * 3, torch.rand(k, 3, generator=g) * 3
        basis = (torch.sin(xs.reshape(1, -1, 1) * fx.view(k, 1, 3)) * torch.cos(ys.reshape(1, -1, 1) * fy.view(k, 1, 3)))
        self.register_buffer("mean", mean)
        self.register_buffer("basis", basis.reshape(k, -1) * (0.05 / math.sqrt(k)))
        self.register_buffer("sigma", torch.cat([torch.ones(dims[0]), torch.full((dims[1],), 0.01)]))
        idx = torch.arange(n * n).view(n, n)
        a, b, c, d = idx[:-1, :-1].reshape(-1), idx[:-1, 1:].reshape(-1), idx[1:, :-1].reshape(-1), idx[1:, 1:].reshape(-1)
        self.register_buffer("tri", torch.cat([torch.stack([a, b, c], 1), torch.stack([b, d, c], 1)], 0).contiguous())

    def random_input(self, batch):                                    # face_model.py:69-70
        return torch.randn(batch, self.sigma.numel(), device=self.sigma.device) * self.sigma

    def forward(self, x):                                             # face_model.py:71-72
        return (self.mean + x @ self.basis).view(x.shape[0], -1, 3)


def random_pose(v, p=(.5, .1, .05, .1, .1, .1, .15)):
    """reference utils_3d.py:360-378 `random_apply_pose3D`: random yaw/pitch/roll, translation, log-scale."""
    b = v.shape[0]
    z = torch.randn(b, 7, device=v.device) * torch.tensor(p, device=v.device)
    cy, sy, cx, sx, cz, sz = z[:, 0].cos(), z[:, 0].sin(), z[:, 1].cos(), z[:, 1].sin(), z[:, 2].cos(), z[:, 2].sin()
    one, zero = torch.ones_like(cy), torch.zeros_like(cy)
    ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).view(b, 3, 3)
    rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], -1).view(b, 3, 3)
    rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], -1).view(b, 3, 3)
    rot = torch.exp(z[:, 6]).view(b, 1, 1) * (ry @ rx @ rz)
    return v @ rot + z[:, 3:6].view(b, 1, 3)


def vertex_normals(v, tri):
    """reference utils_3d.py:379-404 `mesh_point_normal`: area-weighted face normals scattered to vertices, normalised."""
    a, b, c = v[:, tri[:, 0]], v[:, tri[:, 1]], v[:, tri[:, 2]]
    fn = torch.cross(b - a, c - a, dim=-1)
    vn = torch.zeros_like(v)
    for j in range(3):
        vn.index_add_(1, tri[:, j], fn)
    return F.normalize(vn, dim=-1, eps=1e-8)


# ---------------------------------------------------------------- losses (reference train.py:105-134)
def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    grad_real, = autograd.grad(real_pred.sum(), real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_path_regularize(fake_img, latents, mean_path_length, decay=0.01):
    noise = torch.randn_like(fake_img) / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    grads = autograd.grad((fake_img * noise).sum(), latents, create_graph=True, allow_unused=True)
    path_lengths = 0
    for g in grads:
        if g is not None:
            path_lengths = path_lengths + torch.sqrt(g.reshape(g.shape[0], -1).pow(2).sum(1))
    path_mean = mean_path_length + decay * (path_lengths.mean() - mean_path_length)
    return (path_lengths - path_mean).pow(2).mean(), path_mean.detach()


def requires_grad(model, flag):
    for p in model.parameters():
        p.requires_grad = flag


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (reference default, train.py:432)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=16, help="timed iterations (a multiple of 16 covers whole regulariser cycles)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mesh-n", type=int, default=189)
    ap.add_argument("--conv-backend", default="tcgen05", choices=["cudnn", "tcgen05"],
                    help="who runs the ModulatedConv2d contractions of G (the Discriminator's plain convs stay on cuDNN)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="keep cuDNN's heuristic algorithm choice for the Discriminator")
    ap.add_argument("--d-nchw", action="store_true", help="keep the Discriminator in NCHW (default: channels_last, no layout conversions)")
    ap.add_argument("--profile", action="store_true", help="print the top CUDA kernels of the timed iterations (torch.profiler)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)                                     # reference distributed.py:93-95
    from stylerenderer_b200 import _lib, layers
    from stylerenderer_b200.model import Discriminator, GeneratorWithMap

    layers.set_conv_backend(args.conv_backend)
    # the Discriminator's plain convolutions are cuDNN's: let it pick its algorithms by measurement (the heuristic picks
    # 20 TFLOP/s "sm80 indexed" kernels for the 64-channel layers at 256^2) and keep D in channels_last so that no
    # NCHW<->NHWC conversion kernels run around them (our upfirdn2d / fused_leaky_relu take channels_last as is)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    G = GeneratorWithMap(args.size, 512, 8, channel_multiplier=2).to(dev)
    D = Discriminator(args.size, channel_multiplier=2).to(dev)
    if not args.d_nchw:
        D = D.to(memory_format=torch.channels_last)
    g_ema = copy.deepcopy(G).eval()
    face = SyntheticMorphableModel(args.mesh_n).to(dev)
    g_reg, d_reg = 4, 16
    g_ratio, d_ratio = g_reg / (g_reg + 1), d_reg / (d_reg + 1)        # train.py:529-536
    g_optim = torch.optim.Adam(G.parameters(), lr=0.002 * g_ratio, betas=(0.0, 0.99 ** g_ratio))
    d_optim = torch.optim.Adam(D.parameters(), lr=0.002 * d_ratio, betas=(0.0, 0.99 ** d_ratio))
    g_mod, d_mod = G, D
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        # find_unused_parameters: the reference builds the ToRGB list twice and never uses the second copy (SURVEY.md
        # section 4 quirk 4), so six modules' parameters never receive gradients
        G = DDP(G, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True, find_unused_parameters=True)
        D = DDP(D, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True)
    B, tri = args.batch, face.tri
    mean_path = torch.zeros((), device=dev)

    from stylerenderer_b200 import mesh as mesh_frontend

    def sample_mesh(n):
        with torch.no_grad():
            vert = random_pose(face(face.random_input(n))).contiguous()
            return vert, mesh_frontend.mesh_point_normal(vert, tri)          # sr_mesh_vertex_normals_f32

    def iteration(i):
        nonlocal mean_path
        real = torch.randn(B, 3, args.size, args.size, device=dev)
        if not args.d_nchw:
            real = real.contiguous(memory_format=torch.channels_last)
        # ---- D step (train.py:245-268)
        requires_grad(g_mod, False); requires_grad(d_mod, True)
        vert, norm = sample_mesh(B)
        fake, _, _ = g_mod([torch.randn(B, 512, device=dev)], (vert, norm, tri))
        d_loss = d_logistic_loss(D(real), D(fake))
        d_mod.zero_grad(set_to_none=True)
        d_loss.backward()
        d_optim.step()
        if i % d_reg == 0:                                              # R1 (train.py:281-289): double backward through D
            real.requires_grad = True
            with layers.double_backward():                              # the tensor-core blocks are first-order only
                real_pred = D(real)
                r1 = d_r1_loss(real_pred, real)
                d_mod.zero_grad(set_to_none=True)
                (10 / 2 * r1 * d_reg + 0 * real_pred[0]).backward()
            d_optim.step()
        # ---- G step (train.py:292-333)
        requires_grad(g_mod, True); requires_grad(d_mod, False)
        vert, norm = sample_mesh(B)
        fake, _, _ = G([torch.randn(B, 512, device=dev)], (vert, norm, tri))
        g_loss = F.softplus(-d_mod(fake)).mean()
        g_mod.zero_grad(set_to_none=True)
        g_loss.backward()
        g_optim.step()
        if i % g_reg == 0:                                              # path length (train.py:335-354): double backward through G
            pb = max(1, B // 2)
            v = vert[:pb].clone().requires_grad_(True)
            n = norm[:pb].clone().requires_grad_(True)
            with layers.double_backward():
                # the unwrapped module: DDP(find_unused_parameters=True) hands back copies of the outputs, which
                # cannot be differentiated against; the gradient all-reduce of this step is issued explicitly below
                fake, latents, normals = g_mod([torch.randn(pb, 512, device=dev)], (v, n, tri), return_latents=True,
                                               return_normals=True)
                path_loss, mean_path = g_path_regularize(fake, [latents] + normals, mean_path)
                g_mod.zero_grad(set_to_none=True)
                (2 * g_reg * path_loss + 0 * fake[0, 0, 0, 0]).backward()
            if world > 1:
                grads = [p.grad for p in g_mod.parameters() if p.grad is not None]
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat)
                flat.div_(world)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            g_optim.step()
        with torch.no_grad():                                           # EMA (train.py:100-104,358)
            for pe, p in zip(g_ema.parameters(), g_mod.parameters()):
                pe.mul_(0.999).add_(p.detach(), alpha=0.001)

    for i in range(args.warmup):
        iteration(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(args.iters):
                iteration(i)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90), file=sys.stderr)
    else:
        for i in range(args.iters):
            iteration(i)
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "GAR train step (G+D+rasterize+R1/16+path/4) images/sec", "value": round(world * B * args.iters / (ms * 1e-3), 2),
            "unit": "images/s", "n_gpus": world, "iters": args.iters, "ms_per_iter": round(ms / args.iters, 2),
            "scaling": "weak", "dtype": "f32 storage, tf32 tensor-core convs",
            "conv_backend": {"generator": args.conv_backend, "discriminator": args.conv_backend + " (ResBlock convs; 3-channel stem, final conv and regulariser iterations on cuDNN)"},
            "config": {"workload": "GeneratorWithMap + Discriminator 256x256 (BASELINE.json configs[3])", "per_gpu_batch": B,
                       "parallelism": f"ddp{world} (NCCL gradient all-reduce, broadcast_buffers=False)",
                       "mesh": f"{args.mesh_n ** 2} verts / {tri.shape[0]} tris"},
            "gpu_launches_per_iter": (_lib.launch_count() - n0) // args.iters}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
