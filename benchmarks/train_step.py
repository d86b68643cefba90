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


def dead_parameters(gen):
    """The reference builds the ToRGB list twice (model.py:86 + :122) and its forward only
    walks the first half: the second half's parameters never receive a gradient.  They are frozen for the whole run, so
    DDP's reducer never waits for them (no find_unused_parameters graph walk per step)."""
    if len(gen.to_rgbs) != len(gen.convs):               # one ToRGB per resolution = no duplicates
        return []
    return [p for m in list(gen.to_rgbs)[len(gen.to_rgbs) // 2:] for p in m.parameters()]


def flat_grad_views(params, device):
    """One flat fp32 buffer holding the gradients of `params`; every p.grad becomes a view of it with the parameter's own
    layout (channels_last Discriminator weights are dense permutations), so backward passes accumulate straight into the
    buffer the collective reduces -- DDP's gradient_as_bucket_view with a single bucket (reference distributed.py:98-105)."""
    flat = torch.zeros(sum(p.numel() for p in params), device=device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].as_strided(p.size(), p.stride())
        off += p.numel()
    return flat


def average_gradients(flat, world, dist):
    """The data-parallel exchange of the train step: ONE all-reduce of the flat gradient buffer, averaged over ranks."""
    if world > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:                                                           # gloo (CPU tests) has no AVG
            dist.all_reduce(flat)
            flat.div_(world)


def requires_grad(params, flag):
    for p in params:
        p.requires_grad = flag


def cpu_train_step_rate(size, iters=2):
    """cpu_baseline of the train-step workload: the oracle's CPU restatement of the reference modules
    (oracle/torch_ref.GeneratorWithMap / Discriminator + oracle/sr_oracle.c rasteriser), one D step + one G step per
    iteration at batch 1 on the host cores (no regulariser iterations: a bounded sample, and they only add work)."""
    import time
    from oracle import cpu as O, torch_ref as T
    try:
        threads = max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        threads = max(1, os.cpu_count() or 1)
    prev = torch.get_num_threads()
    torch.set_num_threads(threads)
    torch.manual_seed(0)

    def rast(v, tex, tri, h, w):
        return O.rasterize(v.detach(), tex.detach(), tri, h)[0]
    G = T.GeneratorWithMap(size, 512, 8, channel_multiplier=2, rasterize=rast)
    D = T.Discriminator(size, channel_multiplier=2)
    face = SyntheticMorphableModel(189)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            vert = random_pose(face(face.random_input(1))).contiguous()
            norm = vertex_normals(vert, face.tri)
        real = torch.randn(1, 3, size, size)
        with torch.no_grad():
            fake, _, _ = G([torch.randn(1, 512)], (vert, norm, face.tri))
        for p in D.parameters():
            p.grad = None
        d_logistic_loss(D(real), D(fake)).backward()
        fake, _, _ = G([torch.randn(1, 512)], (vert, norm, face.tri))
        for p in G.parameters():
            p.grad = None
        F.softplus(-D(fake)).mean().backward()
        if it:
            times.append(time.perf_counter() - t0)
    torch.set_num_threads(prev)
    return len(times) / sum(times), threads


def run(args):
    """One process of the train-step benchmark (rank from the environment).  Returns the result dict on rank 0, None
    elsewhere.  `args`: batch, size, iters, warmup, mesh_n, conv_backend, no_cudnn_benchmark, d_nchw, profile, e2e."""
    import time
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)                                     # reference distributed.py:93-95
    from stylerenderer_b200 import _lib, layers
    from stylerenderer_b200.model import Discriminator, GeneratorWithMap

    layers.set_conv_backend(args.conv_backend)
    from stylerenderer_b200 import tc_conv
    tc_conv.set_precision(getattr(args, "precision", "tf32"))       # operand mode of the tensor-core convolutions
    # the Discriminator's 3-channel stem / 513-channel final conv are cuDNN's: let it pick its algorithms by measurement
    # and keep D in channels_last so that no NCHW<->NHWC conversion kernels run around them
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    G = GeneratorWithMap(args.size, 512, 8, channel_multiplier=2).to(dev)
    D = Discriminator(args.size, channel_multiplier=2).to(dev)
    if not args.d_nchw:
        D = D.to(memory_format=torch.channels_last)
    g_ema = copy.deepcopy(G).eval()
    face = SyntheticMorphableModel(args.mesh_n).to(dev)
    g_reg, d_reg = 4, 16
    g_ratio, d_ratio = g_reg / (g_reg + 1), d_reg / (d_reg + 1)        # train.py:529-536
    dead = dead_parameters(G)
    requires_grad(dead, False)
    dead_ids = {id(p) for p in dead}
    g_params = [p for p in G.parameters() if id(p) not in dead_ids]
    d_params = list(D.parameters())
    use_graph = getattr(args, "execution", "graph") == "graph"
    # same update as train.py:548-557, one multi-tensor kernel; capturable = the step counter lives on the device
    g_optim = torch.optim.Adam(g_params, lr=0.002 * g_ratio, betas=(0.0, 0.99 ** g_ratio), fused=True, capturable=use_graph)
    d_optim = torch.optim.Adam(d_params, lr=0.002 * d_ratio, betas=(0.0, 0.99 ** d_ratio), fused=True, capturable=use_graph)
    g_mod, d_mod = G, D
    buckets = None
    if world > 1 and not use_graph:
        from torch.nn.parallel import DistributedDataParallel as DDP
        G = DDP(G, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True)
        D = DDP(D, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True)
    B, tri = args.batch, face.tri
    ema_params, g_all_params = list(g_ema.parameters()), list(g_mod.parameters())
    mean_path = torch.zeros((), device=dev)
    real_host = torch.randn(B, 3, args.size, args.size).pin_memory()
    loss_host = torch.zeros(4).pin_memory()
    losses = {}

    from stylerenderer_b200 import mesh as mesh_frontend
    sizes = [2 ** i for i in range(2, int(math.log2(args.size)) + 1)]
    pose_sigma = torch.tensor(mesh_frontend.DEFAULT_POSE_SIGMA, device=dev)

    def sample_mesh(n):
        """Fused front-end (sr_mesh_normal_pyramid_f32): morphable-model GEMM -> pose -> vertex normals -> the normal map at
        every generator resolution as NCHW planes, no index / coefficient buffers (train.py:249-251: under no_grad)."""
        with torch.no_grad():
            T = mesh_frontend.pose_matrices(torch.randn(n, 7, device=dev) * pose_sigma)
            return mesh_frontend.normal_pyramid(face(face.random_input(n)), tri, sizes, pose=T)

    def iteration(i, e2e=False):
        nonlocal mean_path
        if e2e:                                                         # the "real" batch arrives from pinned host memory
            real = real_host.to(dev, non_blocking=True)
        else:
            real = torch.randn(B, 3, args.size, args.size, device=dev)
        if not args.d_nchw:
            real = real.contiguous(memory_format=torch.channels_last)
        # ---- D step (train.py:245-268)
        requires_grad(g_params, False); requires_grad(d_params, True)
        vert, norm, maps = sample_mesh(B)
        fake, _, _ = g_mod([torch.randn(B, 512, device=dev)], maps)
        d_loss = d_logistic_loss(D(real), D(fake))
        d_mod.zero_grad(set_to_none=True)
        d_loss.backward()
        d_optim.step()
        losses["d"] = d_loss.detach()
        if i % d_reg == 0:                                              # R1 (train.py:281-289): double backward through D
            real.requires_grad = True
            with layers.double_backward():                              # the tensor-core blocks are first-order only
                real_pred = D(real)
                r1 = d_r1_loss(real_pred, real)
                d_mod.zero_grad(set_to_none=True)
                (10 / 2 * r1 * d_reg + 0 * real_pred[0]).backward()
            d_optim.step()
            losses["r1"] = r1.detach()
        # ---- G step (train.py:292-333)
        requires_grad(g_params, True); requires_grad(d_params, False)
        vert, norm, maps = sample_mesh(B)
        fake, _, _ = G([torch.randn(B, 512, device=dev)], maps)
        g_loss = F.softplus(-d_mod(fake)).mean()
        g_mod.zero_grad(set_to_none=True)
        g_loss.backward()
        g_optim.step()
        losses["g"] = g_loss.detach()
        if i % g_reg == 0:                                              # path length (train.py:335-354): double backward through G
            pb = max(1, B // 2)
            v = vert[:pb].clone().requires_grad_(True)
            n = norm[:pb].clone().requires_grad_(True)
            with layers.double_backward():
                # the unwrapped module (the loss differentiates the output w.r.t. intermediate tensors, which DDP's
                # output copies would break); the gradient all-reduce of this step is issued explicitly below
                fake, latents, normals = g_mod([torch.randn(pb, 512, device=dev)], (v, n, tri), return_latents=True,
                                               return_normals=True)
                path_loss, mean_path = g_path_regularize(fake, [latents] + normals, mean_path)
                g_mod.zero_grad(set_to_none=True)
                (2 * g_reg * path_loss + 0 * fake[0, 0, 0, 0]).backward()
            if world > 1:
                grads = [p.grad for p in g_params if p.grad is not None]
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat)
                flat.div_(world)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            g_optim.step()
            losses["path"] = path_loss.detach()
        with torch.no_grad():                                           # EMA (train.py:100-104,358), 2 launches instead of ~500
            torch._foreach_mul_(ema_params, 0.999)
            torch._foreach_add_(ema_params, g_all_params, alpha=0.001)
        if e2e:                                                         # the loop reads its scalars every iteration (train.py:360-372)
            loss_host[0].copy_(losses["d"], non_blocking=True)
            loss_host[1].copy_(losses["g"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    # ---- execution = "graph": every phase of the iteration (D step, R1, G step, path length: forward + backward into flat
    # gradient buffers) and the two optimiser steps are captured ONCE as CUDA graphs and replayed; the data-parallel exchange is
    # one NCCL all-reduce (AVG) of the flat gradient buffer between the backward graph and the optimiser graph -- the same
    # gradients DDP's reducer would average (reference distributed.py:98-105), without its per-parameter hooks, bucket
    # copies and the ~3000 eager launches per iteration whose host cost is what limits the eager step when 8 ranks share
    # the box's host cores (profiles/r2_train_step_scaling.md).
    graphs, graph_launches = {}, {}
    if use_graph:
        flat_g, flat_d = flat_grad_views(g_params, dev), flat_grad_views(d_params, dev)
        real_static = torch.randn(B, 3, args.size, args.size, device=dev)
        if not args.d_nchw:
            real_static = real_static.contiguous(memory_format=torch.channels_last)
        loss_buf = torch.zeros(4, device=dev)
        pb = max(1, B // 2)
        vert_buf = torch.zeros(pb, args.mesh_n ** 2, 3, device=dev)
        norm_buf = torch.zeros(pb, args.mesh_n ** 2, 3, device=dev)

        def d_phase():
            requires_grad(g_params, False); requires_grad(d_params, True)
            _, _, maps = sample_mesh(B)
            fake, _, _ = g_mod([torch.randn(B, 512, device=dev)], maps)
            d_loss = d_logistic_loss(d_mod(real_static), d_mod(fake))
            flat_d.zero_()
            d_loss.backward()
            loss_buf[0].copy_(d_loss.detach())

        def r1_phase():
            requires_grad(g_params, False); requires_grad(d_params, True)
            real = real_static.detach().requires_grad_(True)
            with layers.double_backward():
                real_pred = d_mod(real)
                r1 = d_r1_loss(real_pred, real)
                flat_d.zero_()
                (10 / 2 * r1 * d_reg + 0 * real_pred[0]).backward()
            loss_buf[1].copy_(r1.detach())

        def g_phase():
            requires_grad(g_params, True); requires_grad(d_params, False)
            vert, norm, maps = sample_mesh(B)
            vert_buf.copy_(vert[:pb]); norm_buf.copy_(norm[:pb])
            fake, _, _ = g_mod([torch.randn(B, 512, device=dev)], maps)
            g_loss = F.softplus(-d_mod(fake)).mean()
            flat_g.zero_()
            g_loss.backward()
            loss_buf[2].copy_(g_loss.detach())

        def path_phase():
            requires_grad(g_params, True); requires_grad(d_params, False)
            v = vert_buf.clone().requires_grad_(True)
            n = norm_buf.clone().requires_grad_(True)
            with layers.double_backward():
                fake, latents, normals = g_mod([torch.randn(pb, 512, device=dev)], (v, n, tri), return_latents=True,
                                               return_normals=True)
                path_loss, new_mean = g_path_regularize(fake, [latents] + normals, mean_path)
                flat_g.zero_()
                (2 * g_reg * path_loss + 0 * fake[0, 0, 0, 0]).backward()
            mean_path.copy_(new_mean)
            loss_buf[3].copy_(path_loss.detach())

        def d_opt():
            d_optim.step()

        def g_opt():
            g_optim.step()
            with torch.no_grad():
                torch._foreach_mul_(ema_params, 0.999)
                torch._foreach_add_(ema_params, g_all_params, alpha=0.001)

        phases = {"d": d_phase, "r1": r1_phase, "g": g_phase, "path": path_phase, "d_opt": d_opt, "g_opt": g_opt}
        order = ["d", "d_opt", "r1", "d_opt", "g", "g_opt", "path", "g_opt"]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                   # warm-up: optimiser state, cuDNN choices, lazy inits
            for _ in range(3):
                for name in order:
                    phases[name]()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = None
        for name, fn in phases.items():
            gr = torch.cuda.CUDAGraph()
            n_before = _lib.launch_count()
            with torch.cuda.graph(gr, pool=pool):
                fn()
            pool = pool or gr.pool()
            graphs[name], graph_launches[name] = gr, _lib.launch_count() - n_before

        def reduce_(flat):
            average_gradients(flat, world, dist)

        def iteration_graph(i, e2e=False):
            if e2e:
                real_static.copy_(real_host, non_blocking=True)
            else:
                real_static.normal_()
            graphs["d"].replay(); reduce_(flat_d); graphs["d_opt"].replay()
            if i % d_reg == 0:
                graphs["r1"].replay(); reduce_(flat_d); graphs["d_opt"].replay()
            graphs["g"].replay(); reduce_(flat_g); graphs["g_opt"].replay()
            if i % g_reg == 0:
                graphs["path"].replay(); reduce_(flat_g); graphs["g_opt"].replay()
            if e2e:
                loss_host.copy_(loss_buf, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        eager_iteration, iteration = iteration, iteration_graph

    def timed(n_iters, e2e=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n_iters):
            iteration(i, e2e)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(args.warmup, 3)):
        iteration(i)
    n0 = _lib.launch_count()
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
            ms = timed(args.iters)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90), file=sys.stderr)
        # which torch ops (with their input shapes) still own device time: the library / elementwise leftovers
        print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=250, max_name_column_width=50,
                                                                  max_shapes_column_width=110), file=sys.stderr)
    else:
        ms = timed(args.iters)
    launches = (_lib.launch_count() - n0) // args.iters
    phase_ms = None
    if use_graph:                                                       # device time of every phase graph (3 replays each)
        phase_ms = {}
        for name, gr in graphs.items():
            torch.cuda.synchronize()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            for _ in range(3):
                gr.replay()
            e_.record()
            torch.cuda.synchronize()
            phase_ms[name] = round(s_.elapsed_time(e_) / 3, 3)
        phase_ms["per_iteration"] = round(phase_ms["d"] + phase_ms["g"] + phase_ms["d_opt"] + phase_ms["g_opt"]
                                          + (phase_ms["r1"] + phase_ms["d_opt"]) / d_reg + (phase_ms["path"] + phase_ms["g_opt"]) / g_reg, 3)
    if use_graph:                                                       # replays do not pass through the C ABI: count the captures
        gl = graph_launches
        launches = int(gl["d"] + gl["g"] + gl["r1"] / d_reg + gl["path"] / g_reg)
        lb = loss_buf.tolist()
        losses.update({"d": torch.tensor(lb[0]), "r1": torch.tensor(lb[1]), "g": torch.tensor(lb[2]), "path": torch.tensor(lb[3])})
    ms_e2e = timed(args.iters, e2e=True) if getattr(args, "e2e", True) else None

    # ---- the collective: DDP's gradient all-reduce (bucket layout from the reducer) and its stand-alone cost
    comm = None
    if world > 1:
        def bucket_info(m):
            try:
                sizes = [int(x) for x in m._get_ddp_logging_data().get("bucket_sizes", "").split(",") if x.strip()]
                return {"buckets": len(sizes), "bytes": sum(sizes)}
            except Exception:                                           # noqa: BLE001
                return None
        n_g = sum(p.numel() for p in g_params)
        n_d = sum(p.numel() for p in d_params)
        flat = torch.zeros(n_g + n_d, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e.record()
        torch.cuda.synchronize()
        ar_ms = s.elapsed_time(e) / 10
        comm = {"collective": ("NCCL all-reduce (AVG) of the flat G / D gradient buffers, one call per backward pass, between the "
                               "backward graph and the optimiser graph" if use_graph else
                               "NCCL all-reduce of the G and D gradients (DDP buckets, overlapped with the backward)"),
                "gradient_bytes_per_iteration": 4 * (n_g + n_d),
                "generator": {"buckets": 1, "bytes": 4 * n_g} if use_graph else bucket_info(G),
                "discriminator": {"buckets": 1, "bytes": 4 * n_d} if use_graph else bucket_info(D),
                "allreduce_ms_standalone": round(ar_ms, 3),
                "allreduce_busbw_GBps": round(2 * (world - 1) / world * 4 * (n_g + n_d) / (ar_ms * 1e-3) / 1e9, 1)}

    # ---- sanity: a benchmark of a diverged model measures nothing
    finite = all(bool(torch.isfinite(v).all()) for v in losses.values())
    finite = finite and all(bool(torch.isfinite(p).all()) for p in list(g_params)[:40] + list(d_params)[:20])
    if not finite:
        raise RuntimeError(f"train step diverged: losses {dict((k, float(v)) for k, v in losses.items())}")

    state = dict(iteration=(eager_iteration if use_graph else iteration), lib=_lib, dev=dev, world=world, rank=rank, local=local,
                 B=B, tri=tri, execution="cuda_graph_replay" if use_graph else "eager")
    if rank != 0:
        return None, state
    res = {
        "metric": "GAR train step (G+D+rasterize+R1/16+path/4) images/sec", "value": round(world * B * args.iters / (ms * 1e-3), 2),
        "unit": "images/s", "n_gpus": world, "iters": args.iters, "ms_per_iter": round(ms / args.iters, 2),
        "scaling": "weak",
        "dtype": ("bf16 conv operands (tcgen05 kind::f16), fp32 accumulation / storage / parameter gradients"
                  if getattr(args, "precision", "tf32") == "bf16" else "f32 storage, tf32 tensor-core convs"),
        "precision": getattr(args, "precision", "tf32"),
        "conv_backend": {"generator": args.conv_backend, "discriminator": args.conv_backend + " (ResBlock convs; 3-channel stem, final conv on cuDNN)"},
        "config": {"workload": "GeneratorWithMap + Discriminator 256x256 (BASELINE.json configs[3])", "per_gpu_batch": B,
                   "parallelism": (f"dp{world} (flat-buffer NCCL gradient all-reduce between graph replays, dead ToRGB copies frozen)"
                                   if use_graph else
                                   f"ddp{world} (NCCL gradient all-reduce, broadcast_buffers=False, dead ToRGB copies frozen)"),
                   "execution": "cuda_graph_replay (4 phase graphs + 2 optimiser graphs)" if use_graph else "eager",
                   "mesh": f"{args.mesh_n ** 2} verts / {tri.shape[0]} tris"},
        "gpu_launches_per_iter": launches, "losses": {k: round(float(v), 5) for k, v in losses.items()},
        "phase_ms": phase_ms,
        "collective": comm}
    if ms_e2e is not None:
        res["e2e"] = {"value": round(world * B * args.iters / (ms_e2e * 1e-3), 2), "unit": "images/s",
                      "h2d_bytes_per_step": real_host.numel() * 4, "d2h_bytes_per_step": 8,
                      "ms_per_step": round(ms_e2e / args.iters, 2),
                      "what": "the real-image batch from pinned host memory every iteration, D / G losses read back"}
    return res, state


def default_args(**over):
    ns = argparse.Namespace(batch=16, size=256, iters=16, warmup=3, mesh_n=189, conv_backend="tcgen05",
                            no_cudnn_benchmark=False, d_nchw=False, profile=False, e2e=True, precision="bf16",
                            execution="graph")
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (reference default, train.py:432)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=16, help="timed iterations (a multiple of 16 covers whole regulariser cycles)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mesh-n", type=int, default=189)
    ap.add_argument("--conv-backend", default="tcgen05", choices=["cudnn", "tcgen05"],
                    help="who runs the ModulatedConv2d contractions of G and the ResBlock convs of D")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="keep cuDNN's heuristic algorithm choice for the Discriminator")
    ap.add_argument("--d-nchw", action="store_true", help="keep the Discriminator in NCHW (default: channels_last, no layout conversions)")
    ap.add_argument("--profile", action="store_true", help="print the top CUDA kernels of the timed iterations (torch.profiler)")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false")
    ap.add_argument("--execution", default="graph", choices=["graph", "eager"],
                    help="graph: CUDA-graph replay of the phases + flat-buffer all-reduce (default); eager: torch DDP, eager launches")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="operand mode of the tensor-core convolutions (BASELINE.json configs[3] asks for bf16)")
    args = ap.parse_args()
    res, _ = run(args)
    if res is not None:
        print(json.dumps(res), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
