#!/usr/bin/env python
"""Latent inversion (BASELINE.json configs[4] / SURVEY.md 8(d) config 5): per-face w+ latents [14,512] optimised with
Adam so that G(w+) matches a target image under a perceptual + pixel loss.  Faces are sharded over the GPUs of the box --
one process per GPU, NO data-path collective; `value` = faces x Adam steps per second over all ranks.

  python benchmarks/inversion.py [--faces 64] [--batch 32] [--steps 20] [--warmup 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      benchmarks/inversion.py --faces 64        # 64 faces PER GPU (config 5: 512 faces on 8 GPUs)

What runs where: the generator forward + backward-to-the-latents is this repository's chained tcgen05 path with every
weight frozen, so the weight-gradient GEMMs are skipped (fused.StyledLayerTC).  The reference ships no inversion script
and its LPIPS backbone needs ImageNet VGG16 weights that are not available offline; the perceptual term here is the
LPIPS formula (reference lpips/networks_basic.py:64-92: unit-normalised features of 5 VGG16 stages, squared difference,
1x1 heads, spatial mean) over a VGG16-shaped stack with SEEDED RANDOM weights -- the same flops and memory traffic as
LPIPS-VGG (reference lpips/pretrained_networks.py:97-137), not its metric values.  Its 3x3 conv + bias + ReLU layers whose
channel counts are multiples of 128 (10 of the 13) run on this repository's tensor-core kernels too (fused.PlainConvTC with
alpha = 0: forward + dgrad, no weight gradient); conv1_1 / conv1_2 / conv2_1 (3 -> 64 -> 64 -> 128) stay on cuDNN.
The whole Adam step -- G forward, perceptual stack, backward to the latents, fused Adam -- is captured ONCE as a CUDA graph
and replayed (`--no-graph` times the eager step).  Default: 64 faces per GPU, 100 timed Adam steps (a bounded sample of
config 5's 1000; `--steps 1000` runs them all).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import bench  # noqa: E402

VGG16_STAGES = [[64, 64], [128, 128], [256, 256, 256], [512, 512, 512], [512, 512, 512]]   # torchvision cfg "D"


class ConvReLU(nn.Module):
    """3x3 conv (pad 1) + bias + ReLU of the VGG16 stages.  On CUDA, with channel counts the tensor-core kernels take, it
    is one fused.PlainConvTC block (implicit GEMM + bias + ReLU epilogue; backward = mask pass + dgrad GEMM, the frozen
    weight needs no wgrad); otherwise stock torch ops (CPU tests, the 3- and 64-channel layers)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.empty(cout))
        self.stride, self.padding = 1, 1                               # what fused.plain_conv_supported looks at

    def forward(self, x):
        if x.is_cuda:
            from stylerenderer_b200 import fused, layers
            if layers.get_conv_backend() == "tcgen05" and fused.plain_conv_supported(self, x) == "s1":
                return fused.PlainConvTC.apply(x, self.weight, self.bias, 1.0, "s1", 0.0, 1.0)
        return torch.relu(nn.functional.conv2d(x, self.weight, self.bias, padding=1))


class PerceptualStack(nn.Module):
    """LPIPS-shaped distance (reference lpips/networks_basic.py:27-92, pretrained_networks.py:97-137) with seeded
    random weights: ScalingLayer -> 5 VGG16 stages (features after relu1_2 ... relu5_3) -> normalise over channels ->
    squared difference -> non-negative 1x1 head -> spatial mean -> sum over stages."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor([.458, .448, .450]).view(1, 3, 1, 1))
        stages, cin = [], 3
        for i, widths in enumerate(VGG16_STAGES):
            layers = [nn.MaxPool2d(2, 2)] if i else []
            for c in widths:
                conv = ConvReLU(cin, c)
                with torch.no_grad():                                  # He init keeps activations O(1) through 13 layers
                    conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (9 * cin)) ** 0.5)
                    conv.bias.zero_()
                layers.append(conv)
                cin = c
            stages.append(nn.Sequential(*layers))
        self.stages = nn.ModuleList(stages)
        self.heads = nn.ParameterList([nn.Parameter(torch.rand(1, w[-1], 1, 1, generator=g) / w[-1]) for w in VGG16_STAGES])
        for p in self.parameters():
            p.requires_grad_(False)

    def features(self, img):
        x = (img - self.shift) / self.scale
        feats = []
        for st in self.stages:
            x = st(x)
            feats.append(x * torch.rsqrt((x * x).sum(1, keepdim=True) + 1e-20))
        return feats

    def distance(self, feats, target_feats):
        d = 0
        for f, t, h in zip(feats, target_feats, self.heads):
            d = d + ((f - t) ** 2 * h).sum(1).mean((1, 2))
        return d                                                        # [B]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--faces", type=int, default=64, help="faces per GPU (config 5: 64)")
    ap.add_argument("--batch", type=int, default=32, help="faces per generator launch set")
    ap.add_argument("--steps", type=int, default=100, help="timed Adam steps per face (config 5: 1000)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="time the eager Adam step instead of the CUDA-graph replay")
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--pixel-weight", type=float, default=0.1)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "inversion needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from stylerenderer_b200 import _lib, layers
    layers.set_conv_backend("tcgen05")

    G = bench.build_generator(dev).eval()                               # same weights on every rank (seed 0)
    for p in G.parameters():
        p.requires_grad_(False)
    P = PerceptualStack().to(dev).to(memory_format=torch.channels_last)
    n_faces, B = args.faces, args.batch
    assert n_faces % B == 0, "--faces must be a multiple of --batch"
    torch.manual_seed(bench.rank_seed(777, rank))                       # this rank's shard of target faces
    with torch.no_grad():
        w_mean = G.style(torch.randn(4096, 512, device=dev)).mean(0)    # start of every optimisation: the mean latent
        targets, target_feats = [], []
        for _ in range(n_faces // B):
            img, _ = G([torch.randn(B, 512, device=dev)], randomize_noise=False)
            targets.append(img)
            target_feats.append(P.features(img))
    latents = [w_mean.view(1, 1, 512).repeat(B, G.n_latent, 1).clone().requires_grad_(True) for _ in range(n_faces // B)]
    opts = [torch.optim.Adam([w], lr=args.lr, fused=True, capturable=True) for w in latents]
    loss_host = torch.empty(n_faces // B, B).pin_memory()
    tc_convs = sum(1 for m in P.modules() if isinstance(m, ConvReLU) and m.weight.shape[1] % 128 == 0 and m.weight.shape[0] % 128 == 0)

    def adam_step():
        """One Adam step for every face of this rank (chunks of B faces)."""
        for k, (w, opt) in enumerate(zip(latents, opts)):
            opt.zero_grad(set_to_none=True)
            img, _ = G([w], input_is_latent=True, randomize_noise=False)
            per_face = P.distance(P.features(img), target_feats[k]) + args.pixel_weight * ((img - targets[k]) ** 2).mean((1, 2, 3))
            per_face.sum().backward()
            opt.step()
            loss_host[k].copy_(per_face.detach(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    adam_step()
    barrier()
    first = loss_host.mean().item()
    for _ in range(max(args.warmup - 1, 2)):
        adam_step()
    barrier()
    n0 = _lib.launch_count()
    adam_step()
    launches = _lib.launch_count() - n0
    step, execution = adam_step, "eager"
    if not args.no_graph:                                # capture the whole Adam step once, replay it
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                adam_step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            adam_step()
        step, execution = graph.replay, "cuda_graph_replay"
        for _ in range(3):
            step()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with bench.ClockSampler(local) as clocks:
        s.record()
        for _ in range(args.steps):
            step()
        e.record()
        barrier()
    ms = bench.max_over_ranks_ms(s.elapsed_time(e), dev)
    last = loss_host.mean().item()
    assert last == last and last < first, f"inversion does not converge: loss {first} -> {last}"
    roof = None
    if rank == 0:                                        # live per-kernel timing of one eager Adam step (CUDA events per C-ABI call)
        peaks = bench.measured_peaks()
        tf = bench.measure_tf32_peak(dev, sustain_s=1.0)
        peaks.update({"tf32_tflops": tf["tf32_tflops"], "tf32_tflops_sustained": tf["tf32_tflops_sustained"],
                      "tf32_how": "measured in this run: " + tf["how"]})
        names = ["sr_fused_bias_act_f32", "sr_fused_lrelu_backward_f32", "sr_upfirdn2d_f32"] + list(_lib.CONV_EXPORTS)
        with bench.KernelTimer(_lib, names) as kt:
            torch.cuda.synchronize()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record()
            adam_step()
            e2.record()
            torch.cuda.synchronize()
        roof = bench.dominant_kernel_roofline(kt.stats(), peaks, s2.elapsed_time(e2), 1)
    if rank == 0:
        value = bench.whole_job_rate(n_faces, args.steps, ms, world)
        print(json.dumps({
            "metric": "latent inversion face-steps/sec @256px (G fwd + bwd to w+ latents, perceptual + pixel loss, Adam)",
            "value": round(value, 1), "unit": "face-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": f"latent inversion: {n_faces} faces/GPU in chunks of {B}, w+ [14,512] per face, Adam "
                                   "(BASELINE.json configs[4]; bounded sample of its 1000 steps)",
                       "global_faces": world * n_faces, "parallelism": f"dp{world} (face-sharded, no collective)",
                       "perceptual": f"LPIPS formula over a VGG16-shaped stack with seeded random weights (ImageNet weights "
                                     f"unavailable offline); {tc_convs} of its 13 conv layers on the tensor-core kernels, "
                                     "conv1_1 / conv1_2 / conv2_1 on cuDNN",
                       "execution": execution,
                       "frozen_weights": "generator / perceptual weight-gradient GEMMs skipped"},
            "clocks": clocks.summary(), "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None,
            "e2e": {"value": round(value, 1), "unit": "face-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n_faces,
                    "what": "latents, targets and optimiser state live on the device for the whole optimisation; every step "
                            "reads the per-face losses back to pinned host memory (inside the timed region)"},
            "loss_first_step": round(first, 5), "loss_last_step": round(last, 5),
            "seconds_for_1000_steps": round(ms / args.steps, 3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
