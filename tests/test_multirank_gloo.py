"""world_size-2 `gloo` test (CPU) of the host-side logic of the N > 1 bench path: the path shards by image with no
data-path collective, so what the ranks share is (a) distinct per-rank seeds, (b) a barrier and
(c) the max-over-ranks time that turns per-rank work into the whole-job rate."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    torch.manual_seed(bench.rank_seed(1234, rank))
    z = torch.randn(4, 8)                                   # this rank's shard of latents
    gathered = [torch.empty_like(z) for _ in range(world)]
    dist.all_gather(gathered, z)
    dist.barrier()
    local_ms = 10.0 * (rank + 1)                            # rank 1 is the slow one
    ms = bench.max_over_ranks_ms(local_ms)
    value = bench.whole_job_rate(32, 5, ms, world)
    if rank == 0:
        torch.save({"ms": ms, "value": value, "distinct": not torch.equal(gathered[0], gathered[1])}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_max_time_and_distinct_shards(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["distinct"], "ranks must draw different shards"
    assert r["ms"] == 20.0, "whole-job time is the slowest rank's"
    assert abs(r["value"] - 2 * 32 * 5 / 20e-3) < 1e-6


def test_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.max_over_ranks_ms(3.5) == 3.5
    assert bench.whole_job_rate(32, 10, 200.0, 1) == 32 * 10 / 0.2


def _grad_worker(rank, world, port, out):
    """The train step's data-parallel exchange (benchmarks/train_step.py: flat gradient buffer viewed by every p.grad + one
    averaged all-reduce) against torch's DistributedDataParallel on the same two-rank problem."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sr_train_step", os.path.join(ROOT, "benchmarks", "train_step.py"))
    ts = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ts)

    def net():
        torch.manual_seed(7)                               # identical replicas
        m = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.LeakyReLU(0.2), torch.nn.Conv2d(8, 4, 1))
        return m.to(memory_format=torch.channels_last)     # strided parameters, like the Discriminator's
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 3, 9, 9)                            # this rank's shard
    ref = torch.nn.parallel.DistributedDataParallel(net())
    ref(x).square().mean().backward()
    mine = net()
    params = list(mine.parameters())
    flat = ts.flat_grad_views(params, torch.device("cpu"))
    for _ in range(2):                                     # a second pass must start from zeroed buffers, as the phase graphs do
        flat.zero_()
        mine(x).square().mean().backward()
        assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in params)
        ts.average_gradients(flat, world, dist)
    err = max((p.grad - q.grad).abs().max().item() for p, q in zip(params, ref.module.parameters()))
    if rank == 0:
        torch.save({"err": err, "numel": flat.numel(), "want": sum(p.numel() for p in params)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_ddp(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["numel"] == r["want"]
    assert r["err"] < 1e-6, r
