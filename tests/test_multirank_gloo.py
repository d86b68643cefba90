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
