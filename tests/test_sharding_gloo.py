"""N>1 path on CPU: two gloo ranks shard a segment range the way bench.py / bn_pool_run do and
gather results in caller order; no data-path collective is involved."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from birdnet_b200.multi_gpu import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 255, 256, 257, 1000, 28800):
        for world in (1, 2, 4, 8):
            for batch in (1, 32, 256):
                spans = [shard_range(n, r, world, batch) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                for (a, b), (c, d) in zip(spans, spans[1:]):
                    assert b == c and a <= b
                # every boundary except the last falls on a whole batch
                assert all(b % batch == 0 for _, b in spans[:-1])


def test_headline_config_is_balanced():
    # config 5: 28,800 segments, ctx batch 256 -> 113 batches; 8 ranks get 14 or 15 batches
    sizes = [hi - lo for lo, hi in (shard_range(28800, r, 8, 256) for r in range(8))]
    assert sum(sizes) == 28800 and max(sizes) - min(sizes) <= 256


def _worker(rank, world, port, n, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world, batch)
    # stand-in for the per-device engine: a per-segment function of the segment index only
    local = torch.arange(lo, hi, dtype=torch.float32) * 2.0 + 1.0
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    bufs = [torch.zeros(int(s.item()), dtype=torch.float32) for s in sizes]
    dist.all_gather(bufs, local) if len({int(s.item()) for s in sizes}) == 1 else None
    if bufs and len({int(s.item()) for s in sizes}) != 1:      # ragged: gather via padded tensors
        m = max(int(s.item()) for s in sizes)
        pad = torch.zeros(m)
        pad[: hi - lo] = local
        padded = [torch.zeros(m) for _ in range(world)]
        dist.all_gather(padded, pad)
        bufs = [p[: int(s.item())] for p, s in zip(padded, sizes)]
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # bench.py: max-over-ranks time
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(bufs).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(1000, 256), (512, 256)])
def test_two_rank_gloo_gather_in_caller_order(tmp_path, n, batch):
    port = 29500 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, batch, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(got, np.arange(n, dtype=np.float32) * 2.0 + 1.0)
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
