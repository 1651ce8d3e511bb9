"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding plan (no data-path collective;
the only communication is the final gather of per-rank results)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tools_b200.sharding import shard


def test_shard_partition():
    for total in (0, 1, 7, 1000, 1_000_000, 4_194_304):
        for world in (1, 2, 3, 4, 8):
            spans = [shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard(10, 2, 2)


def _worker(rank, world, port, total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard(total, rank, world)
    # stand-in for the per-target result: a pure function of the GLOBAL target index, which is what
    # keying the samplers by (seed, first_index + row) guarantees on the device
    local = torch.arange(lo, hi, dtype=torch.int64) * 3 + 1
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    bufs = [torch.zeros(int(s.item()), dtype=torch.int64) for s in sizes]
    pad = max(int(s.item()) for s in sizes)
    padded = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)]
    mine = torch.zeros(pad, dtype=torch.int64)
    mine[: hi - lo] = local
    dist.all_gather(padded, mine)
    if rank == 0:
        full = torch.cat([p[: int(s.item())] for p, s in zip(padded, sizes)])
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()
    del bufs


def _gather_worker(rank, world, port, total, dim, out_path):
    """The product's own collective (tools_b200.sharding.gather_shards: the call bench.py makes over NCCL) on gloo."""
    from tools_b200.sharding import gather_shards

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard(total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.int64)
    # per-target int16 "preimages" and int64 targets as pure functions of the GLOBAL target index -- what keying the
    # samplers by (seed, first_index + row) gives on the device
    e16 = ((idx[:, None] * 7 + torch.arange(dim)[None, :] * 3) % 2001 - 1000).to(torch.int16)
    u = idx[:, None] * 5 + torch.arange(4)[None, :]
    ge = gather_shards(dist, e16, rank, world, dst=0)
    gu = gather_shards(dist, u, rank, world, dst=0)
    if rank == 0:
        assert ge.dtype == torch.int16 and ge.shape == (total, dim)
        np.save(out_path, ge.numpy())
        np.save(out_path + ".u.npy", gu.numpy())
    else:
        assert ge is None and gu is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_helper(tmp_path):
    """Shards produced per rank (keyed by the global target index) and gathered by the product's helper equal the
    single-process result, in global target order."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total, dim = 512, 24
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_gather_worker, args=(2, port, total, dim, out), nprocs=2, join=True)
    idx = np.arange(total)
    want = ((idx[:, None] * 7 + np.arange(dim)[None, :] * 3) % 2001 - 1000).astype(np.int16)
    assert np.array_equal(np.load(out), want)
    assert np.array_equal(np.load(out + ".u.npy"), idx[:, None] * 5 + np.arange(4)[None, :])


def test_two_rank_gloo_gather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total = 1001
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = np.load(out)
    assert np.array_equal(got, np.arange(total) * 3 + 1)
