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
