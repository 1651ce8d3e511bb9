"""Target sharding across GPUs: targets are independent given the key, so each rank takes a
contiguous slice of the global target index range and the key is replicated (no collective on the
data path).  Passing the slice start as `first_index` keeps the Philox streams aligned with the
single-GPU run."""


def shard(total: int, rank: int, world: int):
    """Contiguous slice [start, stop) of range(total) for `rank` of `world`; sizes differ by <= 1."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_shards(dist, local, rank, world, dst=0, group=None):
    """ONE gather of equally shaped per-rank tensors to `dst` (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU
    tests).  Returns the concatenation along dim 0 on `dst` (rank order = global target order, because shard() hands out
    contiguous slices in rank order), None elsewhere."""
    # a gather only moves bytes: int16 (which neither NCCL nor gloo has as a reduction type) travels as uint8
    import torch

    rows = local.shape[0]
    wire = local.contiguous().view(torch.uint8).reshape(rows, -1) if local.dtype == torch.int16 else local.contiguous()
    if rank == dst:
        out = wire.new_empty((world * rows,) + tuple(wire.shape[1:]))
        dist.gather(wire, list(out.split(rows)), dst=dst, group=group)
        if wire is not local and local.dtype == torch.int16:
            out = out.view(torch.int16).reshape((world * rows,) + tuple(local.shape[1:]))
        return out
    dist.gather(wire, None, dst=dst, group=group)
    return None


def gather_domain_i16(torch, dist, lib, ffi, e, u, rank, world, dev, stream, max_bytes=64 << 30, dst=0):
    """Final gather of the per-rank Domain results over NVLink: each rank narrows its int32 shard to int16
    on the device (qf_narrow_i32_i16_dev -- |e_i| <= 6 s r fits for the reference's parameter sets; overflow is detected and
    reported) and ONE NCCL gather brings the shards to `dst`; the targets travel the same way (as int64) so that `dst` can
    verify the gathered preimages.  Device-timed on the library's stream.  Returns a dict {narrow_ms, gather_ms, bytes, GBps,
    ...}; on `dst` also "gathered": (e_all int16 [world * B, ...], u_all).  Skipped (returns {"skipped": why}) when the gathered
    tensor would not fit max_bytes on `dst` -- C4's 4 Mi x 32849 preimages are 276 GB: those results stay sharded."""
    b = e.shape[0]
    count = e.numel()
    total_bytes = world * count * 2
    if total_bytes > max_bytes:
        return {"skipped": f"gathered int16 results would be {total_bytes / 1e9:.1f} GB: kept sharded"}
    with torch.cuda.stream(stream):
        e16 = torch.empty(e.shape, dtype=torch.int16, device=dev)
        ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        # untimed warm-up of the collective (NCCL sets up its NVLink channels lazily, on the first call of a kind)
        gather_shards(dist, e16[: min(b, 128)].contiguous(), rank, world, dst)
        stream.synchronize()
        dist.barrier()
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ev0.record(stream)
        st = lib.qf_narrow_i32_i16_dev(ffi.ptr(e.data_ptr()), ffi.ptr(e16.data_ptr()), count, ffi.ptr(ovf.data_ptr()),
                                       ffi.ptr(stream.cuda_stream))
        assert st == 0
        ev1.record(stream)
        out_e = gather_shards(dist, e16, rank, world, dst)
        ev2.record(stream)
        out_u = gather_shards(dist, u, rank, world, dst)
        stream.synchronize()
    assert int(ovf.item()) == 0, "a preimage entry does not fit int16: gather the int32 form instead"
    ms_narrow, ms_gather = ev0.elapsed_time(ev1), ev1.elapsed_time(ev2)
    recv = (world - 1) * count * 2
    info = {"collective": "NCCL gather to rank %d (NVLink / NVSwitch)" % dst, "dtype": "int16", "ranks": world,
            "bytes_received_by_dst": recv, "narrow_ms": ms_narrow, "gather_ms": ms_gather,
            "GBps_into_dst": recv / ms_gather / 1e6 if ms_gather > 0 else None,
            "per_target_bytes": (count // b) * 2}
    if rank == dst:
        info["gathered"] = (out_e, out_u)
    return info
