#!/usr/bin/env python
"""Pinned device->host / host->device copy bandwidth of this box (what bounds the end-to-end samp_p number once the
computation is faster than the copy of its results): python scripts/d2h_probe.py"""
import torch

dev = torch.device("cuda:0")
for mb in (64, 468, 936):
    n = mb * 1024 * 1024 // 2
    d = torch.zeros(n, dtype=torch.int16, device=dev)
    h = torch.empty(n, dtype=torch.int16).pin_memory()
    s = torch.cuda.Stream()
    for direction in ("d2h", "h2d"):
        with torch.cuda.stream(s):
            for _ in range(2):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            for _ in range(5):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            b.record(s)
        s.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{direction} {mb} MiB pinned: {ms:.2f} ms  {n * 2 / ms / 1e6:.1f} GB/s", flush=True)
