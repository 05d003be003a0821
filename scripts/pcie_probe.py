#!/usr/bin/env python3
"""H2D / D2H bandwidth from pinned host memory vs transfer size and stream count
(sizing of the chunks of pk_solve_ik_*_host)."""
import torch

dev = torch.device("cuda", 0)
for mb in [0.4, 0.8, 1.6, 4.7, 16, 64]:
    n = int(mb * 2**20 // 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{direction} {mb:6.1f} MiB  {ms * 1e3:8.1f} us  {n * 4 / ms / 1e6:7.1f} GB/s")
# two streams, both directions at once
n = int(4.7 * 2**20 // 4)
h1, h2 = (torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2))
d1, d2 = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mb_up, mb_dn in [(4.7, 4.7), (4.7, 1.8)]:
    nu, nd = int(mb_up * 2**20 // 4), int(mb_dn * 2**20 // 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(20):
        with torch.cuda.stream(s1):
            d1[:nu].copy_(h1[:nu], non_blocking=True)
        with torch.cuda.stream(s2):
            h2[:nd].copy_(d2[:nd], non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    print(f"duplex {mb_up} MiB up + {mb_dn} MiB down concurrently: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per pair")
