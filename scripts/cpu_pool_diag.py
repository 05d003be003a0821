#!/usr/bin/env python3
"""Diagnostics of the CPU arm's thread pool on the GPU box (no GPU work): per-step times for
several thread counts and process set-ups."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
print("mode", mode, "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except OSError as e:
    print("cpu.max: n/a", e)
print(open("/proc/self/status").read().split("Cpus_allowed_list:")[1].split("\n")[0].strip())
if mode == "setaff":
    os.sched_setaffinity(0, os.sched_getaffinity(0))
if mode == "cuda":
    import torch
    torch.cuda.init()
    torch.zeros(1, device="cuda")
if mode == "node1":
    os.sched_setaffinity(0, set(range(32, 64)) | set(range(96, 128)))
    os.sched_setaffinity(0, set(range(128)))
import bench

arm = bench.CpuArm(65536)
for th in (128, 64, 32, 16, 128):
    arm.step(th)
    ts = [arm.step(th)[0] * 1e3 for _ in range(12)]
    print(f"threads {th:4d}: ms per step", " ".join(f"{t:.1f}" for t in ts), f" -> {65536 / (sorted(ts)[len(ts)//2] * 1e-3):.3e} steps/s")
print("loadavg", open("/proc/loadavg").read().strip())
