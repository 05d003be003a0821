#!/usr/bin/env python3
"""Per-step timeline of a short burst of BatchedIK.solve_host calls (the driver's --steps 20):
GPU completion time of every step (events on the caller's stream) next to the host time at
which its submission returned.  Shows whether a burst is limited by the host enqueue rate,
by a DMA ramp after idle, or by the link."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pink_b200 import BatchedIK, FrameTask, PostureTask, workloads
from pink_b200.engine import get_engine
from pink_b200.limits import ConfigurationLimit, VelocityLimit
from pink_b200.robots import load_robot_description


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    print("numa:", bench.bind_to_gpu_numa_node(0))
    dev = torch.device("cuda", 0)
    robot = load_robot_description("ur5_description")
    model = robot.model
    eng = get_engine(model, dev)
    B = 65536
    rng = np.random.default_rng(1)
    q = workloads.sample_configurations(eng.table, B, rng)
    qt = workloads.perturb_configurations(eng.table, q, rng)
    oMf, _ = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=dev))
    f = eng.table.frame_names.index("tool0")
    T = oMf[:, f].reshape(B, 12).contiguous()
    ft = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)
    ft.set_target(T)
    pt = PostureTask(cost=1e-3)
    pt.set_target(workloads.ur5_posture_reference(model))
    ik = BatchedIK(model, [ft, pt], workloads.UR5_DT, damping=workloads.UR5_DAMPING,
                   limits=[ConfigurationLimit(model), VelocityLimit(model)], device=dev, batch_size=B)
    NS = 2
    q_h = [torch.as_tensor(q, dtype=torch.float32).pin_memory() for _ in range(NS)]
    t_h = [T.cpu().pin_memory() for _ in range(NS)]
    v_h = [torch.empty((B, 6), dtype=torch.float32).pin_memory() for _ in range(NS)]
    s_h = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(NS)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for k in range(5):
        ik.solve_host(q_h[k % NS], t_h[k % NS], v_h[k % NS], s_h[k % NS])
    torch.cuda.synchronize()
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)

    def link():
        return "gen%d x%d pstate %d sm %d MHz mem %d MHz" % (
            pynvml.nvmlDeviceGetCurrPcieLinkGeneration(hnd), pynvml.nvmlDeviceGetCurrPcieLinkWidth(hnd),
            pynvml.nvmlDeviceGetPerformanceState(hnd), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM))

    gaps = [0.0, 0.0, 0.3, 0.3, 0.0, 0.0, 0.02, 0.02, 1.0, 1.0, 0.0, 0.0]
    do_flush = os.environ.get("BURST_FLUSH", "1") == "1"
    for rep, gap in enumerate(gaps):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        if gap:
            time.sleep(gap)  # idle gap before the burst
        print(f"rep {rep} gap {gap} flush {do_flush}: before: {link()}")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        host = []
        t0 = time.perf_counter()
        ev[0].record()
        for k in range(steps):
            ik.solve_host(q_h[k % NS], t_h[k % NS], v_h[k % NS], s_h[k % NS])
            ev[k + 1].record()
            host.append((time.perf_counter() - t0) * 1e6)
        torch.cuda.synchronize()
        gpu = [ev[0].elapsed_time(e) * 1e3 for e in ev[1:]]
        d = np.diff([0.0] + gpu)
        print(f"        after: {link()}")
        print(f"rep {rep}: total {gpu[-1]:.0f} us = {gpu[-1] / steps:.1f} us/step; per-step gpu deltas:",
              " ".join(f"{x:.0f}" for x in d))
        print("        host submit-return times:", " ".join(f"{x:.0f}" for x in host))


if __name__ == "__main__":
    main()
