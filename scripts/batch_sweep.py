#!/usr/bin/env python3
"""BASELINE config 5: UR5 throughput vs batch size (1K .. 1M instances) on one GPU;
under torchrun every rank runs its shard of the same total batch.  One JSON line per
batch size: device-resident kernel time (CUDA-graph replay over rotating buffer sets
larger than L2 where the batch allows) and the algorithmic HBM rate."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pink_b200 import BatchedIK, FrameTask, PostureTask, workloads
from pink_b200.engine import get_engine
from pink_b200.robots import load_robot_description


def main():
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    robot = load_robot_description("ur5_description")
    model = robot.model
    eng = get_engine(model, dev)
    table = eng.table
    f = table.frame_names.index("tool0")
    for B in [1024, 4096, 16384, 65536, 262144, 1048576]:
        nbuf = max(2, min(32, (192 << 20) // (B * 96)))
        rng = np.random.default_rng(workloads.SEED)
        qs, ts, vs, ss = [], [], [], []
        for _ in range(nbuf):
            q = workloads.sample_configurations(table, B, rng)
            qt = workloads.perturb_configurations(table, q, rng)
            oMf, _ = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=dev))
            qs.append(torch.as_tensor(q, dtype=torch.float32, device=dev))
            ts.append(oMf[:, f].reshape(B, 12).contiguous())
            vs.append(torch.empty((B, 6), dtype=torch.float32, device=dev))
            ss.append(torch.empty((B,), dtype=torch.int32, device=dev))
        ft = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)
        ft.set_target(ts[0])
        pt = PostureTask(cost=1e-3)
        pt.set_target(workloads.ur5_posture_reference(model))
        ik = BatchedIK(model, [ft, pt], workloads.UR5_DT, damping=workloads.UR5_DAMPING, device=dev, batch_size=B)
        for k in range(nbuf):
            ik.solve(qs[k], ts[k], vs[k], ss[k])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for k in range(nbuf):
                    ik.solve(qs[k], ts[k], vs[k], ss[k])
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        reps = max(3, int(2e8 // (B * nbuf)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * nbuf)
        bad = int(sum(int((s != 0).sum().item()) for s in ss))
        print(json.dumps({"batch": B, "kernel_us": ms * 1e3, "ik_steps_per_s": B / (ms * 1e-3),
                          "hbm_gbs_algorithmic": B * 96 / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": B * 96 / (ms * 1e-3) / 1e9 / peak,
                          "buffer_sets": nbuf, "working_set_mib": nbuf * B * 96 / 2**20, "nonzero_status": bad}), flush=True)
        del graph, ik, qs, ts, vs, ss
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
