#!/usr/bin/env python3
"""BASELINE config 5: UR5 throughput vs TOTAL batch size (1K .. 1M instances), strong
scaling: under torchrun the total batch is cut into equal contiguous shards, one per rank
(one rank per GPU); alone it is the 1-GPU curve.

    python scripts/batch_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/batch_sweep.py

Per total batch size, one JSON line (rank 0): device-resident time per step of the solve
alone and of solve + all-gather of v on every rank (gather fused into the kernel over NVLink
peer memory, wait on a side stream), both as CUDA-graph replays over rotating buffer sets,
median of 5 regions, max over ranks; whole-job IK steps/s and the algorithmic HBM rate per
GPU as a fraction of the measured peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from pink_b200 import BatchedIK, FrameTask, PostureTask, parallel, workloads
from pink_b200.engine import get_engine
from pink_b200.robots import load_robot_description


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, _ = bench.load_peaks()
    robot = load_robot_description("ur5_description") if rank == 0 else None
    model = parallel.broadcast_model(robot.model if rank == 0 else None, dev)
    eng = get_engine(model, dev)
    table = eng.table
    f = table.frame_names.index("tool0")
    steps = 20
    totals = [int(x) for x in os.environ.get("SWEEP_BATCHES", "1024,4096,16384,65536,262144,1048576").split(",")]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(run, regions=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = []
        for _ in range(regions):
            flush.zero_()
            barrier()
            torch.cuda._sleep(300000)
            e0.record()
            run()
            e1.record()
            barrier()
            out.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(out))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for total in totals:
        B = total // world
        if B < 32:
            continue
        nbuf = max(2, min(32, (192 << 20) // (B * 96)))
        rng = np.random.default_rng(workloads.SEED + 1000 * rank)
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

        def step(k):
            i = k % nbuf
            ik.solve(qs[i], ts[i], vs[i], ss[i])

        for k in range(nbuf):
            step(k)
        torch.cuda.synchronize()
        g = bench.GraphedSteps(torch, dev, step, steps)
        g.run()
        ms_solve = timed(g.run) / steps
        line = {"total_batch": total, "n_gpus": world, "batch_per_gpu": B, "solve_us_per_step": ms_solve * 1e3,
                "ik_steps_per_s": total / (ms_solve * 1e-3),
                "hbm_gbs_algorithmic_per_gpu": B * 96 / (ms_solve * 1e-3) / 1e9,
                "hbm_frac_of_measured": B * 96 / (ms_solve * 1e-3) / 1e9 / peak,
                "buffer_sets": nbuf, "nonzero_status": int(sum(int((s != 0).sum().item()) for s in ss))}
        del g
        if world > 1:
            peer = parallel.PeerGather(B, 6, dev, n_buffers=2)
            side = torch.cuda.Stream(dev)

            def fused(k):
                i = k % nbuf
                cur = torch.cuda.current_stream(dev)
                peer.solve(ik, qs[i], ts[i], ss[i], vs[i], wait=False)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    peer.wait()

            gf = bench.GraphedSteps(torch, dev, fused, steps, end=lambda: torch.cuda.current_stream(dev).wait_stream(side))
            gf.run()
            ms_g = timed(gf.run) / steps
            line["solve_plus_gather_us_per_step"] = ms_g * 1e3
            line["solve_plus_gather_ik_steps_per_s"] = total / (ms_g * 1e-3)
            line["gather"] = "fused into the kernel epilogue over NVLink peer memory, wait on a side stream"
            line["spin_timeouts"] = peer.timeouts()
            del gf
            peer.close()
        if rank == 0:
            print(json.dumps(line), flush=True)
        del ik, qs, ts, vs, ss
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
