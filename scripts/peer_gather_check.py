#!/usr/bin/env python3
"""Multi-rank check of the fused all-gather (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/peer_gather_check.py
Every rank solves its shard of one global batch with the gather fused into the kernel and
compares its full buffer, bit for bit, with an NCCL all-gather of the plain solve and with the
single-call helper solve_ik_all_ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pink_b200 import BatchedIK, FrameTask, PostureTask, parallel, workloads
from pink_b200.engine import get_engine
from pink_b200.robots import load_robot_description


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    robot = load_robot_description("ur5_description") if rank == 0 else None
    model = parallel.broadcast_model(robot.model if rank == 0 else None, dev)
    eng = get_engine(model, dev)
    B = 8192
    rng = np.random.default_rng(5)  # same global batch on every rank
    q = workloads.sample_configurations(eng.table, world * B, rng)
    qt = workloads.perturb_configurations(eng.table, q, rng)
    oMf, _ = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=dev))
    T = oMf[:, eng.table.frame_names.index("tool0")].reshape(world * B, 12).contiguous()
    ft = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)
    ft.set_target(T[:B])
    pt = PostureTask(cost=1e-3)
    pt.set_target(workloads.ur5_posture_reference(model))
    ik = BatchedIK(model, [ft, pt], workloads.UR5_DT, damping=workloads.UR5_DAMPING, device=dev, batch_size=B)
    q_all = torch.as_tensor(q, dtype=torch.float32, device=dev)
    lo, hi = parallel.shard_bounds(world * B)
    v_local, _ = ik.solve(q_all[lo:hi].contiguous(), T[lo:hi].contiguous())
    v_nccl = parallel.all_gather_velocities(v_local)
    peer = parallel.PeerGather(B, 6, dev)
    ok = True
    for it in range(4):
        v_fused, _ = peer.solve(ik, q_all[lo:hi].contiguous(), T[lo:hi].contiguous())
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(v_fused, v_nccl))
    v_one = parallel.solve_ik_all_ranks(ik, q_all, T, gather=peer)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(v_one, v_nccl))
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer gather check:", "OK" if flag.item() else "MISMATCH", "world", world)
    peer.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
