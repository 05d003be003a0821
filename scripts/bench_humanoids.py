#!/usr/bin/env python3
"""Timing of the humanoid configurations (BASELINE configs 3 / 4 without the barrier)
through the current general path: prints one JSON line per configuration."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pink_b200
from pink_b200 import workloads
from pink_b200.engine import get_engine
from pink_b200.model import JointModelFreeFlyer
from pink_b200.robots import load_robot_description

BYTES = {"draco3_description": 460, "g1_description": 536}
BATCH = {"draco3_description": 32768, "g1_description": 16384}


def main():
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    for name, config4 in [("draco3_description", False), ("g1_description", False), ("g1_description", True)]:
        robot = load_robot_description(name, root_joint=JointModelFreeFlyer())
        model = robot.model
        barriers, cm = None, None
        if config4:
            # BASELINE config 4: + sphere-pair self-collision barrier (gain 20, safe displacement
            # gain 1, d_min 0.05 as examples/barriers/kukas_self_collision.py:167-172)
            cm = pink_b200.SphereCollisionModel(model)
            for k, (joint, radius) in enumerate([
                ("left_wrist_yaw_joint", 0.06), ("right_wrist_yaw_joint", 0.06), ("left_elbow_joint", 0.06),
                ("right_elbow_joint", 0.06), ("waist_yaw_joint", 0.13), ("left_knee_joint", 0.07),
                ("right_knee_joint", 0.07), ("left_ankle_roll_joint", 0.06), ("right_ankle_roll_joint", 0.06)]):
                cm.add_sphere(f"s{k}", model.getJointId(joint), (0.0, 0.0, 0.0), radius)
            cm.add_all_collision_pairs()
            barriers = [pink_b200.barriers.SelfCollisionBarrier(8, gain=20.0, safe_displacement_gain=1.0, d_min=0.05)]
        eng = get_engine(model, dev)
        B = int(os.environ.get("PK_HUMANOID_BATCH", BATCH[name]))
        rng = np.random.default_rng(workloads.SEED)
        q = workloads.sample_configurations(eng.table, B, rng)
        qt = workloads.perturb_configurations(eng.table, q, rng, sigma=0.15)
        q_d = torch.as_tensor(q, dtype=torch.float32, device=dev)
        oMf, com = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=dev), want_com=True)
        tasks, damping = workloads.humanoid_task_set(name, model, oMf, com)
        ik = pink_b200.BatchedIK(model, tasks, 1.0 / 200.0, damping=damping, safety_break=False, device=dev, batch_size=B,
                                 barriers=barriers, collision_model=cm)
        targets = torch.cat([t._pk_describe(model)["target"].to(dev) for t in tasks
                             if isinstance(t._pk_describe(model)["target"], torch.Tensor)], dim=1).contiguous()
        v = torch.empty((B, model.nv), dtype=torch.float32, device=dev)
        st = torch.empty((B,), dtype=torch.int32, device=dev)
        for _ in range(3):
            ik.solve(q_d, targets, v, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        e0.record()
        for _ in range(steps):
            ik.solve(q_d, targets, v, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ach = B * BYTES[name] / (ms * 1e-3) / 1e9
        print(json.dumps({
            "config": name + (" (synthetic) + self-collision barrier (36 sphere pairs, 8 closest), tree kernel with the warp-cooperative dual QP"
                              if config4 else " (synthetic), tree kernel"), "batch": B, "nv": model.nv, "ms_per_step": ms,
            "ik_steps_per_s": B / (ms * 1e-3), "hbm_gbs_algorithmic": ach, "hbm_frac_of_measured": ach / peak,
            "status_counts": {int(k): int(c) for k, c in zip(*np.unique(st.cpu().numpy(), return_counts=True))},
        }))


if __name__ == "__main__":
    main()
