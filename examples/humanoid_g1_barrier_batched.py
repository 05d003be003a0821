#!/usr/bin/env python3
"""A batch of floating-base humanoids (G1-class tree) reaching with both hands while a
position barrier keeps the left hand below a ceiling and a spherical barrier keeps the two
hands apart - the batched form of the reference's ``examples/barriers`` scripts.

    python examples/humanoid_g1_barrier_batched.py --batch 4096 --steps 100
"""

import argparse

import numpy as np
import torch

import pink_b200 as pink
from pink_b200 import solve_ik
from pink_b200.barriers import BodySphericalBarrier, PositionBarrier
from pink_b200.model import JointModelFreeFlyer
from pink_b200.robots import load_robot_description
from pink_b200.tasks import FrameTask, PostureTask

HANDS = ("left_wrist_yaw_link", "right_wrist_yaw_link")
FEET = ("left_ankle_roll_link", "right_ankle_roll_link")


def run(batch: int = 256, steps: int = 50, device: str = "cuda", dt: float = 5e-3, seed: int = 0):
    """Returns ``(lowest ceiling margin seen, smallest hand distance seen, final hand error)``."""
    robot = load_robot_description("g1_description", root_joint=JointModelFreeFlyer())
    model = robot.model
    rng = np.random.default_rng(seed)
    q0 = np.tile(robot.q0, (batch, 1))
    q0[:, 7:] += 0.02 * rng.standard_normal((batch, model.nq - 7))
    configuration = pink.Configuration(model, robot.data, torch.as_tensor(q0, dtype=torch.float32, device=device))

    tasks = []
    for frame in FEET + ("pelvis",):
        task = FrameTask(frame, position_cost=10.0, orientation_cost=1.0)
        task.set_target_from_configuration(configuration)  # hold feet and pelvis where they are
        tasks.append(task)
    hand_tasks = []
    for frame, side in zip(HANDS, (+1.0, -1.0)):
        task = FrameTask(frame, position_cost=4.0, orientation_cost=0.0, lm_damping=1e-2)
        target = configuration.get_transform_frame_to_world(frame)
        target[:, 2, 3] += 0.25                                # up ...
        target[:, 1, 3] -= side * 0.20                         # ... and towards the other hand
        task.set_target(target)
        tasks.append(task)
        hand_tasks.append(task)
    posture = PostureTask(cost=1e-2)
    posture.set_target(robot.q0)
    tasks.append(posture)

    start = configuration.get_transform_frame_to_world(HANDS[0])
    ceiling = float(start[:, 2, 3].max()) + 0.10               # the left hand may rise 10 cm, not 25
    barriers = [
        PositionBarrier(HANDS[0], indices=[2], p_max=np.array([ceiling]), gain=5.0),
        BodySphericalBarrier(HANDS, d_min=0.12, gain=5.0),
    ]

    margin, closest = float("inf"), float("inf")
    for _ in range(steps):
        velocity = solve_ik(configuration, tasks, dt, solver="quadprog", barriers=barriers, safety_break=False)
        configuration.integrate_inplace(velocity, dt)
        left = configuration.get_transform_frame_to_world(HANDS[0])[:, :, 3]
        right = configuration.get_transform_frame_to_world(HANDS[1])[:, :, 3]
        margin = min(margin, float((ceiling - left[:, 2]).min()))
        closest = min(closest, float(torch.linalg.norm(left - right, dim=1).min()))
    error = torch.stack([torch.linalg.norm(t.compute_error(configuration)[:, :3], dim=1) for t in hand_tasks]).max()
    return margin, closest, float(error)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    margin, closest, error = run(args.batch, args.steps)
    print(f"{args.batch} humanoids, {args.steps} steps: ceiling margin >= {margin:.4f} m, hands never closer than "
          f"{closest:.4f} m (d_min 0.12), remaining hand error {error:.3f} m (the barriers hold the hands back)")
