"""Seeded synthetic inputs of the BASELINE configurations.

Pure numpy sampling (fp64); forward kinematics of the sampled "target
configurations" is done by the caller (the CUDA library in ``bench.py``, the
oracle in the tests) so that this module carries no kinematics of its own.
"""

from __future__ import annotations

import numpy as np

SEED = 20260922

# examples/arm_ur5.py:33-63
UR5_TASKS = dict(frame="tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0, posture_cost=1e-3)
UR5_DT = 1.0 / 200.0
UR5_DAMPING = 1e-12


def ur5_posture_reference(model) -> np.ndarray:
    """``q_ref`` of ``examples/arm_ur5.py:46-51``: pan = lift = elbow = 1."""
    q = np.zeros(model.nq)
    for name in ("shoulder_pan_joint", "shoulder_lift_joint", "elbow_joint"):
        q[model.joints[model.getJointId(name)].idx_q] = 1.0
    return q


def _finite_limits(table):
    rq = 7 if table.free_flyer else 0
    lo = np.array(table.q_min[rq:], dtype=np.float64)
    hi = np.array(table.q_max[rq:], dtype=np.float64)
    lo = np.where(np.isfinite(lo), lo, -np.pi)
    hi = np.where(np.isfinite(hi), hi, np.pi)
    return lo, hi


def random_quaternions(n: int, rng) -> np.ndarray:
    """Uniform unit quaternions ``[x, y, z, w]``."""
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def sample_configurations(table, B: int, rng, near_limit_fraction: float = 0.05) -> np.ndarray:
    """``q ~ U(0.9 q_min, 0.9 q_max)`` per joint, about 5 % of the instances
    pushed to within 1e-2 rad of a position limit on one joint; free-flyer:
    position ``U([-0.5, 0.5]^3)``, uniform random orientation."""
    lo, hi = _finite_limits(table)
    nj = lo.shape[0]
    qj = rng.uniform(0.9 * lo, 0.9 * hi, size=(B, nj))
    n_near = int(round(near_limit_fraction * B))
    if n_near and nj:
        rows = rng.choice(B, size=n_near, replace=False)
        cols = rng.integers(0, nj, size=n_near)
        upper = rng.random(n_near) < 0.5
        gap = rng.uniform(0.0, 1e-2, size=n_near)
        qj[rows, cols] = np.where(upper, hi[cols] - gap, lo[cols] + gap)
    if not table.free_flyer:
        return qj
    base = np.concatenate([rng.uniform(-0.5, 0.5, size=(B, 3)), random_quaternions(B, rng)], axis=1)
    return np.concatenate([base, qj], axis=1)


def perturb_configurations(table, q: np.ndarray, rng, sigma: float = 0.3) -> np.ndarray:
    """``q + delta`` with ``delta ~ N(0, sigma^2)`` on the joints, clipped into the
    limits; the frame poses at these configurations are the *reachable* targets."""
    lo, hi = _finite_limits(table)
    rq = 7 if table.free_flyer else 0
    out = np.array(q, dtype=np.float64)
    out[:, rq:] = np.clip(out[:, rq:] + rng.normal(0.0, sigma, size=out[:, rq:].shape), lo, hi)
    return out


def random_poses(B: int, rng, half_extent: float = 1.2) -> np.ndarray:
    """*Unreachable* targets: ``p ~ U([-h, h]^3)``, ``R`` uniform on SO(3);
    returns ``[B, 3, 4]`` rows ``[R | p]``."""
    q = random_quaternions(B, rng)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((B, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    p = rng.uniform(-half_extent, half_extent, size=(B, 3))
    return np.concatenate([R, p[:, :, None]], axis=2)


# ---- humanoid configurations (BASELINE configs 3 and 4, without the barrier) -----------

# examples/humanoid_draco3.py:69-91 (frame, position_cost, orientation_cost)
DRACO3_FRAME_TASKS = [
    ("l_foot_contact", 1.0, 1.0),
    ("torso_com_link", 1.0, 0.0),
    ("r_foot_contact", 1.0, 1.0),
    ("r_hand_contact", 4.0, 4.0),
]
DRACO3_POSTURE_COST, DRACO3_DAMPING = 1e-1, 1e-12

# examples/humanoid_g1_com.py:45-74, 106-113
G1_FRAME_TASKS = [
    ("pelvis", 0.0, 10.0),
    ("right_ankle_roll_link", [2.0, 2.0, 200.0], 10.0),
    ("left_ankle_roll_link", [2.0, 2.0, 200.0], 10.0),
    ("right_wrist_yaw_link", 4.0, 0.0),
    ("left_wrist_yaw_link", 4.0, 0.0),
]
G1_POSTURE_COST, G1_COM_COST, G1_DAMPING = 1e-1, 200.0, 0.01


def humanoid_task_set(name: str, model, oMf_target, com_target=None):
    """Task objects of the Draco3 / G1 examples with per-instance targets.

    ``oMf_target`` is ``[B, nframes, 3, 4]`` (frame poses at the target
    configurations, e.g. from ``Engine.forward_kinematics``); ``com_target``
    ``[B, 3]`` for the G1 set.  Returns ``(tasks, damping)``.
    """
    from .tasks import ComTask, FrameTask, PostureTask

    if name.startswith("draco3"):
        specs, posture_cost, damping = DRACO3_FRAME_TASKS, DRACO3_POSTURE_COST, DRACO3_DAMPING
    else:
        specs, posture_cost, damping = G1_FRAME_TASKS, G1_POSTURE_COST, G1_DAMPING
    tasks = []
    for frame, pc, oc in specs:
        t = FrameTask(frame, position_cost=pc, orientation_cost=oc)
        t.set_target(oMf_target[:, model.getFrameId(frame)])
        tasks.append(t)
    posture = PostureTask(cost=posture_cost)
    q_ref = np.zeros(model.nq)
    q_ref[6] = 1.0
    posture.set_target(q_ref)
    tasks.append(posture)
    if not name.startswith("draco3") and com_target is not None:
        com = ComTask(cost=G1_COM_COST)
        com.set_target(com_target)
        tasks.append(com)
    return tasks, damping


# sphere set of BASELINE config 4 (joint, radius): 9 spheres -> 36 pairs, the 8 closest enter
# the barrier (gain 20, safe displacement gain 1, d_min 0.05 as
# examples/barriers/kukas_self_collision.py:167-172)
G1_COLLISION_SPHERES = [
    ("left_wrist_yaw_joint", 0.06), ("right_wrist_yaw_joint", 0.06), ("left_elbow_joint", 0.06),
    ("right_elbow_joint", 0.06), ("waist_yaw_joint", 0.13), ("left_knee_joint", 0.07),
    ("right_knee_joint", 0.07), ("left_ankle_roll_joint", 0.06), ("right_ankle_roll_joint", 0.06),
]
HUMANOID_BYTES_PER_STEP = {"draco3_description": 460, "g1_description": 536}
HUMANOID_BATCH = {"draco3_description": 32768, "g1_description": 16384}      # BASELINE configs 3, 4


def humanoid_problem(name: str, device, batch: int | None = None, with_barrier: bool = False, seed: int = SEED):
    """BASELINE config 3 (``draco3_description``) / config 4 (``g1_description``, optionally
    with the sphere self-collision barrier) as a prepared :class:`BatchedIK` plus its seeded
    device inputs.  Returns ``(ik, q [B, nq], targets [B, stride], model)``."""
    import torch

    import pink_b200
    from .engine import get_engine
    from .model import JointModelFreeFlyer
    from .robots import load_robot_description

    robot = load_robot_description(name, root_joint=JointModelFreeFlyer())
    model = robot.model
    barriers, cm = None, None
    if with_barrier:
        cm = pink_b200.SphereCollisionModel(model)
        for k, (joint, radius) in enumerate(G1_COLLISION_SPHERES):
            cm.add_sphere(f"s{k}", model.getJointId(joint), (0.0, 0.0, 0.0), radius)
        cm.add_all_collision_pairs()
        barriers = [pink_b200.barriers.SelfCollisionBarrier(8, gain=20.0, safe_displacement_gain=1.0, d_min=0.05)]
    eng = get_engine(model, device)
    B = int(batch or HUMANOID_BATCH[name])
    rng = np.random.default_rng(seed)
    q = sample_configurations(eng.table, B, rng)
    qt = perturb_configurations(eng.table, q, rng, sigma=0.15)
    q_d = torch.as_tensor(q, dtype=torch.float32, device=device)
    oMf, com = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=device), want_com=True)
    tasks, damping = humanoid_task_set(name, model, oMf, com)
    ik = pink_b200.BatchedIK(model, tasks, 1.0 / 200.0, damping=damping, safety_break=False, device=device,
                             batch_size=B, barriers=barriers, collision_model=cm)
    targets = torch.cat([t._pk_describe(model)["target"].to(device) for t in tasks
                         if isinstance(t._pk_describe(model)["target"], torch.Tensor)], dim=1).contiguous()
    ik._bench_tasks = (tasks, damping, barriers, cm)  # kept for checks that rebuild the QP
    return ik, q_d, targets, model
