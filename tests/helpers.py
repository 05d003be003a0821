"""Shared fixtures: seeded workloads, product-side task objects and the matching
oracle task dicts, tolerance checks."""

import numpy as np
import torch

from oracle import ik as oik
from oracle import kinematics as okin
from pink_b200 import ComTask, FrameTask, PostureTask, RelativeFrameTask, workloads
from pink_b200.limits import ConfigurationLimit, VelocityLimit
from pink_b200.model import JointModelFreeFlyer
from pink_b200.robots import load_robot_description

# fp32 CUDA vs fp64 oracle on identical inputs
V_ATOL, V_RTOL = 2e-4, 2e-3


def load(name):
    root = None if name.startswith("ur5") else JointModelFreeFlyer()
    robot = load_robot_description(name, root_joint=root)
    return robot, robot.model, robot.model.table()


def frame_targets(table, q_target, frame_name):
    """[B, 3, 4] fp32 poses of `frame_name` at configurations `q_target` (oracle FK)."""
    fk = okin.forward_kinematics(table, q_target)
    R, p = okin.frame_placement(table, fk, table.frame_names.index(frame_name))
    return np.concatenate([R, p[:, :, None]], axis=2).astype(np.float32)


class Scenario:
    """A (model, tasks, limits, inputs) bundle in both product and oracle form."""

    def __init__(self, name, robot, model, table, q, tasks, oracle_tasks, dt, damping, limits="default",
                 safety_break=True):
        self.name, self.robot, self.model, self.table = name, robot, model, table
        self.q32 = np.ascontiguousarray(q, dtype=np.float32)
        self.q64 = self.q32.astype(np.float64)
        self.tasks, self.oracle_tasks = tasks, oracle_tasks
        self.dt, self.damping, self.safety_break = dt, damping, safety_break
        if limits == "default":
            self.limits = [ConfigurationLimit(model), VelocityLimit(model)]
            self.oracle_limits = None
        else:
            self.limits = []
            self.oracle_limits = []

    @property
    def B(self):
        return self.q32.shape[0]

    def problem(self):
        from pink_b200.solve_ik import describe_problem

        prob, parts, descs = describe_problem(self.model, self.B, self.tasks, self.dt, self.damping, self.limits,
                                              self.safety_break)
        targets = torch.cat([p.cpu().float() for p in parts], dim=1).numpy() if parts else None
        return prob, targets, descs

    def oracle_solve(self, n=None):
        n = self.B if n is None else min(n, self.B)
        tasks = [oik._slice_task_range(t, 0, n) for t in self.oracle_tasks]
        return oik.solve_ik_batch(self.table, self.q64[:n], tasks, self.dt, self.damping, self.oracle_limits,
                                  self.safety_break)

    def oracle_build(self):
        return oik.build_ik(self.table, self.q64, self.oracle_tasks, self.dt, self.damping, self.oracle_limits)


def ur5_scenario(B, kind="reachable", seed=workloads.SEED, lm_damping=1.0, posture_cost=1e-3, out_of_limits=0):
    robot, model, table = load("ur5_description")
    rng = np.random.default_rng(seed)
    q = workloads.sample_configurations(table, B, rng)
    if out_of_limits:
        rows = rng.choice(B, size=out_of_limits, replace=False)
        q[rows, 2] = table.q_max[2] + rng.uniform(1e-3, 0.2, size=out_of_limits)
    if kind == "reachable":
        T = frame_targets(table, workloads.perturb_configurations(table, q, rng), "tool0")
    elif kind == "unreachable":
        T = workloads.random_poses(B, rng).astype(np.float32)
    else:  # at target: zero error
        T = frame_targets(table, q.astype(np.float32).astype(np.float64), "tool0")
    q_ref = workloads.ur5_posture_reference(model)
    ft = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=lm_damping)
    ft.set_target(torch.as_tensor(T))
    pt = PostureTask(cost=posture_cost)
    pt.set_target(q_ref)
    T64 = T.astype(np.float64)
    f = table.frame_names.index("tool0")
    otasks = [
        {"type": "frame", "frame": f, "cost": np.ones(6), "gain": 1.0, "lm_damping": lm_damping,
         "target": (T64[:, :, :3], T64[:, :, 3])},
        {"type": "posture", "cost": posture_cost, "gain": 1.0, "lm_damping": 0.0, "target": q_ref},
    ]
    return Scenario(f"ur5-{kind}", robot, model, table, q, [ft, pt], otasks, workloads.UR5_DT, workloads.UR5_DAMPING)


def humanoid_scenario(name, B, seed=workloads.SEED, sigma=0.15, with_com=False, with_relative=False):
    """Draco3-class (examples/humanoid_draco3.py:69-91) or G1-class
    (examples/humanoid_g1_com.py:45-74) task sets on the synthetic trees."""
    robot, model, table = load(name)
    rng = np.random.default_rng(seed)
    q = workloads.sample_configurations(table, B, rng, near_limit_fraction=0.05)
    qt = workloads.perturb_configurations(table, q, rng, sigma=sigma)
    if name.startswith("draco3"):
        specs = [("l_foot_contact", 1.0, 1.0), ("torso_com_link", 1.0, 0.0), ("r_foot_contact", 1.0, 1.0),
                 ("r_hand_contact", 4.0, 4.0)]
        posture_cost, damping = 1e-1, 1e-12
    else:
        specs = [("pelvis", 0.0, 10.0), ("right_ankle_roll_link", [2.0, 2.0, 200.0], 10.0),
                 ("left_ankle_roll_link", [2.0, 2.0, 200.0], 10.0), ("right_wrist_yaw_link", 4.0, 0.0),
                 ("left_wrist_yaw_link", 4.0, 0.0)]
        posture_cost, damping = 1e-1, 0.01
    tasks, otasks = [], []
    for frame, pc, oc in specs:
        T = frame_targets(table, qt, frame)
        t = FrameTask(frame, position_cost=pc, orientation_cost=oc)
        t.set_target(torch.as_tensor(T))
        tasks.append(t)
        T64 = T.astype(np.float64)
        otasks.append({"type": "frame", "frame": table.frame_names.index(frame), "cost": np.array(t.cost),
                       "gain": 1.0, "lm_damping": 0.0, "target": (T64[:, :, :3], T64[:, :, 3])})
    q_ref = q[0].copy()
    pt = PostureTask(cost=posture_cost)
    pt.set_target(q_ref)
    tasks.append(pt)
    otasks.append({"type": "posture", "cost": posture_cost, "gain": 1.0, "lm_damping": 0.0, "target": q_ref})
    if with_com:
        com = okin.center_of_mass(table, okin.forward_kinematics(table, qt)).astype(np.float32)
        ct = ComTask(cost=200.0)
        ct.set_target(torch.as_tensor(com))
        tasks.append(ct)
        otasks.append({"type": "com", "cost": np.full(3, 200.0), "gain": 1.0, "lm_damping": 0.0,
                       "target": com.astype(np.float64)})
    if with_relative:
        a, b = ("l_hand_contact", "r_hand_contact") if name.startswith("draco3") else ("left_wrist_yaw_link", "right_wrist_yaw_link")
        fkt = okin.forward_kinematics(table, qt)
        Ra, pa = okin.frame_placement(table, fkt, table.frame_names.index(a))
        Rb, pb = okin.frame_placement(table, fkt, table.frame_names.index(b))
        from oracle import lie

        Rr, pr = lie.se3_act_inv(Rb, pb, Ra, pa)
        T = np.concatenate([Rr, pr[:, :, None]], axis=2).astype(np.float32)
        rt = RelativeFrameTask(a, b, position_cost=2.0, orientation_cost=0.5, lm_damping=1e-3, gain=0.8)
        rt.set_target(torch.as_tensor(T))
        tasks.append(rt)
        T64 = T.astype(np.float64)
        otasks.append({"type": "relative_frame", "frame": table.frame_names.index(a), "root": table.frame_names.index(b),
                       "cost": np.array(rt.cost), "gain": 0.8, "lm_damping": 1e-3,
                       "target": (T64[:, :, :3], T64[:, :, 3])})
    return Scenario(name, robot, model, table, q, tasks, otasks, 1.0 / 200.0, damping, safety_break=False)


def within_tolerance(v, v_ref, atol=V_ATOL, rtol=V_RTOL):
    """Per-instance boolean: every coordinate within atol + rtol |v_ref|."""
    err = np.abs(np.asarray(v, dtype=np.float64) - v_ref)
    return (err <= atol + rtol * np.abs(v_ref)).all(axis=-1)


# Parity statement of BASELINE.md section 6, by conditioning of the QP's Hessian.  The kernels
# work on the square-root form [diag(d); A] (cond ~ sqrt(cond(H))), so fp32 keeps the standard
# tolerance up to cond(H) ~ 1e5; the G1 example weights (CoM cost 200 next to posture cost 0.1)
# give cond(H) of 1e6..1e7, where the measured distribution (host build and GPU, 1000-instance
# samples) is 98.9 % inside the standard tolerance, 99.8 % inside 2.5x of it, all inside 5x.
PARITY_BINS = [
    # (cond(H) lower edge, upper edge, [(atol, rtol, minimum fraction inside), ...])
    (0.0, 1e5, [(V_ATOL, V_RTOL, 0.999)]),
    (1e5, 1e9, [(V_ATOL, V_RTOL, 0.975), (5e-4, 5e-3, 0.99), (1e-3, 1e-2, 1.0)]),
]


def parity_by_condition(v, v_ref, H, min_bin=20):
    """Checks the binned parity statement; returns a report (list of dicts, one per
    populated bin) with the outliers of the standard tolerance listed.  Bins with fewer than
    ``min_bin`` instances are checked against the loosest bound of their bin only (a fraction
    of a handful of instances is not a statistic)."""
    v = np.asarray(v, dtype=np.float64)
    cond = np.linalg.cond(np.asarray(H, dtype=np.float64))
    err = np.abs(v - v_ref)
    report = []
    for lo, hi, bounds in PARITY_BINS:
        m = (cond >= lo) & (cond < hi)
        if not m.any():
            continue
        entry = {"cond": (lo, hi), "n": int(m.sum()), "fractions": [], "worst_abs_err": float(err[m].max())}
        checks = bounds if m.sum() >= min_bin else bounds[-1:]
        for atol, rtol, need in checks:
            ok = (err[m] <= atol + rtol * np.abs(v_ref[m])).all(axis=-1)
            need_n = need if m.sum() >= min_bin else (1.0 if need == 1.0 else 0.0)
            entry["fractions"].append((atol, rtol, float(ok.mean()), need))
            assert ok.mean() >= need_n, (
                f"cond(H) in [{lo:g}, {hi:g}): {ok.mean():.4f} of {m.sum()} inside {atol:g} + {rtol:g}|v|, "
                f"need {need}; worst |dv| {err[m].max():.3e}")
        std = (err[m] <= V_ATOL + V_RTOL * np.abs(v_ref[m])).all(axis=-1)
        entry["outliers_of_standard_tolerance"] = [int(i) for i in np.nonzero(m)[0][~std]][:32]
        report.append(entry)
    assert (cond < PARITY_BINS[-1][1]).all(), "cond(H) beyond the stated range"
    return report


def random_chain_model(nj, rng, prismatic=(), name="chain"):
    """Fixed-base serial chain with random placements / axes, a tool frame on the last
    joint and an elbow frame mid-chain (exercises NJ = 2..7, prismatic joints, two
    frame tasks and the zero columns of a mid-chain frame)."""
    from pink_b200.model import Model, SE3
    from oracle import lie

    model = Model(name)
    parent = 0
    for j in range(nj):
        R, _ = lie.exp6(np.concatenate([np.zeros(3), rng.normal(size=3) * 0.8]))
        T = SE3(R, rng.uniform(-0.3, 0.3, size=3))
        kind = "prismatic" if j in prismatic else "revolute"
        lim = 0.4 if kind == "prismatic" else 2.5
        parent = model.add_joint(f"j{j}", parent, T, rng.normal(size=3), kind=kind, lower=-lim, upper=lim,
                                 velocity=2.0 + j)
        model.append_inertia(parent, 1.0 + 0.1 * j, rng.uniform(-0.1, 0.1, size=3))
    model.add_frame("tool", parent, SE3(np.eye(3), np.array([0.05, 0.0, 0.1])))
    model.add_frame("elbow", max(1, nj // 2), SE3(np.eye(3), np.array([0.0, 0.02, 0.0])))
    return model


def chain_scenario(nj, B, seed=1, prismatic=(), two_tasks=False, shared_target=False):
    rng = np.random.default_rng(seed)
    model = random_chain_model(nj, rng, prismatic)
    table = model.table()
    q = workloads.sample_configurations(table, B, rng)
    qt = workloads.perturb_configurations(table, q, rng, sigma=0.2)
    tasks, otasks = [], []
    frames = [("tool", 1.0, 0.7)] + ([("elbow", [0.5, 0.0, 2.0], 0.3)] if two_tasks else [])
    for frame, pc, oc in frames:
        T = frame_targets(table, qt, frame)
        t = FrameTask(frame, position_cost=pc, orientation_cost=oc, lm_damping=0.1, gain=0.9)
        if shared_target:
            from pink_b200.model import SE3

            T = np.broadcast_to(T[:1], T.shape).copy()
            t.set_target(SE3(T[0].astype(np.float64)))
        else:
            t.set_target(torch.as_tensor(T))
        tasks.append(t)
        T64 = T.astype(np.float64)
        otasks.append({"type": "frame", "frame": table.frame_names.index(frame), "cost": np.array(t.cost),
                       "gain": 0.9, "lm_damping": 0.1, "target": (T64[:, :, :3], T64[:, :, 3])})
    q_ref = np.zeros(nj)
    pt = PostureTask(cost=0.05, gain=0.5)
    pt.set_target(q_ref)
    tasks.append(pt)
    otasks.append({"type": "posture", "cost": 0.05, "gain": 0.5, "lm_damping": 0.0, "target": q_ref})

    class _R:
        data = None

    return Scenario(f"chain{nj}", _R(), model, table, q, tasks, otasks, 0.01, 1e-8)


def random_tree_model(nj, rng, free_flyer=False, name="tree"):
    """Random joint tree (each joint hangs under a random earlier joint, or the root),
    optional free-flyer root, a few frames on leaves: exercises models beyond the
    32-joint warp kernel up to the C-ABI maximum (58 joints, nv = 64)."""
    from pink_b200.model import Model, SE3
    from oracle import lie

    model = Model(name, free_flyer=free_flyer)
    root = model.root_body  # property: Pinocchio id of the root body (1 with a free-flyer)
    ids = []
    for j in range(nj):
        parent = root if (j == 0 or rng.random() < 0.15) else ids[int(rng.integers(max(0, j - 6), j))]
        R, _ = lie.exp6(np.concatenate([np.zeros(3), rng.normal(size=3) * 0.8]))
        T = SE3(R, rng.uniform(-0.15, 0.15, size=3))
        kind = "prismatic" if rng.random() < 0.1 else "revolute"
        lim = 0.3 if kind == "prismatic" else 2.0
        jid = model.add_joint(f"j{j}", parent, T, rng.normal(size=3), kind=kind, lower=-lim, upper=lim, velocity=3.0)
        model.append_inertia(jid, 0.5 + 0.05 * j, rng.uniform(-0.05, 0.05, size=3))
        ids.append(jid)
    for k, jid in enumerate(ids[-4:]):
        model.add_frame(f"tip{k}", jid, SE3(np.eye(3), rng.uniform(-0.1, 0.1, size=3)))
    return model


def tree_scenario(nj, B, free_flyer=False, seed=3):
    """Frame tasks on the four tip frames + relative frame + posture (+ CoM) on a random tree."""
    rng = np.random.default_rng(seed)
    model = random_tree_model(nj, rng, free_flyer)
    table = model.table()
    q = workloads.sample_configurations(table, B, rng)
    qt = workloads.perturb_configurations(table, q, rng, sigma=0.2)
    tasks, otasks = [], []
    for k, (pc, oc) in enumerate([(1.0, 1.0), (2.0, 0.0), ([1.0, 0.0, 3.0], 0.5)]):
        frame = f"tip{k}"
        T = frame_targets(table, qt, frame)
        t = FrameTask(frame, position_cost=pc, orientation_cost=oc, lm_damping=0.05)
        t.set_target(torch.as_tensor(T))
        tasks.append(t)
        T64 = T.astype(np.float64)
        otasks.append({"type": "frame", "frame": table.frame_names.index(frame), "cost": np.array(t.cost), "gain": 1.0,
                       "lm_damping": 0.05, "target": (T64[:, :, :3], T64[:, :, 3])})
    q_ref = q[0].copy()
    pt = PostureTask(cost=0.05)
    pt.set_target(q_ref)
    tasks.append(pt)
    otasks.append({"type": "posture", "cost": 0.05, "gain": 1.0, "lm_damping": 0.0, "target": q_ref})
    com = okin.center_of_mass(table, okin.forward_kinematics(table, qt)).astype(np.float32)
    ct = ComTask(cost=5.0)
    ct.set_target(torch.as_tensor(com))
    tasks.append(ct)
    otasks.append({"type": "com", "cost": np.full(3, 5.0), "gain": 1.0, "lm_damping": 0.0, "target": com.astype(np.float64)})

    class _R:
        data = None

    return Scenario(f"tree{nj}", _R(), model, table, q, tasks, otasks, 0.01, 1e-6, safety_break=False)
