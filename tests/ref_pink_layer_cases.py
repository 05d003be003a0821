"""Seeded cases shared by ``scripts/make_reference_golden.py`` (which runs the REFERENCE's own
Pink-layer Python on them, build container only) and ``tests/test_reference_pink_layer_golden.py``
(which compares the oracle and the kernels with the frozen outputs).  Each case is one of the
scenarios the parity suites already use (tests/helpers.py, tests/extras.py), in product form
(pink_b200 objects) and oracle form (plain records).  The sphere self-collision barrier is part
of the last two cases only: there the reference's ``SelfCollisionBarrier`` runs on the same sphere
pairs through the sphere-only geometry stand-in of oracle/refshim (the real one needs hpp-fcl /
coal meshes)."""

import types

from tests import extras, helpers

NAMES = ["ur5_arm", "ur5_unreachable", "draco3_relative", "g1_com_relative", "ur5_limits_barriers_constraint",
         "g1_coupling_floating_base_limit", "ur5_velocity_tasks", "ur5_all_barriers", "g1_config4_self_collision", "ur5_benchmark_workload"]


def _from_scenario(sc):
    return types.SimpleNamespace(
        model=sc.model, table=sc.table, B=sc.B, q32=sc.q32, q64=sc.q64, dt=sc.dt, damping=sc.damping,
        tasks=sc.tasks, otasks=sc.oracle_tasks, limits=sc.limits, olimits=sc.oracle_limits,
        barriers=[], obarriers=[], constraints=[], oconstraints=[], collision_model=None, safety_break=False)


def _from_extras(sc, self_collision=False):
    """``self_collision``: keep the sphere self-collision barrier (the reference then runs its
    ``SelfCollisionBarrier`` on the same spheres through the sphere-only geometry stand-in)."""
    keep = [k for k, o in enumerate(sc.obarriers) if self_collision or o["type"] != "self_collision"]
    return types.SimpleNamespace(
        model=sc.model, table=sc.table, B=sc.B, q32=sc.q32, q64=sc.q64, dt=sc.dt, damping=sc.damping,
        tasks=sc.tasks, otasks=sc.otasks, limits=sc.limits, olimits=sc.olimits,
        barriers=[sc.barriers[k] for k in keep], obarriers=[sc.obarriers[k] for k in keep],
        constraints=sc.constraints, oconstraints=sc.oconstraints,
        collision_model=sc.collision_model if self_collision else None, safety_break=False)


def _ur5_velocity_tasks(B, seed):
    """FrameTask with anisotropic costs, gain < 1 and LM term + JointVelocityTask + DampingTask;
    ConfigurationLimit with a non-default gain (``pink/tasks/joint_velocity_task.py``,
    ``damping_task.py``, ``limits/configuration_limit.py:40``)."""
    import numpy as np
    import torch

    from pink_b200 import workloads
    from pink_b200.limits import ConfigurationLimit, VelocityLimit
    from pink_b200.tasks import DampingTask, FrameTask, JointVelocityTask

    robot, model, table = helpers.load("ur5_description")
    rng = np.random.default_rng(seed)
    q = workloads.sample_configurations(table, B, rng)
    qt = workloads.perturb_configurations(table, q, rng, sigma=0.25)
    T = helpers.frame_targets(table, qt, "tool0")
    ft = FrameTask("tool0", position_cost=[1.0, 0.5, 2.0], orientation_cost=[0.3, 0.0, 0.7], lm_damping=0.1, gain=0.7)
    ft.set_target(torch.as_tensor(T))
    T64 = T.astype(np.float64)
    oft = {"type": "frame", "frame": table.frame_names.index("tool0"), "cost": np.array(ft.cost), "gain": 0.7, "lm_damping": 0.1,
           "target": (T64[:, :, :3], T64[:, :, 3])}
    dt = 0.01
    target_v = (rng.normal(size=(B, 6)) * 0.4).astype(np.float32)
    jv = JointVelocityTask(cost=0.2)
    jv.set_target(torch.as_tensor(target_v), dt)
    ojv = {"type": "joint_velocity", "cost": 0.2, "gain": 1.0, "lm_damping": 0.0, "target": target_v.astype(np.float64) * dt,
           "ref_class": "joint_velocity", "ref_dt": dt}
    dm = DampingTask(cost=0.3)
    odm = {"type": "joint_velocity", "cost": 0.3, "gain": 1.0, "lm_damping": 0.0, "target": np.zeros(6), "ref_class": "damping"}
    sc = helpers.Scenario("ur5_velocity_tasks", robot, model, table, q, [ft, jv, dm], [oft, ojv, odm], dt, 1e-6, safety_break=False)
    sc.limits = [ConfigurationLimit(model, config_limit_gain=0.3), VelocityLimit(model)]
    sc.oracle_limits = [("configuration", 0.3), ("velocity", None)]
    return sc


def build(name):
    if name == "ur5_benchmark_workload":  # BASELINE config 2 (bench.py's generator and seed) at B = 64
        return _from_scenario(helpers.ur5_scenario(64, "reachable"))
    if name == "ur5_all_barriers":
        return _from_extras(extras.ur5_extras(10, seed=508), self_collision=True)
    if name == "g1_config4_self_collision":  # BASELINE config 4 with its barrier
        return _from_extras(extras.g1_extras(6, seed=509), self_collision=True)
    if name == "ur5_velocity_tasks":
        return _from_scenario(_ur5_velocity_tasks(10, seed=507))
    if name == "ur5_arm":
        return _from_scenario(helpers.ur5_scenario(12, "reachable", seed=501))
    if name == "ur5_unreachable":
        return _from_scenario(helpers.ur5_scenario(12, "unreachable", seed=502))
    if name == "draco3_relative":
        return _from_scenario(helpers.humanoid_scenario("draco3_description", 6, seed=503, with_relative=True))
    if name == "g1_com_relative":
        return _from_scenario(helpers.humanoid_scenario("g1_description", 6, seed=504, with_com=True, with_relative=True))
    if name == "ur5_limits_barriers_constraint":
        return _from_extras(extras.ur5_extras(10, seed=505))
    if name == "g1_coupling_floating_base_limit":
        return _from_extras(extras.g1_extras(6, seed=506))
    raise KeyError(name)


# ---- closed loops of two reference examples on robots the reference vendors ---------------------
# The same function runs with the reference's modules (scripts/make_reference_golden.py) and with
# this package's (tests/test_reference_pink_layer_golden.py): `api` is the module `pink` /
# `pink_b200`, `robot` a RobotWrapper-like (.model, .data, .q0) of the example's URDF.

EXAMPLE_LOOPS = {"double_pendulum": ("double_pendulum.urdf", 300), "one_dof_configuration_limit": ("simple_pendulum.urdf", 400)}


def run_example_loop(name, api, robot, steps):
    """``examples/double_pendulum.py:36-76`` / ``examples/one_dof_configuration_limit.py:50-84`` without
    the visualiser and the rate limiter: returns ``(q[steps + 1, nq], v[steps, nv])``."""
    import numpy as np

    FrameTask, PostureTask = api.tasks.FrameTask, api.tasks.PostureTask
    dt = 0.01  # RateLimiter(frequency=100.0).period
    if name == "double_pendulum":
        tip = FrameTask("link3", position_cost=1.0, orientation_cost=1e-3)
        posture = PostureTask(cost=1e-2)
        tasks = [tip, posture]
        configuration = api.Configuration(robot.model, robot.data, robot.q0)
        for task in tasks:
            task.set_target_from_configuration(configuration)
        tip.transform_target_to_world.translation[2] -= 0.1
    else:
        task = FrameTask("tip", position_cost=1.0, orientation_cost=0.61)
        tasks = [task]
        goal_configuration = api.Configuration(robot.model, robot.data, np.array([5.5]))
        task.set_target_from_configuration(goal_configuration)
        configuration = api.Configuration(robot.model, robot.data, np.array([0.5]))
    qs, vs, t = [np.array(configuration.q, dtype=np.float64)], [], 0.0
    for _ in range(steps):
        if name == "double_pendulum":
            tasks[0].transform_target_to_world.translation[1] = 0.1 * np.sin(t)
        velocity = api.solve_ik(configuration, tasks, dt, solver="quadprog")
        configuration.integrate_inplace(velocity, dt)
        vs.append(np.array(velocity, dtype=np.float64))
        qs.append(np.array(configuration.q, dtype=np.float64))
        t += dt
    return np.array(qs), np.array(vs)
