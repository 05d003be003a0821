"""Limit- and barrier-level behaviour the reference's own unit tests pin
(``tests/test_velocity_limit.py``, ``test_configuration_limit.py``, ``test_limits.py``,
``test_acceleration_limit.py``, ``test_floating_base_velocity_limit.py``,
``test_barrier.py``, ``test_position_barrier.py``, ``test_body_spherical_barrier.py``,
``test_self_collision_barrier.py``), on the unbatched drop-in classes with the engine
routed to the host build of the kernels (tests/host_engine.py; test harness only)."""

import numpy as np
import pytest

import pink_b200
from pink_b200 import Configuration, FrameTask, SphereCollisionModel, solve_ik
from pink_b200.barriers import BodySphericalBarrier, PositionBarrier, SelfCollisionBarrier
from pink_b200.exceptions import (InvalidCollisionPairs, NegativeMinimumDistance, NoPositionLimitProvided,
                                  PinkError)
from pink_b200.limits import AccelerationLimit, ConfigurationLimit, FloatingBaseVelocityLimit, VelocityLimit
from pink_b200.model import JointModelFreeFlyer, Model, model_from_urdf_string
from pink_b200.robots import load_robot_description
from pink_b200.utils import get_joint_idx, process_collision_pairs
from tests.host_engine import host_engine  # noqa: F401  (fixture)


@pytest.fixture(autouse=True)
def _cpu(host_engine):  # noqa: F811
    yield


@pytest.fixture
def humanoid():
    robot = load_robot_description("g1_description", root_joint=JointModelFreeFlyer())
    return robot, Configuration(robot.model, robot.data, robot.q0)


@pytest.fixture
def arm():
    robot = load_robot_description("ur5_description")
    return robot, Configuration(robot.model, robot.data, np.array([0.3, -1.0, 1.2, -0.4, 0.5, 0.1]))


# ---- VelocityLimit / ConfigurationLimit (tests/test_velocity_limit.py, test_configuration_limit.py, test_limits.py)

@pytest.mark.parametrize("cls", [VelocityLimit, ConfigurationLimit])
def test_limit_dimensions(humanoid, cls):
    robot, _ = humanoid
    limit = cls(robot.model)
    for joint in limit.joints:
        assert joint.idx_q >= 0 and joint.idx_v >= 0
    nb = len(limit.joints)
    assert len(limit.indices) == nb and limit.projection_matrix.shape == (nb, robot.model.nv)
    assert len(cls(Model()).indices) == 0  # unbounded models don't fail


def test_velocity_limit_argument(arm):
    """tests/test_velocity_limit.py:46-63."""
    robot, _ = arm
    nv = robot.model.nv
    limit = VelocityLimit(robot.model, velocity_limit=2.0 * np.ones(nv))
    assert len(limit.indices) == nv
    _, h = limit.compute_qp_inequalities(configuration=None, dt=1e-3)
    assert np.allclose(h, 1e-3 * 2.0)
    with pytest.raises(PinkError):
        VelocityLimit(robot.model, velocity_limit=np.ones(nv + 1))


def test_configuration_limit_far_and_near(humanoid):
    """tests/test_configuration_limit.py:48-99."""
    robot, configuration = humanoid
    dt = 1e-3
    G, h = ConfigurationLimit(robot.model).compute_qp_inequalities(configuration, dt=dt)
    v_lim = np.where(np.isfinite(robot.model.velocityLimit), robot.model.velocityLimit, 0.0)
    assert np.max(+G @ v_lim * dt - h) < -1e-10 and np.max(-G @ v_lim * dt - h) < -1e-10
    slack_vel = 5.5e-4
    robot.model.lowerPositionLimit = configuration.integrate(-slack_vel * configuration.tangent.ones, dt)
    robot.model.upperPositionLimit = configuration.integrate(+slack_vel * configuration.tangent.ones, dt)
    G, h = ConfigurationLimit(robot.model, config_limit_gain=0.5).compute_qp_inequalities(configuration, dt)
    assert np.max(h) < slack_vel * dt + 1e-9 and np.min(h) > -slack_vel * dt - 1e-9


def test_limitless_joints_are_skipped():
    """tests/test_limits.py:48-66, tests/test_configuration_limit.py:101-104: joints without
    a position range carry no configuration rows."""
    urdf = """
    <robot name="mixed">
      <link name="a"/><link name="b"/><link name="c"/><link name="d"/>
      <joint name="j1" type="continuous"><parent link="a"/><child link="b"/><axis xyz="0 0 1"/></joint>
      <joint name="j2" type="revolute"><parent link="b"/><child link="c"/><origin xyz="0.2 0 0"/><axis xyz="0 1 0"/>
        <limit lower="-1" upper="1" velocity="2"/></joint>
      <joint name="j3" type="continuous"><parent link="c"/><child link="d"/><origin xyz="0.2 0 0"/><axis xyz="1 0 0"/>
        <limit velocity="3"/></joint>
    </robot>"""
    model = model_from_urdf_string(urdf)
    assert tuple(ConfigurationLimit(model).indices) == (1,)
    assert tuple(VelocityLimit(model).indices) == (1, 2)


def test_velocity_without_configuration_limits(arm):
    """tests/test_limits.py:68-90: with only the velocity limit the IK may leave the
    configuration range."""
    robot, configuration = arm
    task = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0)
    task.set_target_from_configuration(configuration)
    v = solve_ik(configuration, [task], dt=1e-3, solver="daqp", limits=[robot.model.velocity_limit])
    assert np.allclose(v, 0.0, atol=1e-4)


# ---- AccelerationLimit (tests/test_acceleration_limit.py) -----------------------------------------------

def test_acceleration_limit_dimensions_and_empty_model(arm):
    robot, _ = arm
    a_max = 14.0 * np.ones(robot.model.nv)
    limit = AccelerationLimit(robot.model, a_max)
    assert limit.projection_matrix.shape == (len(limit.indices), robot.model.nv)
    empty = Model()
    empty_limit = AccelerationLimit(empty, np.empty(0))
    assert len(empty_limit.indices) == 0


def test_continuous_joint_has_no_braking_distance():
    """tests/test_acceleration_limit.py:55-94."""
    urdf = """
    <robot name="continuous_joint_robot">
      <link name="base_link"/><link name="link1"/>
      <joint name="joint1" type="continuous"><parent link="base_link"/><child link="link1"/><axis xyz="0 0 1"/></joint>
    </robot>"""
    model = model_from_urdf_string(urdf)
    dt = 5e-3
    limit = AccelerationLimit(model, np.array([14.0]))
    configuration = Configuration(model, model.createData(), np.zeros(model.nq))
    limit.set_last_integration(np.array([3.0]), dt)
    G, h = limit.compute_qp_inequalities(configuration, dt)
    nb = len(limit.indices)
    assert nb == 1 and np.all(-h[nb:] <= h[:nb])


def test_acceleration_limit_has_an_effect(arm):
    """tests/test_acceleration_limit.py:96-151, including the in-place edit of the task's
    target through ``transform_target_to_world``."""
    robot, _ = arm
    a_max = 14.0 * np.ones(robot.model.nv)
    limit = AccelerationLimit(robot.model, a_max)
    task = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0)
    configuration = Configuration(robot.model, robot.data, robot.q0 + np.array([0.0, -1.2, 1.5, -0.3, 0.8, 0.0]))
    task.set_target_from_configuration(configuration)
    target = task.transform_target_to_world
    y0 = target.translation[1]
    target.translation[1] = y0 + 0.05
    assert np.linalg.norm(task.compute_error(configuration)) > 0.04  # the edit is seen
    limits = [robot.model.configuration_limit, robot.model.velocity_limit]
    dt = 5e-3
    v_prev = solve_ik(configuration, [task], dt, solver="scs", limits=limits)
    configuration.integrate_inplace(v_prev, dt)
    limit.set_last_integration(v_prev, dt)
    target.translation[1] = y0 + 0.35
    v_with = solve_ik(configuration, [task], dt, solver="scs", limits=limits + [limit])
    v_without = solve_ik(configuration, [task], dt, solver="scs", limits=limits)
    a_with = (v_with - v_prev) / dt
    a_without = (v_without - v_prev) / dt
    assert np.all(np.abs(a_with) < a_max + 2e-2), np.abs(a_with).max()
    assert not np.all(np.abs(a_without) < a_max + 2e-2)


# ---- FloatingBaseVelocityLimit (tests/test_floating_base_velocity_limit.py:84-164) ---------------------------

def test_floating_base_velocity_limit(humanoid):
    robot, configuration = humanoid
    model = robot.model
    dt = 0.05
    linear_max = np.array([0.3, 0.3, 0.2])
    angular_max = np.array([1.0, 1.0, 1.5])
    root_joint_id = model.getJointId("root_joint")
    base_frame = next(f.name for f in model.frames if f.parentJoint == root_joint_id)
    limit = FloatingBaseVelocityLimit(model=model, base_frame=base_frame, max_linear_velocity=linear_max,
                                      max_angular_velocity=angular_max)
    G, h = limit.compute_qp_inequalities(configuration, dt)
    assert G.shape == (12, model.nv) and h.shape == (12,)
    _, idx_v = get_joint_idx(model, "root_joint")
    assert np.linalg.norm(G[:, idx_v:idx_v + 6]) > 0.0
    assert np.allclose(np.delete(G, np.s_[idx_v:idx_v + 6], axis=1), 0.0)
    dq = np.zeros(model.nv)
    dq[idx_v:idx_v + 6] = 0.999 * dt * np.hstack([linear_max, angular_max])
    assert np.all(G @ dq <= h + 1e-7)
    dq[idx_v] = 1.01 * dt * linear_max[0]
    assert np.any(G @ dq > h + 1e-7)
    # an infinite bound contributes no row
    partial = FloatingBaseVelocityLimit(model=model, base_frame=base_frame,
                                        max_linear_velocity=[0.4, 0.2, np.inf], max_angular_velocity=[np.inf, np.inf, 1.0])
    G, h = partial.compute_qp_inequalities(configuration, dt)
    assert G.shape == (6, model.nv)
    # default frame detection and the missing-root error
    auto = FloatingBaseVelocityLimit(model=model, base_frame=None, max_linear_velocity=linear_max,
                                     max_angular_velocity=angular_max)
    assert model.frames[auto.frame_id].parentJoint == root_joint_id
    fixed = load_robot_description("ur5_description")
    with pytest.raises(ValueError):
        FloatingBaseVelocityLimit(model=fixed.model, base_frame="base_link", max_linear_velocity=linear_max,
                                  max_angular_velocity=angular_max)


# ---- Barrier base class through PositionBarrier (tests/test_barrier.py) -----------------------------------------

def test_barrier_shapes_and_objective(humanoid):
    robot, conf = humanoid
    nv = robot.model.nv
    barrier = PositionBarrier("left_hip_pitch_link", p_min=np.zeros(3), p_max=np.zeros(3))
    H, c = barrier.compute_qp_objective(conf)
    G, h = barrier.compute_qp_inequalities(conf, 1e-3)
    assert H.shape == (nv, nv) and c.shape == (nv,) and G.shape == (barrier.dim, nv) and h.shape == (barrier.dim,)
    assert barrier.compute_barrier(conf).shape == (barrier.dim,)
    assert barrier.compute_jacobian(conf).shape == (barrier.dim, nv)
    assert np.allclose(H, 0.0) and np.allclose(c, 0.0)  # no penalty weight
    weighted = PositionBarrier("left_hip_pitch_link", p_min=np.zeros(3), p_max=np.zeros(3), safe_displacement_gain=1.0)
    H, c = weighted.compute_qp_objective(conf)
    assert not np.allclose(H, 0.0)
    r = repr(PositionBarrier("universe", safe_displacement_gain=0.0, p_min=np.zeros(3)))
    for field in ("gain=", "safe_displacement=", "safe_displacement_gain", "dim"):
        assert field in r


# ---- PositionBarrier (tests/test_position_barrier.py) -------------------------------------------------------------

def test_position_barrier(arm):
    _, configuration = arm
    with pytest.raises(NoPositionLimitProvided):
        PositionBarrier("foo")
    assert PositionBarrier("tool0", p_min=np.zeros(3)).dim == 3
    assert PositionBarrier("tool0", p_max=np.zeros(3)).dim == 3
    assert PositionBarrier("tool0", p_min=np.zeros(3), p_max=np.zeros(3)).dim == 6
    assert PositionBarrier("tool0", p_min=np.zeros(3), gain=1).gain.shape == (3,)
    assert PositionBarrier("tool0", p_max=np.zeros(3), gain=np.array([1, 2, 3])).gain.shape == (3,)
    assert PositionBarrier("tool0", p_min=np.zeros(3), p_max=np.zeros(3), gain=1).gain.shape == (6,)
    assert PositionBarrier("tool0", p_min=np.zeros(3), p_max=np.zeros(3), gain=np.array([1, 2, 3])).gain.shape == (6,)
    p = configuration.get_transform_frame_to_world("tool0").translation
    h = PositionBarrier("tool0", p_min=p - 0.5).compute_barrier(configuration)
    assert np.all(h > 0)
    for violated in range(3):
        p_min = p - 0.5
        p_min[violated] = p[violated] + 1.0
        assert np.any(PositionBarrier("tool0", p_min=p_min).compute_barrier(configuration) < 0)


# ---- BodySphericalBarrier (tests/test_body_spherical_barrier.py) ---------------------------------------------------

def test_body_spherical_barrier(humanoid):
    robot, configuration = humanoid
    ees = ("left_wrist_yaw_link", "right_wrist_yaw_link")
    with pytest.raises(NegativeMinimumDistance):
        BodySphericalBarrier(ees, d_min=-1)
    assert BodySphericalBarrier(ees, d_min=0.2).dim == 1
    assert BodySphericalBarrier(ees, d_min=0.2).gain.shape == (1,)
    pa = configuration.get_transform_frame_to_world(ees[0]).translation
    pb = configuration.get_transform_frame_to_world(ees[1]).translation
    d = np.linalg.norm(pa - pb)
    barrier = BodySphericalBarrier(ees, d_min=0.5 * d)
    J = barrier.compute_jacobian(configuration)
    assert np.asarray(J).reshape(-1).shape[0] == robot.model.nv
    assert barrier.compute_barrier(configuration)[0] > 0
    assert BodySphericalBarrier(ees, d_min=1.5 * d).compute_barrier(configuration)[0] < 0
    # h = |pa - pb|^2 - d_min^2 (pink/barriers/body_spherical_barrier.py:73-100)
    assert abs(barrier.compute_barrier(configuration)[0] - (d * d - 0.25 * d * d)) < 1e-5


# ---- SelfCollisionBarrier (tests/test_self_collision_barrier.py) -----------------------------------------------------

SPHERE_ARM = """
<robot name="sphere_arm">
  <link name="base"><collision><origin xyz="0 0 0.05"/><geometry><sphere radius="0.08"/></geometry></collision></link>
  <link name="l1"><collision><origin xyz="0 0 0.15"/><geometry><sphere radius="0.06"/></geometry></collision></link>
  <link name="l2"><collision><origin xyz="0 0 0.15"/><geometry><sphere radius="0.05"/></geometry></collision>
                  <collision><origin xyz="0 0 0.30"/><geometry><sphere radius="0.05"/></geometry></collision></link>
  <link name="l3"><collision><origin xyz="0 0 0.15"/><geometry><sphere radius="0.05"/></geometry></collision>
                  <collision><origin xyz="0 0 0.30"/><geometry><sphere radius="0.04"/></geometry></collision></link>
  <joint name="j1" type="revolute"><parent link="base"/><child link="l1"/><origin xyz="0 0 0.1"/><axis xyz="0 0 1"/>
    <limit lower="-3" upper="3" velocity="2"/></joint>
  <joint name="j2" type="revolute"><parent link="l1"/><child link="l2"/><origin xyz="0 0 0.3"/><axis xyz="0 1 0"/>
    <limit lower="-3" upper="3" velocity="2"/></joint>
  <joint name="j3" type="revolute"><parent link="l2"/><child link="l3"/><origin xyz="0 0 0.4"/><axis xyz="0 1 0"/>
    <limit lower="-3" upper="3" velocity="2"/></joint>
</robot>"""


@pytest.fixture
def sphere_arm():
    model = model_from_urdf_string(SPHERE_ARM)
    collision_model = SphereCollisionModel.from_urdf_string(model, SPHERE_ARM)
    collision_data = process_collision_pairs(model, collision_model)
    return model, collision_model, collision_data


def test_self_collision_barrier(sphere_arm):
    model, collision_model, collision_data = sphere_arm
    n_pairs = len(collision_model.collisionPairs)
    assert n_pairs == 13  # 6 spheres, the two same-link pairs excluded
    with pytest.raises(NegativeMinimumDistance):
        SelfCollisionBarrier(n_collision_pairs=10, d_min=-1)
    with pytest.raises(InvalidCollisionPairs):
        SelfCollisionBarrier(n_collision_pairs=-1, d_min=0.02)
    configuration = Configuration(model, model.createData(), np.zeros(model.nq), collision_model=collision_model,
                                  collision_data=collision_data)
    with pytest.raises(InvalidCollisionPairs):
        SelfCollisionBarrier(n_collision_pairs=n_pairs + 5, d_min=0.02).compute_barrier(configuration)
    assert SelfCollisionBarrier(n_collision_pairs=10, d_min=0.02).dim == 10
    assert SelfCollisionBarrier(n_collision_pairs=10, d_min=0.02).gain.shape == (10,)
    # the ABI carries at most 24 inequality rows per barrier set; all pairs still fit here
    barrier = SelfCollisionBarrier(n_collision_pairs=n_pairs, d_min=0.02)
    J = barrier.compute_jacobian(configuration)
    assert J.ndim == 2 and J.shape == (n_pairs, model.nv)
    h = barrier.compute_barrier(configuration)
    assert h.shape == (n_pairs,)
    # stretched out: every non-adjacent pair is clear of the margin
    far = [k for k, (i, j) in enumerate(collision_model.collisionPairs)
           if abs(collision_model.parents[i] - collision_model.parents[j]) > 1]
    assert far and np.all(np.sort(h)[-len(far):] > 0)
    # folded: the last link's spheres come down onto the base
    folded = Configuration(model, model.createData(), np.array([0.0, 2.6, 2.6]), collision_model=collision_model,
                           collision_data=collision_data)
    assert np.min(barrier.compute_barrier(folded)) < 0
    # closest pairs are the ones kept (tests/test_self_collision_barrier.py:139-159)
    few = SelfCollisionBarrier(n_collision_pairs=5, d_min=0.02)
    h_few = few.compute_barrier(configuration)
    assert h_few.shape == (5,)
    for h_i in h_few:
        assert np.sum(h < h_i - 1e-6) < few.dim
    # collision data is shared and follows the configuration constructed last, as pin.GeometryData does
    configuration = Configuration(model, model.createData(), np.zeros(model.nq), collision_model=collision_model,
                                  collision_data=collision_data)
    distances = np.array([r.min_distance for r in configuration.collision_data.distanceResults]) - 0.02
    assert configuration.collision_data.enable_contact and distances.shape == (n_pairs,)
    assert np.allclose(np.sort(distances), np.sort(h), atol=1e-5)
    for h_i in h_few:
        assert np.sum(distances < h_i - 1e-6) < few.dim
    r0 = configuration.collision_data.distanceResults[0]
    assert abs(np.linalg.norm(r0.getNearestPoint2() - r0.getNearestPoint1()) - abs(r0.min_distance)) < 1e-9
    assert collision_model.geometryObjects[0].parentJoint == collision_model.parents[0]
    assert pink_b200.__version__


def test_opposing_barrier_rows_pin_a_direction(humanoid):
    """A position barrier with p_min == p_max (the frame held on a plane) gives two
    opposing rows: the QP keeps the frame in the plane and moves it within
    (the degenerate lo == hi case itself is pinned in test_hostsim_dualqp.py)."""
    robot, configuration = humanoid
    frame = "left_wrist_yaw_link"
    p = configuration.get_transform_frame_to_world(frame).translation
    task = FrameTask(frame, position_cost=1.0, orientation_cost=0.0)
    T = configuration.get_transform_frame_to_world(frame)
    T.translation[:] = p + np.array([0.05, 0.02, 0.08])
    task.set_target(T)
    barrier = PositionBarrier(frame, indices=[2], p_min=np.array([p[2]]), p_max=np.array([p[2]]), gain=1.0)
    dt = 5e-3
    limits = [robot.model.configuration_limit, robot.model.velocity_limit]
    for force_generic in (False, True):
        import os
        if force_generic:
            os.environ["PK_TREE"] = "0"
        try:
            v = solve_ik(configuration, [task], dt, solver="quadprog", barriers=[barrier], limits=limits)
        finally:
            os.environ.pop("PK_TREE", None)
        J = configuration.get_frame_jacobian(frame)
        R = configuration.get_transform_frame_to_world(frame).rotation
        world_linear = R @ (J[:3] @ v)
        assert np.linalg.norm(world_linear[:2]) > 1e-2  # moves in the plane
        assert abs(world_linear[2]) < 1e-3 * max(1.0, np.linalg.norm(world_linear))  # not out of it
