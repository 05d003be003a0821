"""Task-level behaviour the reference's own unit tests pin (``tests/test_frame_task.py``,
``test_posture_task.py``, ``test_com_task.py``, ``test_relative_frame_task.py``,
``test_joint_velocity_task.py``, ``test_damping_task.py``, ``test_low_acceleration_task.py``,
``test_linear_holonomic_task.py``, ``test_joint_coupling_task.py``), on the unbatched
drop-in classes with the engine routed to the host build of the kernels
(tests/host_engine.py; test harness only).  Tolerances are those of fp32 kernels."""

import numpy as np
import pytest

from pink_b200 import (ComTask, Configuration, DampingTask, FrameTask, JointCouplingTask, JointVelocityTask,
                       LinearHolonomicTask, LowAccelerationTask, PostureTask, RelativeFrameTask)
from pink_b200.exceptions import TargetNotSet, TaskDefinitionError, TaskJacobianNotSet
from pink_b200.model import Frame, JointModelFreeFlyer
from pink_b200.robots import load_robot_description
from pink_b200.spatial import SE3
from pink_b200.utils import get_joint_idx
from tests.host_engine import host_engine  # noqa: F401  (fixture)

ATOL = 2e-5


@pytest.fixture(autouse=True)
def _cpu(host_engine):  # noqa: F811
    yield


@pytest.fixture
def humanoid():
    """G1-class floating-base tree with a user-added operational frame
    (tests/test_frame_task.py:28-48)."""
    robot = load_robot_description("g1_description", root_joint=JointModelFreeFlyer())
    joint_name = robot.model.names[-1]
    robot.model.addFrame(Frame("ee_frame", robot.model.getJointId(joint_name), SE3.Identity(), "OP_FRAME"))
    robot.data = robot.model.createData()
    return Configuration(robot.model, robot.data, robot.q0)


@pytest.fixture
def arm():
    robot = load_robot_description("ur5_description")
    return Configuration(robot.model, robot.data, np.array([0.3, -1.0, 1.2, -0.4, 0.5, 0.1]))


def objective_value(H, c, qd):
    return qd @ H @ qd + c @ qd


# ---- FrameTask (tests/test_frame_task.py) ------------------------------------------------

def test_frame_target_from_configuration_and_copy(humanoid):
    task = FrameTask("left_ankle_roll_link", position_cost=1.0, orientation_cost=0.1)
    with pytest.raises(TargetNotSet):
        task.compute_error(humanoid)
    task.set_target_from_configuration(humanoid)
    T = humanoid.get_transform_frame_to_world("left_ankle_roll_link")
    assert np.allclose(np.asarray(T), np.asarray(task.transform_target_to_world))
    target = SE3(np.eye(3), np.array([0.0, 1.0, 0.0]))
    task.set_target(target)
    target.translation[1] += 12.0
    assert abs(task.transform_target_to_world.translation[1] - 1.0) < 1e-12
    r = repr(task)
    for field in ("frame=", "gain=", "orientation_cost=", "position_cost=", "lm_damping="):
        assert field in r  # tests/test_frame_task.py:63-72


@pytest.mark.parametrize("frame", ["right_ankle_roll_link", "ee_frame"])
def test_frame_zero_error_when_target_at_body(humanoid, frame):
    """Also on a user-added OP_FRAME (tests/test_frame_task.py:112-121, 217-226)."""
    task = FrameTask(frame, [1.0, 1.0, 1.0], [1.0, 1.0, 1.0])
    task.set_target(humanoid.get_transform_frame_to_world(frame))
    e = task.compute_error(humanoid)
    J = task.compute_jacobian(humanoid)
    assert e.shape == (6,) and J.shape == (6, humanoid.model.nv)
    assert np.linalg.norm(e) < 1e-6
    assert np.allclose(J, -humanoid.get_frame_jacobian(frame), atol=ATOL)


def test_frame_unit_cost_qp_objective(humanoid):
    frame = "right_wrist_yaw_link"
    task = FrameTask(frame, position_cost=1.0, orientation_cost=0.1)
    task.set_target(humanoid.get_transform_frame_to_world(frame) * SE3(np.eye(3), np.array([0.0, 0.01, 0.0])))
    J = task.compute_jacobian(humanoid)
    e = task.compute_error(humanoid)
    task.set_position_cost(1.0)
    task.set_orientation_cost(1.0)
    task.lm_damping = 0.0
    H, c = task.compute_qp_objective(humanoid)
    assert np.allclose(J.T @ J, H, atol=ATOL)
    assert np.allclose(e.T @ J, c, atol=ATOL)


def test_frame_zero_costs_same_as_disabling_lines(humanoid):
    frame = "left_wrist_yaw_link"
    task = FrameTask(frame, position_cost=1.0, orientation_cost=0.1)
    task.set_target(humanoid.get_transform_frame_to_world(frame) * SE3(np.eye(3), np.array([0.1, 0.02, 0.01])))
    J = task.compute_jacobian(humanoid)
    e = task.compute_error(humanoid)
    qd = np.random.default_rng(0).random(J.shape[1])
    cases = [(1.0, 0.0, slice(0, 3)), (0.0, 1.0, slice(3, 6)),
             ([1.0, 0.0, 0.0], 0.0, slice(0, 1)), ([0.0, 1.0, 0.0], 0.0, slice(1, 2)),
             ([0.0, 0.0, 1.0], 0.0, slice(2, 3)), (0.0, [1.0, 0.0, 0.0], slice(3, 4)),
             (0.0, [0.0, 1.0, 0.0], slice(4, 5)), (0.0, [0.0, 0.0, 1.0], slice(5, 6))]
    for position_cost, orientation_cost, rows in cases:
        task.set_position_cost(position_cost)
        task.set_orientation_cost(orientation_cost)
        task.lm_damping = 0.0
        H, c = task.compute_qp_objective(humanoid)
        expect = objective_value(J[rows].T @ J[rows], e[rows].T @ J[rows], qd)
        assert abs(objective_value(H, c, qd) - expect) < 1e-4 * max(1.0, abs(expect))


def test_frame_lm_damping(humanoid):
    """No effect at the target, a damping under error (tests/test_frame_task.py:183-215)."""
    frame = "left_wrist_yaw_link"
    task = FrameTask(frame, position_cost=1.0, orientation_cost=0.1)
    task.set_target(humanoid.get_transform_frame_to_world(frame))
    task.lm_damping = 1e-8
    H_1, c_1 = task.compute_qp_objective(humanoid)
    task.lm_damping = 1e-4
    H_2, c_2 = task.compute_qp_objective(humanoid)
    assert np.allclose(H_1, H_2, atol=1e-7) and np.allclose(c_1, c_2, atol=1e-7)
    task.set_target(humanoid.get_transform_frame_to_world(frame) * SE3(np.eye(3), np.array([0.0, 2.0, 0.0])))
    task.lm_damping = 1e-8
    H_1, c_1 = task.compute_qp_objective(humanoid)
    task.lm_damping = 1e-2
    H_2, c_2 = task.compute_qp_objective(humanoid)
    mu = np.diag(H_2 - H_1)
    assert np.all(mu > 0.0) and np.allclose(mu, mu[0], rtol=1e-3)
    assert np.allclose(c_1, c_2, atol=ATOL)
    # mu = lm_damping * |W e|^2 (pink/tasks/task.py:160-164)
    e = task.compute_error(humanoid)
    w = np.array([1.0, 1.0, 1.0, 0.1, 0.1, 0.1])
    assert abs(mu[0] - (1e-2 - 1e-8) * np.sum((w * e) ** 2)) < 1e-3 * mu[0] + 1e-7


def test_frame_inconsistent_cost(humanoid):
    task = FrameTask("ee_frame", [1.0, 1.0, 1.0], [1.0, 1.0, 1.0])
    task.cost = 42.0
    with pytest.raises(TaskDefinitionError):
        task.set_position_cost(1.0)
    with pytest.raises(TaskDefinitionError):
        task.set_orientation_cost(1.0)


# ---- PostureTask (tests/test_posture_task.py) ---------------------------------------------

def test_posture_task_semantics(humanoid):
    task = PostureTask(cost=1.0)
    assert "cost=" in repr(task) and "gain=" in repr(task)
    with pytest.raises(TargetNotSet):
        task.compute_error(humanoid)
    task.set_target_from_configuration(humanoid)
    assert np.allclose(task.target_q, humanoid.q)
    q = humanoid.q.copy()
    task.set_target(q)
    q[7] += 12.0
    assert abs(task.target_q[7] - humanoid.q[7]) < 1e-12
    e = task.compute_error(humanoid)
    J = task.compute_jacobian(humanoid)
    nv = humanoid.model.nv
    assert e.shape == (nv - 6,) and J.shape == (nv - 6, nv)
    assert np.linalg.norm(e) < 1e-7
    # unit cost: (J^T J, e^T J) (tests/test_posture_task.py:76-89)
    target = humanoid.q.copy()
    target[7:] += 0.1
    task.set_target(target)
    e = task.compute_error(humanoid)
    assert np.allclose(e, -0.1, atol=1e-6)
    H, c = task.compute_qp_objective(humanoid)
    assert np.allclose(J.T @ J, H, atol=ATOL) and np.allclose(e.T @ J, c, atol=ATOL)
    # zero cost disables the task (tests/test_posture_task.py:91-100)
    task.cost = 0.0
    H, c = task.compute_qp_objective(humanoid)
    qd = np.random.default_rng(1).random(nv)
    assert abs(objective_value(H, c, qd)) < 1e-9


# ---- ComTask (tests/test_com_task.py) ---------------------------------------------------------

def test_com_task_semantics(humanoid):
    task = ComTask(cost=1.0)
    assert "cost=" in repr(task) and "target_com=" in repr(task)
    with pytest.raises(TargetNotSet):
        task.compute_error(humanoid)
    with pytest.raises(TargetNotSet):
        task.compute_jacobian(humanoid)
    task.set_target_from_configuration(humanoid)
    com = humanoid.get_center_of_mass()
    assert np.allclose(task.target_com, com)
    target = np.array(com)
    task.set_target(target)
    target[1] += 12.0
    assert abs(task.target_com[1] - com[1]) < 1e-12
    e = task.compute_error(humanoid)
    J = task.compute_jacobian(humanoid)
    assert e.shape == (3,) and J.shape == (3, humanoid.model.nv) and np.linalg.norm(e) < 1e-6
    task.set_target(com + np.array([0.0, 0.01, 0.0]))
    e = task.compute_error(humanoid)
    assert np.allclose(e, [0.0, -0.01, 0.0], atol=1e-6)
    H, c = task.compute_qp_objective(humanoid)
    assert np.allclose(J.T @ J, H, atol=ATOL) and np.allclose(e.T @ J, c, atol=ATOL)
    zero = ComTask(cost=0.0)
    zero.set_target(com)
    H, c = zero.compute_qp_objective(humanoid)
    qd = np.random.default_rng(2).random(humanoid.model.nv)
    assert abs(objective_value(H, c, qd)) < 1e-9


def test_com_cost_validation():
    """tests/test_com_task.py:113-125."""
    for bad in (-1.0, [-1.0, -1.0, -1.0], -np.ones(3)):
        with pytest.raises(AssertionError):
            ComTask(cost=bad)
    task = ComTask(cost=1.0)
    task.set_cost(cost=[1.0, 1.0, 1.0])
    task.set_cost(cost=np.ones(3))


# ---- RelativeFrameTask (tests/test_relative_frame_task.py) -----------------------------------------

def test_relative_task_with_universe_root_matches_frame_task(humanoid):
    """tests/test_relative_frame_task.py:60-110: with the world as root the relative task
    is the frame task with error and Jacobian negated."""
    frame = "right_wrist_yaw_link"
    relative_task = RelativeFrameTask(frame, "universe", position_cost=1.0, orientation_cost=0.1)
    frame_task = FrameTask(frame, position_cost=1.0, orientation_cost=0.1)
    for field in ("frame=", "root=", "gain=", "orientation_cost=", "position_cost=", "lm_damping="):
        assert field in repr(relative_task)
    relative_task.set_target_from_configuration(humanoid)
    frame_task.set_target_from_configuration(humanoid)
    rng = np.random.default_rng(3)
    q = humanoid.q.copy()
    q[7:] += 0.2 * rng.standard_normal(q.size - 7)
    q = np.clip(q, humanoid.model.lowerPositionLimit, humanoid.model.upperPositionLimit)
    q[0:3] += [0.1, -0.2, 0.05]
    moved = Configuration(humanoid.model, humanoid.data, q)
    assert np.linalg.norm(frame_task.compute_error(moved)) > 1e-2
    assert np.allclose(-relative_task.compute_error(moved), frame_task.compute_error(moved), atol=ATOL)
    assert np.allclose(-relative_task.compute_jacobian(moved), frame_task.compute_jacobian(moved), atol=5e-5)


def test_relative_task_target_from_configuration(humanoid):
    task = RelativeFrameTask("left_wrist_yaw_link", "pelvis", position_cost=1.0, orientation_cost=0.1)
    with pytest.raises(TargetNotSet):
        task.compute_error(humanoid)
    task.set_target_from_configuration(humanoid)
    T = humanoid.get_transform("left_wrist_yaw_link", "pelvis")
    assert np.allclose(np.asarray(T), np.asarray(task.transform_target_to_root), atol=1e-6)
    assert np.linalg.norm(task.compute_error(humanoid)) < 1e-5


# ---- JointVelocityTask / DampingTask / LowAccelerationTask -------------------------------------------

def test_joint_velocity_task_semantics(humanoid):
    """tests/test_joint_velocity_task.py:39-78."""
    nv = humanoid.model.nv
    task = JointVelocityTask(cost=1.0)
    assert "cost=" in repr(task)
    with pytest.raises(TargetNotSet):
        task.compute_error(humanoid)
    task.set_target(np.zeros(nv), 3e-3)
    with pytest.raises(TaskDefinitionError):
        task.compute_error(humanoid)
    task.set_target(np.zeros(nv - 6), 3e-3)
    assert task.compute_error(humanoid).shape[0] == task.compute_jacobian(humanoid).shape[0]
    for dt in (3e-3, 7e-3):
        task.set_target(np.ones(nv - 6), dt)
        assert abs(task.compute_error(humanoid)[0] - dt) < 1e-9


def test_damping_and_low_acceleration_objectives(arm):
    """tests/test_damping_task.py:27-39, tests/test_low_acceleration_task.py:27-43."""
    nv = arm.model.nv
    task = DampingTask(cost=1.0)
    assert "cost=" in repr(task) and "gain=" not in repr(task) and "lm_damping=" not in repr(task)
    H, c = task.compute_qp_objective(arm)
    assert np.linalg.norm(H - np.eye(nv)) < 1e-6 and np.linalg.norm(c) < 1e-9
    task = LowAccelerationTask(cost=1.0)
    assert "cost=" in repr(task) and "gain=" not in repr(task) and "lm_damping=" not in repr(task)
    v_prev = np.array([1.0, 2.0, 3.0, 4.0, -3.0, -2.0])
    dt = 1.234e-2
    task.set_last_integration(v_prev, dt)
    H, c = task.compute_qp_objective(arm)
    assert np.linalg.norm(H - np.eye(nv)) < 1e-6 and np.linalg.norm(c + v_prev * dt) < 1e-6


# ---- LinearHolonomicTask / JointCouplingTask -----------------------------------------------------------

def knee_rows(configuration, n):
    A = np.zeros((n, configuration.model.nv))
    names = [("right_knee_joint", "right_ankle_pitch_joint"), ("left_knee_joint", "left_ankle_pitch_joint")]
    for a, b in names[:max(n, 1)]:
        A[:, get_joint_idx(configuration.model, a)[1]] = 1.0
        A[:, get_joint_idx(configuration.model, b)[1]] = -1.0
    return A


def test_linear_holonomic_task_semantics(humanoid):
    """tests/test_linear_holonomic_task.py:35-120."""
    with pytest.raises(TaskDefinitionError):
        LinearHolonomicTask(A=np.ones((3, 4)), b=np.ones(5), q_0=None)
    task = LinearHolonomicTask(A=knee_rows(humanoid, 1), b=np.zeros(1), q_0=None, cost=1.0)
    assert "cost=" in repr(task) and "gain=" in repr(task)
    wrong = LinearHolonomicTask(A=np.zeros((1, humanoid.model.nq)), b=np.zeros(1), q_0=None, cost=1.0)
    with pytest.raises(TaskJacobianNotSet):
        wrong.compute_error(humanoid)
    with pytest.raises(TaskJacobianNotSet):
        wrong.compute_jacobian(humanoid)
    rng = np.random.default_rng(4)
    q = humanoid.q.copy()
    q[7:] += 0.1 * rng.standard_normal(q.size - 7)
    moved = Configuration(humanoid.model, humanoid.data, q)
    task = LinearHolonomicTask(A=knee_rows(humanoid, 2), b=np.zeros(2), q_0=None, cost=[1.0, 1.0])
    e = task.compute_error(moved)
    J = task.compute_jacobian(moved)
    H, c = task.compute_qp_objective(moved)
    assert e.shape == (2,) and np.linalg.norm(e) > 1e-3
    assert np.allclose(J.T @ J, H, atol=ATOL) and np.allclose(e.T @ J, c, atol=ATOL)
    zero = LinearHolonomicTask(A=knee_rows(humanoid, 2), b=np.zeros(2), q_0=None, cost=[0.0, 0.0])
    H, c = zero.compute_qp_objective(moved)
    qd = rng.random(humanoid.model.nv)
    assert abs(objective_value(H, c, qd)) < 1e-9


def test_joint_coupling_task_semantics(humanoid):
    """tests/test_joint_coupling_task.py:33-70."""
    task = JointCouplingTask(["right_knee_joint", "right_ankle_pitch_joint"], [1.0, -1.0], 100.0, humanoid)
    assert "cost=" in repr(task) and "gain=" in repr(task)
    unit = JointCouplingTask(["right_knee_joint", "right_ankle_pitch_joint"], [1.0, -1.0], 1.0, humanoid)
    q = humanoid.q.copy()
    q[get_joint_idx(humanoid.model, "right_knee_joint")[0]] += 0.2
    moved = Configuration(humanoid.model, humanoid.data, q)
    e = unit.compute_error(moved)
    J = unit.compute_jacobian(moved)
    H, c = unit.compute_qp_objective(moved)
    assert e.shape == (1,) and abs(e[0] - 0.2) < 1e-6
    assert np.allclose(J.T @ J, H, atol=ATOL) and np.allclose(e.T @ J, c, atol=ATOL)
    zero = JointCouplingTask(["right_knee_joint", "right_ankle_pitch_joint"], [1.0, -1.0], 0.0, humanoid)
    H, c = zero.compute_qp_objective(moved)
    qd = np.random.default_rng(5).random(humanoid.model.nv)
    assert abs(objective_value(H, c, qd)) < 1e-9


# ---- Configuration / utils (tests/test_configuration.py:348-476, tests/test_utils.py:27-60) ---------------------

def test_configuration_semantics(humanoid):
    from pink_b200.exceptions import ConfigurationError, FrameNotFound, NotWithinConfigurationLimits
    from pink_b200.utils import VectorSpace, custom_configuration_vector

    model = humanoid.model
    robot = load_robot_description("g1_description", root_joint=JointModelFreeFlyer())
    # copy_data=False works directly on the caller's data (test_configuration.py:348-380)
    shared = Configuration(robot.model, robot.data, robot.q0, copy_data=False)
    assert shared.data is robot.data
    copied = Configuration(robot.model, robot.data, robot.q0)
    assert copied.data is not robot.data
    T = humanoid.get_transform_frame_to_world("pelvis")
    assert np.allclose(T.np[3, :], [0.0, 0.0, 0.0, 1.0])
    with pytest.raises(FrameNotFound):
        humanoid.get_transform_frame_to_world("foo")
    with pytest.raises(FrameNotFound):
        humanoid.get_frame_jacobian("does_not_exist")
    humanoid.check_limits()
    q = robot.q0.copy()
    q[-10] += 1e4
    with pytest.raises(NotWithinConfigurationLimits):
        Configuration(robot.model, robot.data, q).check_limits()
    # q is a read-only copy (test_configuration.py:420-432)
    original_q = robot.q0.copy()
    configuration = Configuration(robot.model, robot.data, original_q)
    original_q[2] = 42.0
    assert configuration.q[2] != 42.0
    with pytest.raises(ValueError):
        configuration.q[2] += 3.0
    # tangent space helpers
    v = np.arange(model.nv, dtype=float)
    assert np.allclose(configuration.tangent.eye.dot(v), v)
    assert np.sum(configuration.tangent.ones) == model.nv
    assert np.sum(configuration.tangent.zeros) == 0.0 and len(configuration.tangent.zeros) == model.nv
    tangent = VectorSpace(model.nv)
    assert tangent.eye.shape == (model.nv, model.nv) and tangent.ones.shape == (model.nv,)
    # in-place integration (test_configuration.py:469-476)
    q0 = configuration.q.copy()
    configuration.integrate_inplace(configuration.tangent.ones, dt=1e-3)
    assert np.linalg.norm(configuration.q - q0) > 2e-3
    # custom configuration vectors (test_utils.py:27-49)
    q = custom_configuration_vector(robot, left_knee_joint=0.2, right_knee_joint=-0.2)
    assert abs(q[get_joint_idx(model, "left_knee_joint")[0]] - 0.2) < 1e-12
    assert abs(q[get_joint_idx(model, "right_knee_joint")[0]] + 0.2) < 1e-12
    with pytest.raises(ConfigurationError):
        custom_configuration_vector(robot, left_knee_joint=[0.1, 0.2])
