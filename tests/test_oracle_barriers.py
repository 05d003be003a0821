"""Oracle restatement of barriers, the two opt-in limits, LinearHolonomicTask and
equality constraints, pinned by the invariants the reference's tests assert
(finite differences, signs of the safe set, closest-pair selection, known rows).

Each test names the reference test it transplants."""

import numpy as np
import pytest

from oracle import barriers as obar
from oracle import ik as oik
from oracle import kinematics as okin
from oracle import limits as olim
from oracle import qp as oqp
from oracle import tasks as otk
from tests import helpers

MODELS = ["ur5_description", "g1_description"]


def _random_q(table, rng, spread=0.25 * np.pi):
    return okin.integrate(table, okin.neutral(table), spread * (2.0 * rng.random(table.nv) - 1.0))


def _barriers(table, name, rng):
    if name.startswith("ur5"):
        fa, fb, fc = (table.frame_names.index(n) for n in ("tool0", "forearm_link", "upper_arm_link"))
    else:
        fa, fb, fc = (table.frame_names.index(n) for n in ("left_wrist_yaw_link", "right_wrist_yaw_link", "pelvis"))
    frames = [fa, fb, fc]
    pairs = [(frames[i], frames[j], 0.03 + 0.01 * i, 0.02 + 0.01 * j) for i in range(3) for j in range(i + 1, 3)]
    return [
        {"type": "position", "frame": fa, "indices": [0, 2], "p_min": np.array([-0.4, -0.1]),
         "p_max": np.array([0.6, 0.9]), "gain": np.array([1.0, 2.0]), "safe_displacement_gain": 1.0},
        {"type": "position", "frame": fb, "p_max": np.array([0.6, 0.6, 1.2]), "gain": 3.0},
        {"type": "body_spherical", "frames": (fa, fb), "d_min": 0.1, "gain": 10.0, "safe_displacement_gain": 3.0},
        {"type": "self_collision", "pairs": pairs, "n_pairs": 3, "d_min": 0.02, "gain": 20.0,
         "safe_displacement_gain": 1.0},
    ]


@pytest.mark.parametrize("name", MODELS)
def test_barrier_jacobians_are_finite_differences(name):
    """The defining property of compute_jacobian (pink/barriers/barrier.py:91-104);
    shapes as tests/test_barrier.py:34-64."""
    robot, model, table = helpers.load(name)
    rng = np.random.default_rng(5)
    for _ in range(3):
        q = _random_q(table, rng)
        fk = okin.forward_kinematics(table, q)
        for barrier in _barriers(table, name, rng):
            h0 = obar.barrier_value(table, q, fk, barrier)
            J = obar.barrier_jacobian(table, q, fk, barrier)
            assert J.shape == (h0.shape[0], table.nv)
            G, h = obar.barrier_qp_inequalities(table, q, fk, barrier, 0.01)
            assert G.shape == J.shape and h.shape == h0.shape
            eps = 1e-6
            Jfd = np.zeros_like(J)
            for i in range(table.nv):
                d = np.zeros(table.nv)
                d[i] = eps
                qp_, qm_ = okin.integrate(table, q, d), okin.integrate(table, q, -d)
                hp = obar.barrier_value(table, qp_, okin.forward_kinematics(table, qp_), barrier)
                hm = obar.barrier_value(table, qm_, okin.forward_kinematics(table, qm_), barrier)
                Jfd[:, i] = (hp - hm) / (2 * eps)
            if barrier["type"] == "self_collision":
                # rows follow the same closest-pair selection; compare as sets of rows
                order = np.argsort(h0)
                order_fd = np.argsort(h0)
                assert np.abs(J[order] - Jfd[order_fd]).max() < 1e-5
            else:
                assert np.abs(J - Jfd).max() < 1e-5, barrier["type"]


def test_barrier_objective_and_sign_conventions():
    """tests/test_barrier.py:66-88 (no penalty weight -> H = c = 0; weight -> H != 0),
    tests/test_position_barrier.py:78-93, tests/test_body_spherical_barrier.py:77-89."""
    robot, model, table = helpers.load("ur5_description")
    q = np.array([0.3, -1.2, 1.0, 0.2, 0.4, -0.3])
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("tool0")
    _, p = okin.frame_placement(table, fk, f)
    inside = {"type": "position", "frame": f, "p_min": p - 0.1, "p_max": p + 0.1, "gain": 1.0}
    assert np.all(obar.barrier_value(table, q, fk, inside) > 0)
    outside = {"type": "position", "frame": f, "p_min": p + 0.1, "gain": 1.0}
    assert np.any(obar.barrier_value(table, q, fk, outside) < 0)
    H, c = obar.barrier_qp_objective(table, q, fk, inside)
    assert np.allclose(H, 0) and np.allclose(c, 0)
    weighted = dict(inside, safe_displacement_gain=2.0)
    H, c = obar.barrier_qp_objective(table, q, fk, weighted)
    J = obar.barrier_jacobian(table, q, fk, weighted)
    assert np.allclose(H, 2.0 / (J * J).sum() * np.eye(6)) and np.allclose(c, 0)
    g = table.frame_names.index("forearm_link")
    _, p2 = okin.frame_placement(table, fk, g)
    dist = np.linalg.norm(p - p2)
    assert obar.barrier_value(table, q, fk, {"type": "body_spherical", "frames": (f, g), "d_min": 0.5 * dist})[0] > 0
    assert obar.barrier_value(table, q, fk, {"type": "body_spherical", "frames": (f, g), "d_min": 2.0 * dist})[0] < 0
    # class-K function of the spherical barrier (body_spherical_barrier.py:65)
    b = {"type": "body_spherical", "frames": (f, g), "d_min": 0.5 * dist, "gain": 7.0}
    h = obar.barrier_value(table, q, fk, b)
    _, rhs = obar.barrier_qp_inequalities(table, q, fk, b, 0.01)
    assert np.allclose(rhs, 7.0 * h / (1 + np.abs(h)))
    # position barrier: per-index gains are tiled over [min; max] (position_barrier.py:84-85)
    b = {"type": "position", "frame": f, "indices": [0, 1], "p_min": p[:2] - 0.1, "p_max": p[:2] + 0.2,
         "gain": np.array([1.0, 2.0])}
    _, rhs = obar.barrier_qp_inequalities(table, q, fk, b, 0.01)
    assert np.allclose(rhs, [0.1, 0.2, 0.2, 0.4])


def test_self_collision_keeps_the_closest_pairs():
    """tests/test_self_collision_barrier.py:139-159: fewer than `dim` pairs are
    closer than any selected pair."""
    robot, model, table = helpers.load("g1_description")
    rng = np.random.default_rng(2)
    frames = [table.frame_names.index(n) for n in
              ("left_wrist_yaw_link", "right_wrist_yaw_link", "pelvis", "left_ankle_roll_link", "right_ankle_roll_link")]
    pairs = [(frames[i], frames[j], 0.05, 0.04) for i in range(5) for j in range(i + 1, 5)]
    for dim in (1, 3, len(pairs)):
        barrier = {"type": "self_collision", "pairs": pairs, "n_pairs": dim, "d_min": 0.02, "gain": 1.0}
        q = _random_q(table, rng)
        fk = okin.forward_kinematics(table, q)
        h = obar.barrier_value(table, q, fk, barrier)
        assert h.shape == (dim,)
        all_d = np.array([d for d, _, _ in obar.sphere_pair_distances(table, fk, pairs)]) - 0.02
        for h_i in h:
            assert np.sum(all_d < h_i) < dim
    # colliding spheres: negative barrier (tests/test_self_collision_barrier.py:118-137)
    big = [(frames[0], frames[1], 5.0, 5.0)]
    fk = okin.forward_kinematics(table, okin.neutral(table))
    assert obar.barrier_value(table, okin.neutral(table), fk, {"type": "self_collision", "pairs": big, "n_pairs": 1, "d_min": 0.0})[0] < 0


def test_floating_base_velocity_limit_rows():
    """tests/test_floating_base_velocity_limit.py:48-140: only root columns are
    non-zero, 2 rows per finite bound, inside/outside displacements."""
    robot, model, table = helpers.load("g1_description")
    rng = np.random.default_rng(3)
    q = _random_q(table, rng)
    fk = okin.forward_kinematics(table, q)
    frame = table.frame_names.index("pelvis")
    twist_max = np.array([0.4, 0.2, np.inf, np.inf, np.inf, 1.0])
    dt = 0.01
    G, h = olim.floating_base_velocity_rows(table, fk, frame, twist_max, dt)
    assert G.shape == (6, table.nv) and h.shape == (6,)
    assert np.linalg.norm(G[:, :6]) > 0 and np.allclose(G[:, 6:], 0)
    assert np.allclose(h, dt * np.array([0.4, 0.2, 1.0, 0.4, 0.2, 1.0]))
    dq_in = np.zeros(table.nv)
    dq_in[:6] = 0.5 * dt * np.array([0.4, 0.2, 5.0, 5.0, 5.0, 1.0])
    dq_out = dq_in.copy()
    dq_out[0] = 2.0 * dt * 0.4
    # pelvis is the root body itself: the rows select base twist coordinates
    assert np.all(G @ dq_in <= h + 1e-12) and np.any(G @ dq_out > h + 1e-12)
    assert olim.floating_base_velocity_rows(table, fk, frame, np.full(6, np.inf), dt) is None


def test_acceleration_limit_rows():
    """pink/limits/acceleration_limit.py:119-200 on known numbers, and
    tests/test_acceleration_limit.py:54-92 (a joint without configuration limits
    keeps lower <= upper)."""
    robot, model, table = helpers.load("ur5_description")
    q = np.array([0.3, -1.2, 1.0, 0.2, 0.4, -0.3])
    a_max = np.array([10.0, 20.0, np.inf, 5.0, 0.0, 8.0])
    dt = 5e-3
    prev = np.array([0.01, -0.02, 0.0, 0.001, 0.0, 0.0])
    G, h = olim.acceleration_limit_rows(table, q, a_max, prev, dt)
    idx = [0, 1, 3, 5]
    assert G.shape == (8, 6)
    assert np.allclose(G[:4], np.eye(6)[idx]) and np.allclose(G[4:], -np.eye(6)[idx])
    a = a_max[idx]
    up = np.minimum(a * dt * dt + prev[idx], dt * np.sqrt(2 * a * (table.q_max[idx] - q[idx])))
    lo = np.minimum(a * dt * dt - prev[idx], dt * np.sqrt(2 * a * (q[idx] - table.q_min[idx])))
    assert np.allclose(h, np.concatenate([up, lo]))
    from pink_b200.model import Model, SE3

    m = Model("continuous")
    m.add_joint("joint1", 0, SE3(np.eye(3), np.zeros(3)), np.array([0.0, 0.0, 1.0]), kind="revolute",
                lower=-np.inf, upper=np.inf, velocity=np.inf)
    t = m.table()
    G, h = olim.acceleration_limit_rows(t, np.zeros(1), np.array([14.0]), np.array([3.0 * dt]), dt)
    assert -h[1] <= h[0]


def test_linear_task_and_joint_coupling():
    """tests/test_linear_holonomic_task.py:64-118, tests/test_joint_coupling_task.py:44-70:
    unit cost H = J^T J, c = e^T J; Jacobian is a finite difference of the error."""
    robot, model, table = helpers.load("g1_description")
    rng = np.random.default_rng(9)
    q = _random_q(table, rng)
    fk = okin.forward_kinematics(table, q)
    A = np.zeros((2, table.nv))
    A[0, 10], A[0, 11] = 1.0, -1.0
    A[1, 6:] = rng.normal(size=table.nv - 6)
    task = {"type": "linear", "A": A, "b": np.array([0.1, -0.2]), "q0": None, "cost": np.ones(2), "gain": 1.0}
    e, J = otk.task_error_jacobian(table, q, fk, task)
    H, c = otk.task_qp_objective(table, q, fk, task)
    # c = +gain e^T W^T W J (pink/tasks/task.py:145-166); the reference test evaluates at e = 0
    assert np.allclose(J.T @ J, H) and np.allclose(e @ J, c)
    eps = 1e-6
    for i in range(table.nv):
        d = np.zeros(table.nv)
        d[i] = eps
        ep = otk.task_error_jacobian(table, okin.integrate(table, q, d), None, task)[0]
        em = otk.task_error_jacobian(table, okin.integrate(table, q, -d), None, task)[0]
        assert np.abs((ep - em) / (2 * eps) - J[:, i]).max() < 1e-6
    # full A (root columns included) still differentiates correctly
    A2 = rng.normal(size=(3, table.nv))
    q0 = _random_q(table, rng)
    task2 = {"type": "linear", "A": A2, "b": np.zeros(3), "q0": q0, "cost": np.ones(3)}
    e, J = otk.task_error_jacobian(table, q, fk, task2)
    for i in range(table.nv):
        d = np.zeros(table.nv)
        d[i] = eps
        ep = otk.task_error_jacobian(table, okin.integrate(table, q, d), None, task2)[0]
        em = otk.task_error_jacobian(table, okin.integrate(table, q, -d), None, task2)[0]
        assert np.abs((ep - em) / (2 * eps) - J[:, i]).max() < 1e-5


def test_equality_constraints_and_barriers_in_solve_ik():
    """pink/solve_ik.py:125-149: a task passed as constraint is met exactly
    (J dq = -gain e); barrier rows G dq <= h hold at the solution; removing an
    inactive barrier does not change the solution."""
    robot, model, table = helpers.load("ur5_description")
    rng = np.random.default_rng(4)
    scn = helpers.ur5_scenario(8)
    f = table.frame_names.index("tool0")
    for i in range(8):
        q = scn.q64[i]
        tasks = [oik._slice_task(t, i) for t in scn.oracle_tasks]
        fk = okin.forward_kinematics(table, q)
        _, p = okin.frame_placement(table, fk, f)
        barrier = {"type": "position", "frame": f, "p_max": p + np.array([1e-4, 0.5, 1e-4]), "gain": 5.0}
        coupling = {"type": "linear", "A": np.array([[1.0, 1.0, 0, 0, 0, 0]]), "b": np.zeros(1),
                    "q0": q.copy(), "cost": np.ones(1), "gain": 1.0}
        v, st = oik.solve_ik(table, q, tasks, scn.dt, scn.damping, None, True, [barrier], [coupling])
        assert st == 0
        H, c, G, h, A, b = oik.assemble(table, q, tasks, scn.dt, scn.damping, None, [barrier], [coupling])
        dq = v * scn.dt
        assert np.abs(A @ dq - b).max() < 1e-10
        assert np.all(G @ dq <= h + 1e-9)
        res = oqp.solve_qp(H, c, G, h, A, b)
        assert np.abs(res.x - dq).max() < 1e-12
        loose = dict(barrier, p_max=p + 10.0)
        v2, _ = oik.solve_ik(table, q, tasks, scn.dt, scn.damping, None, True, [loose], [coupling])
        v3, _ = oik.solve_ik(table, q, tasks, scn.dt, scn.damping, None, True, None, [coupling])
        assert np.abs(v2 - v3).max() < 1e-9
