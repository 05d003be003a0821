"""The model-independent invariants the reference's own tests assert, run on the
oracle (this is what pins it; see oracle/__init__.py "PARITY UNPINNED").

Each test names the reference test it transplants."""

import numpy as np
import pytest

from oracle import ik as oik
from oracle import kinematics as okin
from oracle import lie
from oracle import limits as olim
from oracle import tasks as otk
from tests import helpers

MODELS = ["ur5_description", "draco3_description", "g1_description"]


def _fd_jacobian(table, q, task, h=1e-6):
    def err(qq):
        return otk.task_error_jacobian(table, qq, okin.forward_kinematics(table, qq), task)[0]

    cols = []
    for i in range(table.nv):
        d = np.zeros(table.nv)
        d[i] = h
        cols.append((err(okin.integrate(table, q, d)) - err(okin.integrate(table, q, -d))) / (2 * h))
    return np.stack(cols, axis=1)


def _random_q(table, rng, spread=0.1 * np.pi):
    return okin.integrate(table, okin.neutral(table), spread * (2.0 * rng.random(table.nv) - 1.0))


@pytest.mark.parametrize("name", MODELS)
def test_task_jacobians_are_finite_differences_of_errors(name):
    """tests/test_jacobians.py:47-99 (h = 1e-6, tol 1e-5 frame / 1e-6 posture)."""
    robot, model, table = helpers.load(name)
    rng = np.random.default_rng(42)
    frame = {"ur5_description": "tool0", "draco3_description": "r_hand_contact", "g1_description": "left_wrist_yaw_link"}[name]
    other = {"ur5_description": "forearm_link", "draco3_description": "l_foot_contact", "g1_description": "right_ankle_roll_link"}[name]
    f, r = table.frame_names.index(frame), table.frame_names.index(other)
    for _ in range(4):
        q = _random_q(table, rng)
        Rt, pt = lie.exp6(rng.normal(size=6) * 0.7)
        tasks = [
            ({"type": "frame", "frame": f, "cost": np.ones(6), "target": (Rt, pt)}, 1e-5),
            ({"type": "relative_frame", "frame": f, "root": r, "cost": np.ones(6), "target": (Rt, pt)}, 1e-5),
            ({"type": "posture", "cost": 1.0, "target": okin.neutral(table)}, 1e-6),
            ({"type": "com", "cost": np.ones(3), "target": np.zeros(3)}, 1e-6),
        ]
        fk = okin.forward_kinematics(table, q)
        for task, tol in tasks:
            _, J = otk.task_error_jacobian(table, q, fk, task)
            assert np.abs(J - _fd_jacobian(table, q, task)).max() < tol, task["type"]


@pytest.mark.parametrize("name", MODELS)
def test_frame_task_at_target(name):
    """tests/test_frame_task.py:112-121: e == 0 and J == -bJ_b at the target."""
    robot, model, table = helpers.load(name)
    rng = np.random.default_rng(1)
    q = _random_q(table, rng)
    fk = okin.forward_kinematics(table, q)
    f = table.nframes - 1
    task = {"type": "frame", "frame": f, "cost": np.ones(6), "target": okin.frame_placement(table, fk, f)}
    e, J = otk.task_error_jacobian(table, q, fk, task)
    assert np.linalg.norm(e) < 1e-10
    assert np.allclose(J, -okin.frame_jacobian_local(table, fk, f))


def test_unit_cost_objective_and_zero_cost_rows():
    """tests/test_frame_task.py:123-181: H == J^T J, c == e^T J; zero cost == deleted rows."""
    robot, model, table = helpers.load("draco3_description")
    rng = np.random.default_rng(2)
    q = _random_q(table, rng)
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("r_hand_contact")
    Rf, pf = okin.frame_placement(table, fk, f)
    Rt, pt = lie.se3_mul(Rf, pf, np.eye(3), np.array([0.1, 0.02, 0.01]))
    task = {"type": "frame", "frame": f, "cost": np.ones(6), "target": (Rt, pt), "lm_damping": 0.0}
    e, J = otk.task_error_jacobian(table, q, fk, task)
    H, c = otk.task_qp_objective(table, q, fk, task)
    assert np.allclose(J.T @ J, H) and np.allclose(e @ J, c)
    for cost, rows in [((1.0, 1.0, 1.0, 0.0, 0.0, 0.0), slice(0, 3)), ((0.0, 0.0, 0.0, 1.0, 1.0, 1.0), slice(3, 6)),
                       ((0.0, 1.0, 0.0, 0.0, 0.0, 0.0), slice(1, 2))]:
        t2 = dict(task, cost=np.array(cost))
        H2, c2 = otk.task_qp_objective(table, q, fk, t2)
        assert np.allclose(H2, J[rows].T @ J[rows]) and np.allclose(c2, e[rows] @ J[rows])


def test_lm_damping_inert_at_target_active_otherwise():
    """tests/test_frame_task.py:183-215."""
    robot, model, table = helpers.load("ur5_description")
    q = _random_q(table, np.random.default_rng(3))
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("tool0")
    Rf, pf = okin.frame_placement(table, fk, f)
    at = {"type": "frame", "frame": f, "cost": np.ones(6), "target": (Rf, pf)}
    H1, c1 = otk.task_qp_objective(table, q, fk, dict(at, lm_damping=1e-8))
    H2, c2 = otk.task_qp_objective(table, q, fk, dict(at, lm_damping=1e-4))
    assert np.allclose(H1, H2) and np.allclose(c1, c2)
    off = dict(at, target=(Rf, pf + np.array([0.0, 2.0, 0.0])))
    H1, _ = otk.task_qp_objective(table, q, fk, dict(off, lm_damping=1e-8))
    H2, _ = otk.task_qp_objective(table, q, fk, dict(off, lm_damping=1e-4))
    assert np.abs(H1 - H2).max() > 1e-6


def test_relative_task_with_universe_root_is_minus_frame_task():
    """tests/test_relative_frame_task.py:60-110."""
    robot, model, table = helpers.load("g1_description")
    q = _random_q(table, np.random.default_rng(4))
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("left_wrist_yaw_link")
    u = table.frame_names.index("universe")
    Rt, pt = lie.exp6(np.array([0.3, -0.2, 0.8, 0.2, 0.1, -0.3]))
    ef, Jf = otk.task_error_jacobian(table, q, fk, {"type": "frame", "frame": f, "cost": np.ones(6), "target": (Rt, pt)})
    er, Jr = otk.task_error_jacobian(table, q, fk, {"type": "relative_frame", "frame": f, "root": u, "cost": np.ones(6), "target": (Rt, pt)})
    assert np.allclose(er, -ef, atol=1e-12)
    # J_rel = Jlog6(T_tf) J_f and J_frame = -Jlog6(T_tf) J_f
    assert np.allclose(Jr, -Jf, atol=1e-10)


def test_no_task_and_fulfilled_tasks_give_zero_velocity():
    """tests/test_solve_ik.py:79-102, 249-277."""
    robot, model, table = helpers.load("ur5_description")
    q = _random_q(table, np.random.default_rng(5))
    v, st = oik.solve_ik(table, q, [], 1e-3)
    assert st == 0 and np.abs(v).max() < 1e-12
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("tool0")
    tasks = [{"type": "frame", "frame": f, "cost": np.ones(6), "target": okin.frame_placement(table, fk, f)},
             {"type": "posture", "cost": 1e-3, "target": q}]
    v, st = oik.solve_ik(table, q, tasks, 1e-3)
    assert st == 0 and np.abs(v).max() < 1e-8


def test_single_task_converges():
    """tests/test_solve_ik.py:160-210: monotone, a few steps when limits are inactive."""
    robot, model, table = helpers.load("ur5_description")
    q = np.array([0.3, -1.0, 1.2, -0.5, 0.8, 0.1])
    fk = okin.forward_kinematics(table, q)
    f = table.frame_names.index("tool0")
    Rf, pf = okin.frame_placement(table, fk, f)
    task = {"type": "frame", "frame": f, "cost": np.ones(6), "target": (Rf, pf + np.array([0.0, 0.01, 0.0]))}
    dt, errs = 1.0, []
    for _ in range(4):
        v, st = oik.solve_ik(table, q, [task], dt, limits=[])
        q = okin.integrate(table, q, v * dt)
        errs.append(np.linalg.norm(otk.frame_task_error(table, okin.forward_kinematics(table, q), task)))
    assert errs[2] < 1e-8 and all(b <= a for a, b in zip(errs, errs[1:]))


def test_limit_rows():
    """tests/test_configuration_limit.py:32-99, tests/test_velocity_limit.py:30-55."""
    robot, model, table = helpers.load("g1_description")
    q = okin.neutral(table)
    G, h = olim.configuration_limit_rows(table, q)
    n = olim.configuration_limit_indices(table).size
    assert n == table.njoints and G.shape == (2 * n, table.nv) and h.shape == (2 * n,)
    assert (G[:, :6] == 0).all()  # floating base is never configuration-limited
    Gv, hv = olim.velocity_limit_rows(table, 0.01)
    assert Gv.shape == (2 * n, table.nv) and np.allclose(hv[:n], 0.01 * table.v_max[6:])
    # a model without limits has no rows (tests/test_limits.py:35-66)
    import types

    free = types.SimpleNamespace(**vars(table))
    free.q_max = np.full(table.nq, np.inf)
    free.q_min = np.full(table.nq, -np.inf)
    free.v_max = np.full(table.nv, np.inf)
    assert olim.configuration_limit_rows(free, q) is None and olim.velocity_limit_rows(free, 0.01) is None
    H, c, G2, h2 = oik.build_ik(free, q, [], 0.01)
    assert G2 is None and h2 is None


def test_check_limits():
    """tests/test_solve_ik.py:39-65 / pink/configuration.py:181-201."""
    robot, model, table = helpers.load("ur5_description")
    q = np.zeros(6)
    assert not olim.check_limits(table, q)
    q[2] = table.q_max[2] + 1e-3
    assert olim.check_limits(table, q)
    v, st = oik.solve_ik(table, q, [], 1e-3)
    assert st == 2
    v, st = oik.solve_ik(table, q, [], 1e-3, safety_break=False)
    assert st == 0
