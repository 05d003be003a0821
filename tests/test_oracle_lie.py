"""Pin the oracle's SE(3) maps with checks that do not depend on Pinocchio."""

import numpy as np
import scipy.linalg as sl

from oracle import lie


def _rand_twist(rng, scale):
    return rng.normal(size=6) * scale


def test_exp6_log6_against_scipy_expm():
    rng = np.random.default_rng(0)
    for scale in [1e-7, 1e-4, 1e-2, 0.5, 2.0]:
        for _ in range(5):
            xi = _rand_twist(rng, scale)
            R, p = lie.exp6(xi)
            M = np.zeros((4, 4))
            M[:3, :3] = lie.hat(xi[3:])
            M[:3, 3] = xi[:3]
            E = sl.expm(M)
            assert np.abs(E[:3, :3] - R).max() < 1e-12
            assert np.abs(E[:3, 3] - p).max() < 1e-12
            if np.linalg.norm(xi[3:]) < np.pi:
                assert np.abs(lie.log6(R, p) - xi).max() < 1e-9


def test_log3_near_pi():
    rng = np.random.default_rng(1)
    for theta in [np.pi - 0.05, np.pi - 1e-3, np.pi - 1e-6]:
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        R = lie.exp3(axis * theta)
        w, t = lie.log3(R)
        assert abs(t - theta) < 1e-9
        assert np.abs(lie.exp3(w) - R).max() < 1e-7


def test_jlog6_is_right_jacobian_inverse():
    """log6(T exp6(d)) ~ log6(T) + Jlog6(T) d (the property FrameTask relies on,
    pink/tasks/frame_task.py:222-227)."""
    rng = np.random.default_rng(2)
    for scale in [1e-6, 0.3, 1.5, 2.9]:
        R, p = lie.exp6(_rand_twist(rng, scale))
        J = lie.jlog6(R, p)
        Jfd = np.zeros((6, 6))
        h = 1e-6
        for i in range(6):
            d = np.zeros(6)
            d[i] = h
            Rp, pp = lie.se3_mul(R, p, *lie.exp6(d))
            Rm, pm = lie.se3_mul(R, p, *lie.exp6(-d))
            Jfd[:, i] = (lie.log6(Rp, pp) - lie.log6(Rm, pm)) / (2 * h)
        assert np.abs(J - Jfd).max() < 5e-9


def test_action_matrices():
    rng = np.random.default_rng(3)
    R, p = lie.exp6(_rand_twist(rng, 1.0))
    A = lie.action(R, p)
    Ai = lie.action_inverse(R, p)
    assert np.abs(A @ Ai - np.eye(6)).max() < 1e-12
    # Ad_T xi: adjoint of T exp(xi) T^-1
    xi = _rand_twist(rng, 0.1)
    Rl, pl = lie.se3_mul(*lie.se3_mul(R, p, *lie.exp6(xi)), *lie.se3_inv(R, p))
    assert np.abs(lie.log6(Rl, pl) - A @ xi).max() < 1e-10


def test_batched_shapes_and_quaternions():
    rng = np.random.default_rng(4)
    xi = rng.normal(size=(7, 6))
    R, p = lie.exp6(xi)
    assert lie.log6(R, p).shape == (7, 6) and lie.jlog6(R, p).shape == (7, 6, 6)
    q = lie.matrix_to_quat(R)
    assert np.abs(lie.quat_to_matrix(q) - R).max() < 1e-10
    assert np.abs(lie.rpy_to_matrix(0.1, -0.2, 0.3) - lie.exp3([0, 0, 0.3]) @ lie.exp3([0, -0.2, 0]) @ lie.exp3([0.1, 0, 0])).max() < 1e-14
