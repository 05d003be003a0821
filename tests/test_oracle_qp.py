"""Pin the oracle's Goldfarb-Idnani QP against independent exact methods."""

import numpy as np
from scipy.optimize import minimize

from oracle import qp


def _random_qp(rng, n, m, feasible=True):
    M = rng.normal(size=(n, n))
    P = M @ M.T + 1e-2 * np.eye(n)
    q = rng.normal(size=n) * 3
    G = rng.normal(size=(m, n))
    x0 = rng.normal(size=n) * 0.1
    h = G @ x0 + rng.uniform(0, 0.5, size=m)
    return P, q, G, h


def test_matches_bruteforce_active_set_enumeration():
    rng = np.random.default_rng(1)
    for _ in range(200):
        n, m = rng.integers(1, 7), rng.integers(0, 9)
        P, q, G, h = _random_qp(rng, n, m)
        res = qp.solve_qp(P, q, G, h)
        assert res.found
        xb = qp.solve_qp_bruteforce(P, q, G, h)
        assert np.abs(res.x - xb).max() < 1e-9
        assert max(qp.kkt_residuals(P, q, G, h, res.x, res.z)) < 1e-8


def test_box_rows_like_the_default_limits():
    rng = np.random.default_rng(2)
    n = 6
    G = np.vstack([np.eye(n), -np.eye(n), np.eye(n), -np.eye(n)])  # [cfg; -cfg; vel; -vel]
    for _ in range(100):
        M = rng.normal(size=(n, n))
        P = M @ M.T + 1e-3 * np.eye(n)
        q = rng.normal(size=n)
        h = np.concatenate([rng.uniform(0, 0.3, n), rng.uniform(0, 0.3, n), np.full(n, 0.0157), np.full(n, 0.0157)])
        res = qp.solve_qp(P, q, G, h)
        lo = np.maximum(-h[n:2 * n], -h[3 * n:])
        hi = np.minimum(h[:n], h[2 * n:3 * n])
        r = minimize(lambda x: 0.5 * x @ P @ x + q @ x, np.clip(np.zeros(n), lo, hi), jac=lambda x: P @ x + q,
                     bounds=list(zip(lo, hi)), method="L-BFGS-B", options={"ftol": 1e-15, "gtol": 1e-12, "maxiter": 2000})
        assert np.abs(res.x - r.x).max() < 1e-6


def test_equalities_against_slsqp():
    rng = np.random.default_rng(3)
    n = 5
    for _ in range(40):
        P, q, G, h = _random_qp(rng, n, 6)
        A = rng.normal(size=(2, n))
        b = A @ rng.uniform(-0.05, 0.05, n)
        res = qp.solve_qp(P, q, G, h, A, b)
        cons = [{"type": "eq", "fun": lambda x: A @ x - b}, {"type": "ineq", "fun": lambda x: h - G @ x}]
        r = minimize(lambda x: 0.5 * x @ P @ x + q @ x, np.zeros(n), jac=lambda x: P @ x + q, constraints=cons,
                     method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
        if res.found and r.success:
            assert np.abs(res.x - r.x).max() < 1e-5


def test_infeasible_reports_not_found():
    res = qp.solve_qp(np.eye(2), np.zeros(2), np.array([[1.0, 0], [-1.0, 0]]), np.array([-1.0, -1.0]))
    assert not res.found


def test_unconstrained():
    rng = np.random.default_rng(4)
    P, q, _, _ = _random_qp(rng, 4, 0)
    res = qp.solve_qp(P, q)
    assert np.abs(res.x + np.linalg.solve(P, q)).max() < 1e-12
