"""Hostile inputs to the kernel bodies (host build): NaN / Inf / denormal-scale targets and
configurations, targets far out of reach, zero and huge costs.  Every kernel must come back
(the iteration caps hold - a hang on the GPU would stall a whole launch), flag what it cannot
solve in ``status`` and never emit a non-finite or out-of-box velocity with status 0."""

import numpy as np
import pytest

from tests import helpers
from tests.hostsim import HostSim


def run(sc, q, targets, **kw):
    hs = HostSim(sc.model)
    prob, _, _ = sc.problem()
    v, st = hs.solve_ik(prob, np.ascontiguousarray(q, dtype=np.float32),
                        np.ascontiguousarray(targets, dtype=np.float32), **kw)
    return v, st, prob


def check_clean(v, st, prob, nv, dt, rows=None):
    ok = st == 0
    if rows is not None:
        assert (st[rows] != 0).all(), st[rows]
    assert np.isfinite(v[ok]).all()
    vmax = np.array([prob.vel[i] for i in range(nv)])
    bounded = np.isfinite(vmax)
    if ok.any():
        assert (np.abs(v[ok][:, bounded]) <= vmax[bounded] * (1 + 1e-5) + 1e-6).all()
    # flagged instances carry zeros, not garbage
    bad = (st & 1) != 0
    assert np.isfinite(v[bad]).all() and (np.abs(v[bad]).max() if bad.any() else 0.0) == 0.0


def poison(targets, q, kind, rng):
    targets, q = targets.copy(), q.copy()
    B = q.shape[0]
    rows = rng.choice(B, size=max(B // 8, 1), replace=False)
    if kind == "nan_target":
        targets[rows, rng.integers(0, targets.shape[1], size=rows.size)] = np.nan
    elif kind == "inf_target":
        targets[rows, 9] = np.inf  # a translation entry of the first frame target
    elif kind == "nan_q":
        q[rows, rng.integers(0, q.shape[1], size=rows.size)] = np.nan
    elif kind == "far_target":
        targets[rows, 9:12] = 1e6
        rows = None  # solvable: the velocity box caps the step
    elif kind == "zero_rotation_target":
        targets[rows, 0:9] = 0.0  # not a rotation matrix
        rows = None  # garbage in, but must come back finite or flagged
    elif kind == "tiny_target":
        targets[rows, 9:12] = 1e-30
        rows = None
    return targets, q, rows


KINDS = ["nan_target", "inf_target", "nan_q", "far_target", "zero_rotation_target", "tiny_target"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("general_path", [False, True])
def test_ur5_kernels_on_hostile_inputs(kind, general_path):
    sc = helpers.ur5_scenario(64, "reachable", seed=11)
    _, targets, _ = sc.problem()
    rng = np.random.default_rng(5)
    t, q, rows = poison(targets, sc.q32, kind, rng)
    v, st, prob = run(sc, q, t, general_path=general_path)
    check_clean(v, st, prob, 6, sc.dt, rows)
    # the untouched instances are unaffected by their poisoned neighbours
    v0, st0, _ = run(sc, sc.q32, targets, general_path=general_path)
    clean = np.ones(64, dtype=bool)
    if rows is not None:
        clean[rows] = False
        np.testing.assert_array_equal(v[clean], v0[clean])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
def test_ur5_sub_warp_kernel_on_hostile_inputs(kind, lanes):
    """The same on the sub-warp chain kernel body (pk_coop.cuh), every lanes-per-instance variant."""
    sc = helpers.ur5_scenario(64, "reachable", seed=11)
    _, targets, _ = sc.problem()
    rng = np.random.default_rng(5)
    t, q, rows = poison(targets, sc.q32, kind, rng)
    v, st, prob = run(sc, q, t, path=10 + lanes)
    check_clean(v, st, prob, 6, sc.dt, rows)
    v0, st0, _ = run(sc, sc.q32, targets, path=10 + lanes)
    clean = np.ones(64, dtype=bool)
    if rows is not None:
        clean[rows] = False
        np.testing.assert_array_equal(v[clean], v0[clean])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("general_path", [False, True])
def test_humanoid_kernels_on_hostile_inputs(kind, general_path):
    sc = helpers.humanoid_scenario("g1_description", 16, seed=3, with_com=True)
    _, targets, _ = sc.problem()
    rng = np.random.default_rng(6)
    t, q, rows = poison(targets, sc.q32, kind, rng)
    v, st, prob = run(sc, q, t, general_path=general_path)
    check_clean(v, st, prob, sc.model.nv, sc.dt, rows)


@pytest.mark.parametrize("cost", [0.0, 1e-20, 1e8])
def test_extreme_costs(cost):
    """All-zero, vanishing and huge task weights: Tikhonov damping keeps the QP well posed
    (pink/solve_ik.py:58), the kernels stay finite."""
    sc = helpers.ur5_scenario(32, "reachable", seed=12)
    for t in sc.tasks:
        if hasattr(t, "frame"):
            t.set_position_cost(cost)
            t.set_orientation_cost(cost)
        else:
            t.cost = cost
    for kw in ({"general_path": False}, {"general_path": True}, {"path": 11}, {"path": 12}, {"path": 18}):
        _, targets, _ = sc.problem()
        v, st, prob = run(sc, sc.q32, targets, **kw)
        check_clean(v, st, prob, 6, sc.dt)
        if cost == 0.0:
            assert (st == 0).all() and np.abs(v).max() < 1e-3
