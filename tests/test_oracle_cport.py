"""The oracle's C port (dense H + Goldfarb-Idnani with Givens updates) against the
numpy oracle (explicit projectors): two independent fp64 restatements must agree."""

import numpy as np

from oracle import cport
from tests import helpers
from tests.hostsim import HostSim


def _targets(sc):
    R, p = sc.oracle_tasks[0]["target"]
    return np.concatenate([R, p[:, :, None]], axis=2)[:, None]


def test_c_port_equals_numpy_oracle():
    for kind in ["reachable", "unreachable", "at_target"]:
        sc = helpers.ur5_scenario(250, kind, out_of_limits=4)
        port = cport.CPort(sc.table, sc.oracle_tasks, sc.dt, sc.damping)
        v, st = port.solve(sc.q64, _targets(sc), threads=2)
        v_ref, st_ref = sc.oracle_solve()
        np.testing.assert_array_equal(st, st_ref)
        np.testing.assert_allclose(v, v_ref, atol=1e-8, rtol=1e-8)


def test_kernel_bodies_against_c_port_at_full_batch():
    """65536 instances (BASELINE config 2 size): fp32 kernel body vs fp64 C port."""
    sc = helpers.ur5_scenario(65536, "reachable")
    port = cport.CPort(sc.table, sc.oracle_tasks, sc.dt, sc.damping)
    v_ref, st_ref = port.solve(sc.q64, _targets(sc), threads=8)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    np.testing.assert_array_equal(st, st_ref)
    ok = helpers.within_tolerance(v, v_ref)
    assert ok.mean() >= 0.999, f"{(~ok).sum()} of {ok.size} instances outside tolerance"
