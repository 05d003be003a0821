"""Committed fixtures (tests/golden/, made by scripts/make_golden.py from the fp64
oracle; see tests/golden/README.md): drift guard for the oracle, and parity of the
kernels - host build on the CPU, CUDA library under `-m gpu` - against frozen vectors."""

import os

import numpy as np
import pytest

from pink_b200 import _cabi
from tests import extras, helpers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "ur5_reachable": lambda: helpers.ur5_scenario(96, "reachable"),
    "ur5_unreachable": lambda: helpers.ur5_scenario(96, "unreachable"),
    "draco3": lambda: helpers.humanoid_scenario("draco3_description", 48),
    "g1_com_relative": lambda: helpers.humanoid_scenario("g1_description", 32, with_com=True, with_relative=True),
    "ur5_barriers_constraints": lambda: extras.ur5_extras(64),
    "g1_self_collision": lambda: extras.g1_extras(32),
}
# humanoid task sets mix costs of 0.1 and 200 (cond(H) ~ 1e6): looser fp32 bound, as in
# test_hostsim_parity.py / test_gpu_parity.py
TOL = {"draco3": dict(atol=5e-4, rtol=5e-3), "g1_com_relative": dict(atol=5e-4, rtol=5e-3)}


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _check(name, v, st, gold):
    ok = gold["status"] == 0
    assert ((st[~ok] & _cabi.PK_STATUS_NO_SOLUTION) != 0).all()
    assert (st[ok] == 0).all()
    if name in TOL:
        # humanoid fixtures (32 / 48 instances, cond(H) up to 1e7): every instance inside the
        # loosest bound of helpers.PARITY_BINS, >= 97 % inside the middle one
        assert helpers.within_tolerance(v[ok], gold["v"][ok], atol=1e-3, rtol=1e-2).all(), np.abs(v[ok] - gold["v"][ok]).max()
        assert helpers.within_tolerance(v[ok], gold["v"][ok], **TOL[name]).mean() >= 0.97
    else:
        assert helpers.within_tolerance(v[ok], gold["v"][ok]).all(), np.abs(v[ok] - gold["v"][ok]).max()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_fixtures_and_inputs_are_the_seeded_ones(name):
    sc = CASES[name]()
    gold = _load(name)
    prob, targets, _ = sc.problem()
    np.testing.assert_array_equal(sc.q32, gold["q"])
    if targets is not None:
        np.testing.assert_array_equal(targets.astype(np.float32), gold["targets"])
    v, st = sc.oracle_solve(16)
    np.testing.assert_array_equal(st, gold["status"][:16])
    np.testing.assert_allclose(v, gold["v"][:16], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_build_of_the_kernels_matches_the_fixtures(name):
    from tests.hostsim import HostSim

    sc = CASES[name]()
    gold = _load(name)
    prob, targets, _ = sc.problem()
    v, st = HostSim(sc.model).solve_ik(prob, gold["q"], gold["targets"] if gold["targets"].shape[1] else None)
    _check(name, v, st, gold)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_library_matches_the_fixtures(name):
    """Through the C-ABI (pk_solve_ik_batched via the engine) on the frozen inputs."""
    import torch

    from pink_b200.engine import get_engine

    sc = CASES[name]()
    gold = _load(name)
    prob, _, _ = sc.problem()
    eng = get_engine(sc.model)
    q = torch.as_tensor(gold["q"], device=eng.device)
    t = torch.as_tensor(gold["targets"], device=eng.device) if gold["targets"].shape[1] else None
    v, st = eng.solve_ik(prob, q, t)
    torch.cuda.synchronize()
    _check(name, v.cpu().numpy(), st.cpu().numpy(), gold)
