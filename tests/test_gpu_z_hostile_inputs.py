"""GPU twin of tests/test_hostsim_degenerate_inputs.py (runs last in the `-m gpu` suite):
NaN / Inf / far / malformed inputs through the C-ABI.  Every launch must come back, poisoned
instances carry a non-zero status and zero velocity, their neighbours are bit-identical to
a clean launch."""

import numpy as np
import pytest
import torch

from pink_b200.engine import get_engine
from tests import helpers
from tests.test_hostsim_degenerate_inputs import KINDS, check_clean, poison

pytestmark = pytest.mark.gpu


def gpu_run(sc, q, targets):
    eng = get_engine(sc.model, torch.device("cuda", 0))
    prob, _, _ = sc.problem()
    v, st = eng.solve_ik(prob, torch.as_tensor(np.ascontiguousarray(q, dtype=np.float32), device="cuda"),
                         torch.as_tensor(np.ascontiguousarray(targets, dtype=np.float32), device="cuda"))
    torch.cuda.synchronize()
    return v.cpu().numpy(), st.cpu().numpy(), prob


@pytest.mark.parametrize("kind", KINDS)
def test_ur5_chain_kernel_on_hostile_inputs(kind):
    sc = helpers.ur5_scenario(4096, "reachable", seed=11)
    _, targets, _ = sc.problem()
    t, q, rows = poison(targets, sc.q32, kind, np.random.default_rng(5))
    v, st, prob = gpu_run(sc, q, t)
    check_clean(v, st, prob, 6, sc.dt, rows)
    if rows is not None:
        v0, _, _ = gpu_run(sc, sc.q32, targets)
        clean = np.ones(4096, dtype=bool)
        clean[rows] = False
        np.testing.assert_array_equal(v[clean], v0[clean])


@pytest.mark.parametrize("kind", KINDS)
def test_humanoid_tree_kernel_on_hostile_inputs(kind):
    sc = helpers.humanoid_scenario("g1_description", 256, seed=3, with_com=True)
    _, targets, _ = sc.problem()
    t, q, rows = poison(targets, sc.q32, kind, np.random.default_rng(6))
    v, st, prob = gpu_run(sc, q, t)
    check_clean(v, st, prob, sc.model.nv, sc.dt, rows)
