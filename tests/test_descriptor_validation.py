"""Error behaviour of the problem-descriptor validation (pk_marshal.hpp, shared by the
CUDA library and the host build): malformed descriptors are rejected with a message,
never executed.  Mirrors the reference's argument checks (TaskDefinitionError,
FrameNotFound, NegativeMinimumDistance, InvalidCollisionPairs ...) at the C-ABI level."""

import copy
import ctypes as C

import numpy as np
import pytest

from pink_b200 import _cabi
from tests import extras, helpers
from tests.hostsim import HostSim


def _run(hs, prob, sc):
    _, targets, _ = sc.problem()
    return hs.solve_ik(prob, sc.q32[:4], None if targets is None else targets[:4])


def _fresh(sc):
    prob, _, _ = sc.problem()
    return prob


@pytest.fixture(scope="module")
def ur5():
    sc = extras.ur5_extras(8)
    return sc, HostSim(sc.model)


def test_valid_descriptor_runs(ur5):
    sc, hs = ur5
    v, st = _run(hs, _fresh(sc), sc)
    assert v.shape == (4, 6)


@pytest.mark.parametrize("mutate,fragment", [
    (lambda p: setattr(p, "ntasks", 13), "ntasks"),
    (lambda p: setattr(p, "ntasks", -1), "ntasks"),
    (lambda p: setattr(p, "dt", 0.0), "dt"),
    (lambda p: setattr(p, "target_stride", -3), "target_stride"),
    (lambda p: setattr(p.tasks[0], "frame", 10_000), "frame"),
    (lambda p: setattr(p.tasks[0], "type", 77), "task type"),
    (lambda p: setattr(p.tasks[0], "target_offset", 10_000), "target"),
    (lambda p: p.tasks[0].cost.__setitem__(0, -1.0), "cost"),
    (lambda p: setattr(p, "nbarriers", 9), "nbarriers"),
    (lambda p: setattr(p, "nconstraints", 5), "nconstraints"),
    (lambda p: setattr(p.barriers[0], "type", 9), "barrier type"),
    (lambda p: setattr(p.barriers[0], "frame", -2), "frame"),
    (lambda p: setattr(p.barriers[0], "nidx", 4), "indices"),
    (lambda p: p.barriers[0].indices.__setitem__(0, 3), "index"),
    (lambda p: (setattr(p.barriers[0], "has_min", 0), setattr(p.barriers[0], "has_max", 0)), "p_min or p_max"),
    (lambda p: setattr(p.barriers[0], "dim", 5), "dim"),
    (lambda p: setattr(p.barriers[1], "d_min", -0.1), "negative minimum distance"),
    (lambda p: setattr(p.barriers[1], "gain_function", 5), "gain function"),
    (lambda p: setattr(p.barriers[2], "dim", 99), "dim"),
    (lambda p: setattr(p.barriers[2], "npairs", 300), "pairs"),
    (lambda p: setattr(p.barriers[2], "pair_offset", 4), "pairs out of range"),
    (lambda p: setattr(p.barriers[2], "data_offset", 10_000), "radii"),
    (lambda p: setattr(p.constraints[0], "rows", 7), "rows"),
    (lambda p: setattr(p.constraints[0], "data_offset", 10_000), "extra"),
    (lambda p: setattr(p.constraints[0], "type", _cabi.PK_TASK_POSTURE), "equality"),
    (lambda p: setattr(p, "acc_prev_offset", 10_000), "dq_prev"),
    (lambda p: (setattr(p, "fb_enabled", 1), setattr(p, "fb_frame", 0)), "floating-base"),
    (lambda p: setattr(p, "n_extra", -1), "extra"),
])
def test_malformed_descriptors_are_rejected(ur5, mutate, fragment):
    sc, hs = ur5
    prob = _fresh(sc)
    mutate(prob)
    with pytest.raises(RuntimeError) as err:
        _run(hs, prob, sc)
    assert fragment.lower() in str(err.value).lower(), str(err.value)


def test_too_many_dense_rows_and_linear_root_columns():
    sc = extras.g1_extras(4)
    hs = HostSim(sc.model)
    prob = _fresh(sc)
    prob.barriers[0].dim = 30  # more rows than PK_MAX_INEQ_ROWS (with 36 pairs available)
    with pytest.raises(RuntimeError, match="(?i)dense inequality rows|dim"):
        _run(hs, prob, sc)
    # a LINEAR task acting on floating-base columns is refused (documented restriction)
    prob = _fresh(sc)
    lin = [k for k in range(prob.ntasks) if prob.tasks[k].type == _cabi.PK_TASK_LINEAR][0]
    extra = np.ctypeslib.as_array(prob.extra, shape=(prob.n_extra,))
    extra[prob.tasks[lin].data_offset] = 1.0
    with pytest.raises(RuntimeError, match="(?i)root columns"):
        _run(hs, prob, sc)
