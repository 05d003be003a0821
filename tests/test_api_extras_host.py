"""CPU run of the Python drop-in layer for barriers / constraints / opt-in limits:
the bodies of test_gpu_extras.py with the engine routed to the host build of the
kernels (tests/host_engine.py; test harness only)."""

import pytest

from tests import test_gpu_extras as g
from tests.host_engine import host_engine  # noqa: F401  (fixture)


@pytest.fixture(autouse=True)
def _cpu(monkeypatch, host_engine):  # noqa: F811
    monkeypatch.setattr(g, "DEVICE", "cpu")


def test_ur5_barriers_constraints_limits_match_oracle():
    g.test_ur5_barriers_constraints_limits_match_oracle()


def test_g1_config4_self_collision_barrier_matches_oracle():
    g.test_g1_config4_self_collision_barrier_matches_oracle()


def test_barrier_api_matches_oracle():
    g.test_barrier_api_matches_oracle()


def test_opt_in_limits_api():
    g.test_opt_in_limits_api()
