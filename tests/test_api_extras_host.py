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


def test_barrier_fulfilled_and_violated_like_the_reference():
    """tests/test_solve_ik.py:104-158 on the unbatched API: with the only task fulfilled
    the velocity is zero when the position barrier holds with margin, and non-zero when
    the barrier is violated (the QP must move the frame back)."""
    import numpy as np

    import pink_b200
    from pink_b200.barriers import PositionBarrier
    from pink_b200.model import JointModelFreeFlyer
    from pink_b200.robots import load_robot_description

    robot = load_robot_description("g1_description", root_joint=JointModelFreeFlyer())
    configuration = pink_b200.Configuration(robot.model, robot.data, robot.q0)
    frame = "left_ankle_roll_link"
    task = pink_b200.FrameTask(frame, position_cost=1.0, orientation_cost=1.0)
    task.set_target(configuration.get_transform_frame_to_world(frame))
    p = configuration.get_transform_frame_to_world(frame).translation
    barrier = PositionBarrier(frame, p_min=p - 0.1 * np.ones(3))
    velocity = pink_b200.solve_ik(configuration, [task], dt=5e-3, solver="daqp", barriers=[barrier],
                                  limits=[configuration.model.configuration_limit])
    assert velocity.shape == (robot.model.nv,) and np.allclose(velocity, 0.0, atol=1e-5)
    barrier = PositionBarrier(frame, p_min=p + 0.01 * np.ones(3))
    velocity = pink_b200.solve_ik(configuration, [task], dt=5e-3, solver="daqp", barriers=[barrier],
                                  limits=[configuration.model.configuration_limit])
    assert not np.allclose(velocity, 0.0, atol=1e-3)
    # the frame moves back inside: dh/dt + gain * h >= 0 with h = -0.01
    J = barrier.compute_jacobian(configuration)
    h = barrier.compute_barrier(configuration)
    assert np.all(J @ velocity + 1.0 * h >= -1e-5)


def test_joint_coupling_tasks_on_the_tree_kernel():
    g.test_joint_coupling_tasks_on_the_tree_kernel()


def test_config4_feasibility_and_sample_parity_small(monkeypatch):
    """Same body as the full-size GPU test, on a batch the host build finishes quickly."""
    real = g.extras.g1_extras
    monkeypatch.setattr(g.extras, "g1_extras", lambda B, floating_base_limit=True: real(96, floating_base_limit=floating_base_limit))
    monkeypatch.setattr(g.np.random, "default_rng", lambda seed=0: _SmallChoice(real_rng(seed)))
    g.test_config4_full_batch_feasibility_and_sample_parity(False)
    g.test_config4_full_batch_feasibility_and_sample_parity(True)


import numpy as _np

real_rng = _np.random.default_rng


class _SmallChoice:
    def __init__(self, rng):
        self.rng = rng

    def choice(self, a, size, replace):
        return self.rng.choice(a, size=min(size, 24), replace=replace)

    def __getattr__(self, name):
        return getattr(self.rng, name)
