"""The scripts under examples/ on the GPU (the CPU suite runs them on the host build)."""

import pytest

from tests.test_examples_host import load

pytestmark = pytest.mark.gpu


def test_arm_ur5_batched_on_gpu():
    err, q = load("arm_ur5_batched").run(batch=4096, steps=240, device="cuda")
    assert q.is_cuda and tuple(q.shape) == (4096, 6)
    assert float(err.max()) < 1e-3


def test_humanoid_g1_barrier_batched_on_gpu():
    margin, closest, error = load("humanoid_g1_barrier_batched").run(batch=256, steps=150, device="cuda")
    assert -1e-3 < margin < 0.01
    assert 0.12 - 2e-3 < closest < 0.14
    assert error > 0.05
