"""The scripts under examples/ run end to end, here on CPU tensors with the engine routed to the
host build of the kernels (tests/host_engine.py; test harness only)."""

import importlib.util
import os

import pytest

from tests.host_engine import host_engine  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def test_arm_ur5_batched(host_engine):  # noqa: F811
    err, q = load("arm_ur5_batched").run(batch=48, steps=240, device="cpu")
    assert tuple(q.shape) == (48, 6)
    # after 1.2 s every arm rides its moving target (Levenberg-Marquardt damping slows the
    # approach while the error is large; gain 1 makes the tracking deadbeat afterwards)
    assert float(err.max()) < 1e-3


def test_humanoid_g1_barrier_batched(host_engine):  # noqa: F811
    margin, closest, error = load("humanoid_g1_barrier_batched").run(batch=6, steps=150, device="cpu")
    assert -1e-3 < margin < 0.01          # the left hand comes up to the ceiling and stays below it
    assert 0.12 - 2e-3 < closest < 0.14   # the hands approach d_min = 0.12 and stay apart
    assert error > 0.05                   # ... which is why the hand tasks cannot be met
