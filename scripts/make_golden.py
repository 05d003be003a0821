#!/usr/bin/env python3
"""Generate the fixtures under tests/golden/ (run from the repo root).

The reference (Pink + Pinocchio + quadprog) cannot be imported in this environment
(its compiled dependencies are neither vendored nor installable offline), so these
vectors come from the fp64 oracle of this repository (oracle/), which is pinned by the
reference tests' invariants (tests/test_oracle_*.py) - NOT from the reference itself.
They freeze the oracle's outputs on seeded inputs: the CPU suite checks that the
oracle still reproduces them bit-for-bit-close (drift guard), the kernel parity tests
check the CUDA path (and its host build) against them without re-running the oracle.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import extras, helpers  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def scenarios():
    yield "ur5_reachable", helpers.ur5_scenario(96, "reachable"), None
    yield "ur5_unreachable", helpers.ur5_scenario(96, "unreachable"), None
    yield "draco3", helpers.humanoid_scenario("draco3_description", 48), None
    yield "g1_com_relative", helpers.humanoid_scenario("g1_description", 32, with_com=True, with_relative=True), None
    yield "ur5_barriers_constraints", None, extras.ur5_extras(64)
    yield "g1_self_collision", None, extras.g1_extras(32)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, sc, ex in scenarios():
        s = sc if sc is not None else ex
        prob, targets, _ = s.problem()
        v, st = s.oracle_solve()
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            q=s.q32, targets=np.zeros((s.B, 0), np.float32) if targets is None else targets.astype(np.float32),
            v=v, status=st.astype(np.int32), dt=np.float64(s.dt), damping=np.float64(s.damping),
        )
        print(name, s.B, "instances, feasible", int((st == 0).sum()))


if __name__ == "__main__":
    main()
