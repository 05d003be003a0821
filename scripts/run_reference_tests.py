#!/usr/bin/env python3
"""Run the REFERENCE's own test files (/root/reference/tests, unmodified) against the
reference's own Pink-layer code, with the oracle standing in for the third-party primitives.

Build container only (``/root/reference`` does not travel).  ``oracle/refshim`` supplies the
module names ``pinocchio``, ``qpsolvers`` and ``robot_descriptions`` (oracle-backed stand-ins,
see oracle/refshim/README.md); robots the reference clones from the network are replaced by
offline stand-ins of the same class (oracle/refshim/robot_descriptions/loaders/pinocchio.py)
or the test is skipped.  Every known-answer test of the reference that does not depend on the
geometry of one particular robot therefore exercises the oracle's FK / Jacobians / log6 /
Jlog6 / integrate / difference / CoM / QP through the reference's own assertions
(finite-difference Jacobians of tests/test_jacobians.py, the identities of
tests/test_frame_task.py, convergence in tests/test_solve_ik.py, ...).

    python scripts/run_reference_tests.py [record.txt]     # default profiles/r02j_reference_tests_over_oracle.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
OUT_OF_SCOPE = {
    "test_manipulability_task.py": "ManipulabilityTask is out of scope (SURVEY section 2)",
    "test_rolling_task.py": "wheel tasks are out of scope (SURVEY section 2)",
    "test_omniwheel_task.py": "wheel tasks are out of scope (SURVEY section 2)",
    "test_self_collision_barrier.py": "needs hpp-fcl / coal collision geometry and the iiwa14 package (not available offline)",
}
EXPECTED_FAILURES = {
    "test_configuration.py::TestConfiguration::test_constructor":
        "holds the 6 x 50 data.J of the real JVRC-1 (stand-in robot has nv = 35)",
    "test_position_barrier.py::TestPositionBarrier::test_positive_when_in_safety_zone":
        "asserts the UR3's tool position at q = 0 is positive on every axis (stand-in is the UR5: z = -0.005)",
    "test_body_spherical_barrier.py::TestBodySphericalBarrier::test_negative_when_out_of_safety_zone":
        "asserts the real YuMi's hands are closer than 0.3 m at its q0 (stand-in dual arm: 0.58 m)",
}


def main():
    record = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02j_reference_tests_over_oracle.txt")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "refshim"), ROOT, REFERENCE])
    cmd = [sys.executable, "-m", "pytest", os.path.join(REFERENCE, "tests"), "-p", "no:cacheprovider", "-q", "-rA", "-W", "ignore",
           "--tb=line"] + [f"--ignore={os.path.join(REFERENCE, 'tests', f)}" for f in OUT_OF_SCOPE]
    res = subprocess.run(cmd, cwd="/tmp", env=env, capture_output=True, text=True)
    lines = res.stdout.splitlines()
    outcomes = []
    for ln in lines:
        m = re.match(r"(PASSED|FAILED|SKIPPED|ERROR)\s+(.*)", ln)
        if m:
            outcomes.append((m.group(1), m.group(2).replace("../root/reference/tests/", "").replace("../root/reference/", "")))
    summary = next((ln for ln in reversed(lines) if re.search(r"\d+ passed", ln)), "no summary")
    unexpected = [t for o, t in outcomes if o in ("FAILED", "ERROR") and not any(t.startswith(k) for k in EXPECTED_FAILURES)]
    with open(record, "w") as fh:
        fh.write("# The reference's own tests (/root/reference/tests, unmodified) over its own Pink-layer code, with\n"
                 "# oracle/refshim standing in for pinocchio / qpsolvers / robot_descriptions.  Made by\n"
                 "# scripts/run_reference_tests.py in the build container.  NOT a GPU capture.\n")
        fh.write(f"# pytest summary: {summary.strip('= ')}\n")
        fh.write(f"# unexpected failures: {len(unexpected)}\n#\n# modules not run:\n")
        for f, why in OUT_OF_SCOPE.items():
            fh.write(f"#   {f}: {why}\n")
        fh.write("#\n# failures that come from the stand-in robots (not from the code under test):\n")
        for t, why in EXPECTED_FAILURES.items():
            fh.write(f"#   {t}: {why}\n")
        fh.write("#\n")
        for o, t in sorted(outcomes, key=lambda x: (x[1].split(" ")[0], x[0])):
            fh.write(f"{o:8s} {t}\n")
    print(summary)
    print(f"unexpected failures: {len(unexpected)}", *unexpected, sep="\n  ")
    print("record:", os.path.relpath(record, ROOT))
    return 1 if unexpected else 0


if __name__ == "__main__":
    sys.exit(main())
