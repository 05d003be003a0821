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
    python scripts/run_reference_tests.py --product [record.txt]
        # the same files against THIS package: `pink` -> pink_b200, engine = host build of the kernels (fp32);
        # default record profiles/r02k_reference_tests_on_product_host_build.txt
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


# --product: the module name `pink` is bound to THIS package (tests/refalias_plugin.py), engine = host build of
# the kernels (fp32).  The reference's tests were written for fp64 Pinocchio: what fails here and why.
PRODUCT_EXPECTED_FAILURES = dict(EXPECTED_FAILURES)
PRODUCT_EXPECTED_FAILURES.update({
    "test_configuration.py::TestConfiguration::test_copy_no_forward_kinematics":
        "reads Pinocchio's data.J; this package evaluates kinematics lazily on the device and never materialises data.J (DESIGN section 1)",
    "test_jacobians.py::TestJacobians::test_frame_task":
        "finite differences with step 1e-6 of a function evaluated in fp32 (q itself is quantised at 1e-7)",
    "test_jacobians.py::TestJacobians::test_joint_coupling_task": "idem",
    "test_jacobians.py::TestJacobians::test_posture_task": "idem",
    "test_com_task.py::TestComTask::test_zero_error_when_target_at_body": "asks for |e| < 1e-10; fp32 gives 4e-9",
    "test_low_acceleration_task.py::TestLowAccelerationTask::test_qp_objective": "asks for 1e-10; fp32 gives 2e-9",
    "test_solve_ik.py::TestSolveIK::test_three_tasks_convergence": "asks for |v| < 1e-6 at convergence; fp32 floor 3e-4",
    "test_solve_ik.py::TestSolveIK::test_com_task_convergence":
        "asks for |v| < 2e-5 at convergence, at the fp32 floor: passes on the default host build, 1.2e-4 on the FMA-contracting one",
    "test_frame_task.py::TestFrameTask::test_lm_damping_has_effect_under_error":
        "mu = 1e-8 |e|^2 is below the fp32 resolution of H: H stays singular and the test's own unconstrained solve has no answer",
})


def main():
    product = "--product" in sys.argv
    args = [a for a in sys.argv[1:] if a != "--product"]
    default = "r02k_reference_tests_on_product_host_build.txt" if product else "r02j_reference_tests_over_oracle.txt"
    record = args[0] if args else os.path.join(ROOT, "profiles", default)
    expected = PRODUCT_EXPECTED_FAILURES if product else EXPECTED_FAILURES
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "refshim"), ROOT, REFERENCE])
    cmd = [sys.executable, "-m", "pytest", os.path.join(REFERENCE, "tests"), "-p", "no:cacheprovider", "-q", "-rA", "-W", "ignore",
           "--tb=line"] + [f"--ignore={os.path.join(REFERENCE, 'tests', f)}" for f in OUT_OF_SCOPE]
    if product:
        cmd += ["-p", "tests.refalias_plugin", "--import-mode=importlib"]
    res = subprocess.run(cmd, cwd="/tmp", env=env, capture_output=True, text=True)
    lines = res.stdout.splitlines()
    outcomes = []
    for ln in lines:
        m = re.match(r"(PASSED|FAILED|SKIPPED|ERROR)\s+(.*)", ln)
        if m:
            outcomes.append((m.group(1), m.group(2).replace("../root/reference/tests/", "").replace("../root/reference/", "")))
    summary = next((ln for ln in reversed(lines) if re.search(r"\d+ passed", ln)), "no summary")
    unexpected = [t for o, t in outcomes if o in ("FAILED", "ERROR") and not any(t.startswith(k) for k in expected)]
    with open(record, "w") as fh:
        if product:
            fh.write("# The reference's own tests (/root/reference/tests, unmodified) against THIS package: the module name\n"
                     "# `pink` is bound to pink_b200 (tests/refalias_plugin.py), the engine is the host build of the CUDA kernel\n"
                     "# bodies (fp32), pinocchio / qpsolvers / robot_descriptions are the stand-ins of oracle/refshim.  Made by\n"
                     "# scripts/run_reference_tests.py --product in the build container.  NOT a GPU capture.\n")
        else:
            fh.write("# The reference's own tests (/root/reference/tests, unmodified) over its own Pink-layer code, with\n"
                     "# oracle/refshim standing in for pinocchio / qpsolvers / robot_descriptions.  Made by\n"
                     "# scripts/run_reference_tests.py in the build container.  NOT a GPU capture.\n")
        fh.write(f"# pytest summary: {summary.strip('= ')}\n")
        fh.write(f"# unexpected failures: {len(unexpected)}\n#\n# modules not run:\n")
        for f, why in OUT_OF_SCOPE.items():
            fh.write(f"#   {f}: {why}\n")
        fh.write("#\n# failures with a known cause outside the code under test (stand-in robots; with --product also fp32):\n")
        for t, why in expected.items():
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
