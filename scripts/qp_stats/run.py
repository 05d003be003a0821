#!/usr/bin/env python3
"""QP path statistics on the seeded workloads, from the host build of the kernel bodies
(scripts/qp_stats/count_hooks.cpp).  Tags: chain solvers - see pk_lsq.cuh / pk_coop.cuh
(-(iterations + 2) at the exit of the two-slot loop with the number of free slots, 20 + slots
when the Cholesky rounds take over; rounds >= 1 with the free count); tree kernel - 1000 = one
factorisation with that many free columns, 2000 + k = columns changed before factorisation k + 1.

    python scripts/qp_stats/run.py            # builds /tmp/pk_qp_stats.so and prints the histograms
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SO = "/tmp/pk_qp_stats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
                       "-o", SO, os.path.join(ROOT, "scripts", "qp_stats", "count_hooks.cpp")])
import tests.hostsim as hsmod  # noqa: E402

hsmod._lib = C.CDLL(SO)
hsmod._lib.hs_last_error.restype = C.c_char_p
from tests import helpers  # noqa: E402
from tests.hostsim import HostSim  # noqa: E402


def dump(title):
    print("===", title)
    sys.stdout.flush()
    hsmod._lib.pk_count_dump()


sc = helpers.ur5_scenario(30000, "reachable")
hs = HostSim(sc.model)
prob, targets, _ = sc.problem()
for path, name in ((0, "chain kernel (round-1 solver)"), (11, "sub-warp kernel, L = 1"), (12, "sub-warp kernel, L = 2")):
    hs.solve_ik(prob, sc.q32, targets, path=path)
    dump(f"UR5 benchmark workload, 30000 instances: {name}")
for name, kw, B in (("draco3_description", {}, 400), ("g1_description", {"with_com": True}, 300)):
    sc = helpers.humanoid_scenario(name, B, **kw)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    hs.solve_ik(prob, sc.q32, targets, path=2)
    dump(f"{name}, {B} instances: tree kernel")
