#!/usr/bin/env python
"""Robustness soak of the kernel bodies on the host build (tests/hostsim; test harness, fp32,
same source as the CUDA kernels): many seeds / weights / time steps, checked against the fp64
oracle (humanoids, extras) or the fp64 KKT conditions of each instance's own QP (UR5 at scale).
Prints one line per configuration; `profiles/r01l_hostsim_soak.txt` and
`profiles/r02i_hostsim_soak.txt` are committed runs.

    PYTHONPATH=. python scripts/soak_hostsim.py [--quick] [--lanes 2] [--wide] [--fma]
"""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import ik as oik  # noqa: E402
from tests import extras, helpers  # noqa: E402
from tests.hostsim import HostSim  # noqa: E402


def ur5_kkt(B, path=None):
    which = "chain kernel (one instance per thread)" if path is None else f"sub-warp kernel, {path - 10} lane(s) per instance"
    print(f"# UR5 {which}: status, primal violation, relative stationarity residual of the fp64 KKT system")
    total = 0
    for kind in ("reachable", "unreachable"):
        for seed in (101, 102):
            for dt in (0.005, 0.02, 0.1):
                for lm, pc in ((1.0, 1e-3), (0.0, 1e-3), (1.0, 0.0), (1e-3, 1e-1)):
                    sc = helpers.ur5_scenario(B, kind, seed=seed, lm_damping=lm, posture_cost=pc)
                    sc.dt = dt
                    hs = HostSim(sc.model)
                    prob, targets, _ = sc.problem()
                    v, st = hs.solve_ik(prob, sc.q32, targets, path=path)
                    H, c, G, h = sc.oracle_build()
                    stat, prim, _, _ = oik.kkt_check_batch(H, c, G, h, v.astype(np.float64) * sc.dt)
                    r = stat / (np.abs(c).max(axis=1) + 1e-12)
                    total += B
                    print(f"ur5 {kind:11s} seed {seed} dt {dt:<5} lm {lm:<5} posture {pc:<5} status!=0 {int((st != 0).sum())} "
                          f"iter_limit {int(((st & 8) != 0).sum())} primal {prim.max():.1e} stationarity q99.9 "
                          f"{np.quantile(r, 0.999):.1e} max {r.max():.1e} n>1e-3 {int((r > 1e-3).sum())}", flush=True)
    print(f"# {total} instances")


def ur5_wide(B):
    """More seeds, printing only the configurations with a residual above 3e-4."""
    print("# UR5 chain kernel, seeds 103..112: configurations with relative stationarity > 3e-4, then the total")
    total, worst, bad = 0, 0.0, 0
    for kind in ("reachable", "unreachable"):
        for seed in range(103, 113):
            for dt in (0.005, 0.02, 0.1):
                for lm, pc in ((1.0, 1e-3), (0.0, 1e-3), (1e-3, 1e-1)):
                    sc = helpers.ur5_scenario(B, kind, seed=seed, lm_damping=lm, posture_cost=pc)
                    sc.dt = dt
                    hs = HostSim(sc.model)
                    prob, targets, _ = sc.problem()
                    v, st = hs.solve_ik(prob, sc.q32, targets)
                    H, c, G, h = sc.oracle_build()
                    stat, prim, _, _ = oik.kkt_check_batch(H, c, G, h, v.astype(np.float64) * sc.dt)
                    r = stat / (np.abs(c).max(axis=1) + 1e-12)
                    total += B
                    bad += int((r > 1e-3).sum())
                    worst = max(worst, r.max())
                    if r.max() > 3e-4 or (st != 0).any() or prim.max() > 1e-6:
                        print(f"ur5 {kind} seed {seed} dt {dt} lm {lm} posture {pc}: status!=0 {int((st != 0).sum())} primal "
                              f"{prim.max():.1e} stationarity max {r.max():.1e} n>1e-3 {int((r > 1e-3).sum())}", flush=True)
    print(f"# {total} instances, worst relative stationarity {worst:.1e}, n>1e-3 {bad}")


def chains(B):
    """Every <NJ, NFT> instantiation of both chain kernels on random chains (prismatic joints, mid-chain
    frames), with and without Levenberg-Marquardt term: fp64 KKT certificate."""
    print("# random chains, both chain kernels (one instance per thread | 2 lanes per instance): lines only for findings")
    total = 0
    for nj, kw in ((2, {}), (3, {"prismatic": (1,)}), (4, {"two_tasks": True}), (5, {}), (6, {"two_tasks": True, "prismatic": (0, 4)}),
                   (6, {}), (7, {"two_tasks": True, "prismatic": (2,)}), (7, {})):
        for seed in range(20, 26):
            for dt in (0.005, 0.1):
                for lm in (0.1, 0.0):
                    sc = helpers.chain_scenario(nj, B, seed=seed, **kw)
                    sc.dt = dt
                    for t, o in zip(sc.tasks, sc.oracle_tasks):
                        if o["type"] == "frame":
                            t.lm_damping = lm
                            o["lm_damping"] = lm
                    hs = HostSim(sc.model)
                    prob, targets, _ = sc.problem()
                    H, c, G, h = sc.oracle_build()
                    res = []
                    for path in (None, 12):
                        v, st = hs.solve_ik(prob, sc.q32, targets, path=path)
                        stat, prim, _, _ = oik.kkt_check_batch(H, c, G, h, v.astype(np.float64) * sc.dt)
                        r = stat / (np.abs(c).max(axis=1) + 1e-12)
                        res.append((int((st != 0).sum()), prim.max(), r.max(), int((r > 1e-3).sum())))
                    total += B
                    if any(x[0] or x[1] > 1e-6 or x[3] for x in res):
                        print(f"chain{nj} {kw} seed {seed} dt {dt} lm {lm}: " + " | ".join(
                            f"status!=0 {a} primal {b:.1e} stationarity max {m:.1e} n>1e-3 {d}" for a, b, m, d in res), flush=True)
    print(f"# {total} instances per kernel")


def humanoids(B):
    print("# humanoid tree kernel (box QP) against the oracle: tolerance 5e-4 + 5e-3 |v|")
    for name, kw in (("draco3_description", {}), ("g1_description", {"with_com": True}),
                     ("draco3_description", {"with_relative": True}),
                     ("g1_description", {"with_com": True, "with_relative": True})):
        for seed in (31, 32, 33):
            for sigma in (0.15, 0.6):
                sc = helpers.humanoid_scenario(name, B, seed=seed, sigma=sigma, **kw)
                hs = HostSim(sc.model)
                prob, targets, _ = sc.problem()
                v, st = hs.solve_ik(prob, sc.q32, targets)
                v_ref, st_ref = sc.oracle_solve()
                both = (st == 0) & (st_ref == 0)
                ok = helpers.within_tolerance(v[both], v_ref[both], atol=5e-4, rtol=5e-3)
                print(f"{name[:6]} {str(kw):45s} seed {seed} sigma {sigma} status!=0 {int((st != 0).sum())} "
                      f"oracle!=0 {int((st_ref != 0).sum())} off {int((~ok).sum())}/{int(both.sum())} "
                      f"worst {np.abs(v - v_ref)[both].max():.1e} ({'tree' if hs.used_tree else 'general'})", flush=True)


def with_extras(B_ur5, B_g1):
    print("# barriers + equality constraints + opt-in limits: warp-cooperative dual QP and general path vs oracle")
    for name, fn, B in (("ur5", extras.ur5_extras, B_ur5), ("g1", extras.g1_extras, B_g1)):
        for seed in range(11, 17):
            sc = fn(B, seed=seed)
            hs = HostSim(sc.model)
            prob, targets, _ = sc.problem()
            v, st = hs.solve_ik(prob, sc.q32, targets)
            vg, sg = hs.solve_ik(prob, sc.q32, targets, path=1)
            v_ref, st_ref = sc.oracle_solve()
            feas = st_ref == 0
            mis = int(((st & 1) != 0)[feas].sum() + ((st & 1) == 0)[~feas].sum())
            mis_g = int(((sg & 1) != 0)[feas].sum() + ((sg & 1) == 0)[~feas].sum())
            both, bothg = feas & (st == 0), feas & (sg == 0)
            ok = helpers.within_tolerance(v[both], v_ref[both], atol=5e-4, rtol=5e-3)
            okg = helpers.within_tolerance(vg[bothg], v_ref[bothg], atol=5e-4, rtol=5e-3)
            print(f"{name} seed {seed} feasible {feas.mean():.2f} feasibility verdicts differing from the oracle: "
                  f"tree {mis} general {mis_g}; off tolerance: tree {int((~ok).sum())} general {int((~okg).sum())}; "
                  f"iteration caps hit: {int(((st & 8) != 0).sum())} / {int(((sg & 8) != 0).sum())}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--fma", action="store_true", help="host build with -ffp-contract=fast -mfma (the GPU contracts a*b+c into FMAs; "
                    "the default host build does not): a second rounding pattern for the same source")
    ap.add_argument("--wide", action="store_true", help="also: 10 more UR5 seeds and every chain instantiation on random chains")
    ap.add_argument("--lanes", type=int, default=0, help="also soak the sub-warp chain kernel body with this many lanes per instance (UR5 part)")
    a = ap.parse_args()
    if a.fma:
        import ctypes
        import os
        import subprocess

        import tests.hostsim as hsmod

        so = "/tmp/libpk_hostsim_fma.so"
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=fast", "-mfma",
                               "-o", so, os.path.join(os.path.dirname(os.path.abspath(hsmod.__file__)), "hostsim.cpp")])
        hsmod._lib = ctypes.CDLL(so)
        hsmod._lib.hs_last_error.restype = ctypes.c_char_p
        print("# host build with FMA contraction (-ffp-contract=fast -mfma)")
    t0 = time.time()
    ur5_kkt(4000 if a.quick else 40000)
    if a.lanes:
        ur5_kkt(4000 if a.quick else 40000, path=10 + a.lanes)
    if a.wide:
        ur5_wide(4000 if a.quick else 40000)
        chains(2000 if a.quick else 20000)
    humanoids(60 if a.quick else 300)
    with_extras(80 if a.quick else 400, 40 if a.quick else 200)
    print(f"# {time.time() - t0:.0f} s")
