"""CPU suite: barriers, equality constraints, opt-in limits and constant-Jacobian
tasks through the kernel body of the general path (fp32, host build) against the
fp64 oracle."""

import numpy as np

from pink_b200 import _cabi
from tests import extras, helpers
from tests.hostsim import HostSim


def _check_rows(sc, hs, prob, targets, n):
    G, hG, E, f, lo, hi = hs.constraint_rows(prob, sc.q32[:n], None if targets is None else targets[:n])
    H32, c32, _ = hs.build_ik(prob, sc.q32[:n], None if targets is None else targets[:n])
    nv = sc.table.nv
    for i in range(n):
        H, c, Go, ho, A, b = sc.oracle_assemble(i)
        # objective (barrier safe-displacement terms included)
        assert np.abs(H32[i] - H).max() <= 2e-4 * np.abs(H).max()
        assert np.abs(c32[i] - c).max() <= 2e-4 * (np.abs(c).max() + 1e-3)
        # dense rows = the oracle rows that are not +-e_i, in order
        dense = [r for r in range(Go.shape[0]) if np.count_nonzero(Go[r]) != 1 or abs(abs(Go[r]).max() - 1.0) > 1e-12]
        p = len(dense)
        if sc.obarriers and sc.obarriers[-1]["type"] == "self_collision":
            # closest-pair rows may come in any order: compare as sorted by right-hand side
            k = sc.obarriers[-1]["n_pairs"]
            head, tail = dense[: p - k], dense[p - k:]
            order_o = list(head) + [tail[j] for j in np.argsort(ho[tail], kind="stable")]
            order_k = list(range(p - k)) + [p - k + j for j in np.argsort(hG[i, p - k:p], kind="stable")]
        else:
            order_o, order_k = dense, list(range(p))
        assert np.isinf(hG[i, p:]).all() and not G[i, p:].any()
        scale = np.abs(Go[order_o]).max() + 1e-9
        assert np.abs(G[i, order_k] - Go[order_o]).max() <= 5e-4 * scale, i
        assert np.abs(hG[i, order_k] - ho[order_o]).max() <= 5e-4 * (np.abs(ho[order_o]).max() + 1e-3), i
        if A is not None:
            m = A.shape[0]
            assert np.abs(E[i, :m] - A).max() <= 1e-5 and np.abs(f[i, :m] - b).max() <= 1e-5
        # box = intersection of all +-e_i rows
        hi_o = np.full(nv, np.inf)
        lo_o = np.full(nv, -np.inf)
        for r in range(Go.shape[0]):
            if r in dense:
                continue
            j = int(np.nonzero(Go[r])[0][0])
            if Go[r, j] > 0:
                hi_o[j] = min(hi_o[j], ho[r])
            else:
                lo_o[j] = max(lo_o[j], -ho[r])
        fin = np.isfinite(hi_o)
        assert np.abs(hi[i][fin] - hi_o[fin]).max() <= 1e-6 and np.abs(lo[i][fin] - lo_o[fin]).max() <= 1e-6


def test_ur5_rows_match_oracle():
    sc = extras.ur5_extras(24)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    _check_rows(sc, hs, prob, targets, 24)


def test_ur5_barriers_constraints_limits_match_oracle():
    sc = extras.ur5_extras(300)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree and not hs.used_chain  # barriers + an equality constraint: warp kernel
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)  # thread-per-instance general path
    np.testing.assert_array_equal(st, st_gen)
    v_ref, st_ref = sc.oracle_solve()
    feasible = st_ref == 0
    # infeasible QPs (barrier already violated and unreachable within the velocity box) are flagged alike
    assert ((st & _cabi.PK_STATUS_NO_SOLUTION) != 0)[~feasible].all()
    assert (st[feasible] == 0).all()
    assert feasible.mean() > 0.5
    ok = helpers.within_tolerance(v[feasible], v_ref[feasible])
    assert ok.all(), f"{(~ok).sum()} of {feasible.sum()} off, worst {np.abs(v - v_ref)[feasible].max()}"
    assert helpers.within_tolerance(v_gen[feasible], v_ref[feasible]).all()
    # the barrier rows matter: the same problem without them moves differently
    sc2 = extras.ur5_extras(300, active=False)
    v2_ref, _ = sc2.oracle_solve(60)
    assert np.abs(v2_ref - v_ref[:60])[feasible[:60]].max() > 1e-2


def test_g1_rows_match_oracle():
    sc = extras.g1_extras(6)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    _check_rows(sc, hs, prob, targets, 6)


def test_g1_self_collision_barrier_config_matches_oracle():
    """Config 4 of BASELINE.json: G1-class humanoid, CoM + frame tasks + sphere
    self-collision barrier (+ floating-base limit and a joint coupling task)."""
    sc = extras.g1_extras(64)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree  # floating-base rows and barrier rows: dual QP of the warp kernel
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)  # thread-per-instance general path
    np.testing.assert_array_equal(st, st_gen)
    v_ref, st_ref = sc.oracle_solve()
    feasible = st_ref == 0
    assert feasible.mean() > 0.8
    assert (st[feasible] == 0).all()
    assert helpers.within_tolerance(v_gen[feasible], v_ref[feasible]).all()
    # infeasible QPs (penetrating spheres that cannot separate within the limits) are flagged
    assert ((st & _cabi.PK_STATUS_NO_SOLUTION) != 0)[~feasible].all() and not feasible.all()
    ok = helpers.within_tolerance(v[feasible], v_ref[feasible])
    assert ok.all(), f"{(~ok).sum()} off, worst {np.abs(v - v_ref)[feasible].max()}"


def test_joint_coupling_task_runs_on_the_tree_kernel_body():
    """examples/humanoid_draco3.py:94-107: JointCouplingTasks next to frame and posture
    tasks stay on the warp-cooperative kernel (no barriers / constraints involved)."""
    import torch

    from pink_b200 import JointCouplingTask

    sc = helpers.humanoid_scenario("draco3_description", 48)

    class _Cfg:
        model = sc.model

    names = [n for n in sc.table.joint_names if "knee" in n or "hip_pitch" in n][:4]
    assert len(names) == 4
    jc1 = JointCouplingTask(names[:2], [1.0, -1.0], 100.0, _Cfg())
    jc2 = JointCouplingTask(names[2:], [1.0, -0.5], 50.0, _Cfg(), gain=0.7, lm_damping=1e-3)
    tasks = sc.tasks + [jc1, jc2]
    otasks = sc.oracle_tasks + [
        {"type": "linear", "A": jc1.A, "b": np.zeros(1), "q0": None, "cost": np.full(1, 100.0), "gain": 1.0, "lm_damping": 0.0},
        {"type": "linear", "A": jc2.A, "b": np.zeros(1), "q0": None, "cost": np.full(1, 50.0), "gain": 0.7, "lm_damping": 1e-3},
    ]
    from pink_b200.solve_ik import describe_problem
    from oracle import ik as oik

    prob, parts, _ = describe_problem(sc.model, sc.B, tasks, sc.dt, sc.damping, sc.limits, sc.safety_break)
    targets = torch.cat([p.cpu().float() for p in parts], dim=1).numpy()
    hs = HostSim(sc.model)
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)
    v_ref, st_ref = oik.solve_ik_batch(sc.table, sc.q64, otasks, sc.dt, sc.damping, sc.oracle_limits, sc.safety_break)
    assert (st == 0).all() and (st_gen == 0).all() and (st_ref == 0).all()
    assert helpers.within_tolerance(v, v_ref).all(), np.abs(v - v_ref).max()
    assert helpers.within_tolerance(v_gen, v_ref).all(), np.abs(v_gen - v_ref).max()
    # the coupling matters
    v_plain, _ = oik.solve_ik_batch(sc.table, sc.q64[:8], [oik._slice_task_range(t, 0, 8) for t in sc.oracle_tasks],
                                    sc.dt, sc.damping, sc.oracle_limits, sc.safety_break)
    assert np.abs(v_plain - v_ref[:8]).max() > 1e-3


def test_acceleration_limit_stays_on_the_chain_and_tree_kernels():
    """AccelerationLimit is a box (pink/limits/acceleration_limit.py:119-200): with
    per-instance previous velocities it runs inside the chain kernel (UR5) and the tree
    kernel (G1-class), and all three paths agree with the oracle."""
    import torch

    from pink_b200.limits import AccelerationLimit, ConfigurationLimit, VelocityLimit
    from pink_b200.solve_ik import describe_problem
    from oracle import ik as oik

    for name, B in [("ur5", 200), ("g1", 24)]:
        if name == "ur5":
            sc = helpers.ur5_scenario(B, "reachable")
            a_max = np.array([40.0, 40.0, 60.0, np.inf, 80.0, 80.0])
        else:
            sc = helpers.humanoid_scenario("g1_description", B, with_com=True)
            a_max = np.concatenate([np.full(6, np.inf), np.full(sc.table.nv - 6, 150.0)])
        rng = np.random.default_rng(1)
        v_prev = (rng.normal(size=(B, sc.table.nv)) * 0.3).astype(np.float32)
        acc = AccelerationLimit(sc.model, a_max)
        acc.set_last_integration(torch.as_tensor(v_prev), sc.dt)
        limits = [ConfigurationLimit(sc.model), VelocityLimit(sc.model), acc]
        prob, parts, _ = describe_problem(sc.model, B, sc.tasks, sc.dt, sc.damping, limits, sc.safety_break)
        targets = torch.cat([p.cpu().float() for p in parts], dim=1).numpy()
        hs = HostSim(sc.model)
        v, st = hs.solve_ik(prob, sc.q32, targets)
        assert hs.used_chain if name == "ur5" else hs.used_tree
        v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)
        dq_prev = (torch.as_tensor(v_prev) * sc.dt).numpy().astype(np.float64)
        v_ref, st_ref = oik.solve_ik_batch(sc.table, sc.q64, sc.oracle_tasks, sc.dt, sc.damping,
                                           [("configuration", 0.5), ("velocity", None), ("acceleration", a_max, dq_prev)],
                                           sc.safety_break)
        feasible = st_ref == 0
        assert feasible.mean() > 0.7
        assert (st[feasible] == 0).all() and (st_gen[feasible] == 0).all()
        assert ((st[~feasible] & _cabi.PK_STATUS_NO_SOLUTION) != 0).all()
        tol = dict(atol=5e-4, rtol=5e-3) if name == "g1" else {}
        assert helpers.within_tolerance(v[feasible], v_ref[feasible], **tol).all(), np.abs(v - v_ref)[feasible].max()
        assert helpers.within_tolerance(v_gen[feasible], v_ref[feasible], **tol).all()
        # the limit matters
        v_plain, _ = oik.solve_ik_batch(sc.table, sc.q64[:16], [oik._slice_task_range(t, 0, 16) for t in sc.oracle_tasks],
                                        sc.dt, sc.damping, None, sc.safety_break)
        assert np.abs(v_plain - v_ref[:16])[feasible[:16]].max() > 1e-2


def test_frame_and_com_tasks_as_equality_constraints():
    """solve_ik(..., constraints=[FrameTask, ComTask]) (pink/solve_ik.py:125-149): nine
    equality rows J dq = -gain e on the G1-class model, next to the other tasks."""
    import torch

    from oracle import ik as oik
    from pink_b200.solve_ik import describe_problem

    sc = helpers.humanoid_scenario("g1_description", 24, with_com=True)
    # constraints: left foot (frame task, gain 0.5) and the CoM task; the rest stays in the objective
    foot = [i for i, t in enumerate(sc.tasks) if getattr(t, "frame", "") == "left_ankle_roll_link"][0]
    com = [i for i, t in enumerate(sc.tasks) if type(t).__name__ == "ComTask"][0]
    sc.tasks[foot].gain = 0.05
    sc.oracle_tasks[foot]["gain"] = 0.05
    sc.tasks[com].gain = 0.05
    sc.oracle_tasks[com]["gain"] = 0.05
    cons, ocons = [sc.tasks[foot], sc.tasks[com]], [sc.oracle_tasks[foot], sc.oracle_tasks[com]]
    tasks = [t for i, t in enumerate(sc.tasks) if i not in (foot, com)]
    otasks = [t for i, t in enumerate(sc.oracle_tasks) if i not in (foot, com)]
    prob, parts, _ = describe_problem(sc.model, sc.B, tasks, sc.dt, sc.damping, sc.limits, sc.safety_break, None, cons)
    targets = torch.cat([p.cpu().float() for p in parts], dim=1).numpy()
    hs = HostSim(sc.model)
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree  # equality rows in the dual QP of the warp kernel
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)
    np.testing.assert_array_equal(st, st_gen)
    v_ref, st_ref = oik.solve_ik_batch(sc.table, sc.q64, otasks, sc.dt, sc.damping, sc.oracle_limits, sc.safety_break,
                                       None, ocons)
    feasible = st_ref == 0
    assert feasible.mean() > 0.5 and (st[feasible] == 0).all()
    assert ((st[~feasible] & _cabi.PK_STATUS_NO_SOLUTION) != 0).all()
    assert helpers.within_tolerance(v[feasible], v_ref[feasible], atol=5e-4, rtol=5e-3).all(), \
        np.abs(v - v_ref)[feasible].max()
    assert helpers.within_tolerance(v_gen[feasible], v_ref[feasible], atol=5e-4, rtol=5e-3).all()
    # the equalities hold in the exported rows: E dq = f
    _, _, E, f, _, _ = hs.constraint_rows(prob, sc.q32, targets)
    x = v.astype(np.float64) * sc.dt
    res = np.abs(np.einsum("brn,bn->br", E[:, :9].astype(np.float64), x) - f[:, :9])
    assert res[feasible].max() < 2e-6


def test_barriers_run_on_the_tree_kernel_body():
    """Barriers without equality constraints / floating-base limit stay on the
    warp-cooperative kernel (dual active-set QP in shared memory, pk_treedual.cuh):
    config 4 of BASELINE.json, and a UR5 with position + distance barriers."""
    import torch

    from oracle import ik as oik
    from pink_b200.solve_ik import describe_problem

    sc = extras.g1_extras(64, floating_base_limit=False)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree
    v_ref, st_ref = sc.oracle_solve()
    feasible = st_ref == 0
    assert feasible.mean() > 0.8 and not feasible.all()
    assert (st[feasible] == 0).all() and ((st[~feasible] & _cabi.PK_STATUS_NO_SOLUTION) != 0).all()
    assert not v[~feasible].any()
    assert helpers.within_tolerance(v[feasible], v_ref[feasible]).all(), np.abs(v - v_ref)[feasible].max()
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)
    np.testing.assert_array_equal(st, st_gen)
    np.testing.assert_allclose(v[feasible], v_gen[feasible], rtol=2e-3, atol=2e-4)

    su = extras.ur5_extras(200)
    prob, parts, _ = describe_problem(su.model, su.B, su.tasks, su.dt, su.damping, su.limits, su.safety_break,
                                      su.barriers, None, su.collision_model)
    targets = torch.cat([p.cpu().float() for p in parts], dim=1).numpy()
    hs = HostSim(su.model)
    v, st = hs.solve_ik(prob, su.q32, targets)
    assert hs.used_tree and not hs.used_chain
    v_ref, st_ref = oik.solve_ik_batch(su.table, su.q64, su.otasks, su.dt, su.damping, su.olimits, su.safety_break,
                                       su.obarriers, [])
    feasible = st_ref == 0
    assert feasible.mean() > 0.5
    assert (st[feasible] == 0).all() and ((st[~feasible] & _cabi.PK_STATUS_NO_SOLUTION) != 0).all()
    assert helpers.within_tolerance(v[feasible], v_ref[feasible]).all(), np.abs(v - v_ref)[feasible].max()


def test_random_trees_with_barriers_constraints_and_base_limit():
    """Barriers, equality constraints and the floating-base limit on topologies other than UR5 /
    G1: random trees with prismatic joints, fixed and floating base, on the warp-cooperative
    dual QP and on the general path, against the oracle."""
    for nj, free_flyer, seed in [(7, False, 7), (12, True, 8), (20, False, 9), (28, True, 7)]:
        sc = extras.tree_extras(nj, 40, free_flyer, seed=seed)
        hs = HostSim(sc.model)
        prob, targets, _ = sc.problem()
        v, st = hs.solve_ik(prob, sc.q32, targets)
        assert hs.used_tree
        v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, path=1)
        v_ref, st_ref = sc.oracle_solve()
        feasible = st_ref == 0
        assert feasible.mean() > 0.5
        for vk, sk in ((v, st), (v_gen, st_gen)):
            np.testing.assert_array_equal((sk & _cabi.PK_STATUS_NO_SOLUTION) != 0, ~feasible)
            assert (sk[feasible] == 0).all()
            assert helpers.within_tolerance(vk[feasible], v_ref[feasible], atol=5e-4, rtol=5e-3).all()
    _check_rows(sc, hs, prob, targets, 6)
