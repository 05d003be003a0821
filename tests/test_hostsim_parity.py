"""CPU suite: the kernel bodies (fp32, compiled for the host by tests/hostsim)
against the fp64 oracle.  These run the same source the GPU runs, so logic and
numerics of the CUDA path are covered without a device; the `-m gpu` twins in
test_gpu_parity.py repeat them through the real library."""

import numpy as np
import pytest

from oracle import ik as oik
from tests import helpers
from tests.hostsim import HostSim


@pytest.mark.parametrize("kind", ["reachable", "unreachable", "at_target"])
def test_ur5_chain_kernel_matches_oracle(kind):
    sc = helpers.ur5_scenario(400, kind)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_chain, "UR5 frame+posture must take the register-resident chain kernel"
    v_ref, st_ref = sc.oracle_solve()
    assert (st == 0).all() and (st_ref == 0).all()
    ok = helpers.within_tolerance(v, v_ref)
    if kind == "at_target":
        # e ~ 1e-7 is at the fp32 resolution of forward kinematics and H = J^T J + 1e-6 I
        # is as ill-conditioned as it gets (lm_damping inert at zero error): the noise
        # floor is |dv| ~ 2e-7 |J^-1| / dt, see DESIGN.md "Numerics"
        assert ok.mean() >= 0.97
        assert helpers.within_tolerance(v, v_ref, atol=5e-3, rtol=2e-2).all()
    else:
        assert ok.all(), f"{(~ok).sum()} instances off, worst {np.abs(v - v_ref).max()}"


def test_ur5_general_path_agrees_with_chain_kernel():
    sc = helpers.ur5_scenario(300, "reachable")
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v1, s1 = hs.solve_ik(prob, sc.q32, targets)
    v2, s2 = hs.solve_ik(prob, sc.q32, targets, general_path=True)
    assert not hs.used_chain
    np.testing.assert_array_equal(s1, s2)
    # same mathematics, different linear algebra (Cholesky + refinement vs Householder QR,
    # folded joint constants): agreement to fp32 rounding of the solve
    np.testing.assert_allclose(v1, v2, rtol=5e-4, atol=5e-5)


def test_ur5_full_batch_kkt_certificate():
    """Size-independent property at a large batch: the fp32 solution satisfies
    the fp64 KKT conditions of its own QP (unique minimiser => parity)."""
    sc = helpers.ur5_scenario(20000, "reachable")
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert (st == 0).all()
    H, c, G, h = sc.oracle_build()
    x = v.astype(np.float64) * sc.dt
    stat, prim, lo, hi = oik.kkt_check_batch(H, c, G, h, x)
    scale = np.abs(c).max(axis=1)
    assert prim.max() <= 1e-6
    assert np.quantile(stat / scale, 0.999) <= 1e-4
    assert (stat / scale).max() <= 1e-3


@pytest.mark.parametrize("seed", [104, 107])
def test_ur5_refinement_that_leaves_the_box_activates_the_bound(seed):
    """No Levenberg-Marquardt term, dt = 0.1: one instance of each of these 40000 reaches the
    Cholesky rounds with an ill-conditioned free block whose fp32 refinement leaves the box.
    Clamping alone stalled at a relative stationarity of 4e-2 (found by the round-2 soak,
    profiles/r02i_hostsim_soak.txt); a coordinate left on a bound must become active."""
    sc = helpers.ur5_scenario(40000, "reachable", seed=seed, lm_damping=0.0, posture_cost=1e-3)
    sc.dt = 0.1
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert (st == 0).all()
    H, c, G, h = sc.oracle_build()
    stat, prim, _, _ = oik.kkt_check_batch(H, c, G, h, v.astype(np.float64) * sc.dt)
    assert prim.max() <= 1e-6
    assert (stat / np.abs(c).max(axis=1)).max() <= 1e-4


def test_ur5_out_of_limits_and_safety_break():
    sc = helpers.ur5_scenario(200, "reachable", out_of_limits=7)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    v_ref, st_ref = sc.oracle_solve()
    np.testing.assert_array_equal(st & 3, st_ref)
    assert (st == 2).sum() == 7
    assert np.abs(v[st == 2]).max() == 0.0
    ok = helpers.within_tolerance(v, v_ref)
    assert ok.all()
    # safety_break=False: flagged but solved (configuration.py:195-201)
    sc.safety_break = False
    prob, targets, _ = sc.problem()
    v2, st2 = hs.solve_ik(prob, sc.q32, targets)
    v2_ref, st2_ref = sc.oracle_solve()
    flagged = (st2 & 2) != 0
    assert flagged.sum() == 7
    # instances too far outside have an empty box: NoSolutionFound in both
    np.testing.assert_array_equal((st2 & 1) != 0, st2_ref == 1)
    solved = (st2 & 1) == 0
    assert helpers.within_tolerance(v2[solved], v2_ref[solved]).all()


def test_ur5_no_limits_is_unconstrained_minimiser():
    sc = helpers.ur5_scenario(100, "reachable")
    sc.limits, sc.oracle_limits = [], []
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    H, c, G, h = sc.oracle_build()
    assert G is None
    x_ref = -np.linalg.solve(H, c[..., None])[..., 0]
    np.testing.assert_allclose(v * sc.dt, x_ref, rtol=2e-3, atol=2e-6)


@pytest.mark.parametrize("name,kw", [
    ("draco3_description", {}),
    ("g1_description", {"with_com": True}),
    ("draco3_description", {"with_relative": True}),
])
def test_humanoid_general_path_matches_oracle(name, kw):
    sc = helpers.humanoid_scenario(name, 24, **kw)
    hs = HostSim(sc.model)
    prob, targets, descs = sc.problem()
    # objective and limit rows
    H, c, h4 = hs.build_ik(prob, sc.q32, targets)
    H_ref, c_ref, G_ref, h_ref = sc.oracle_build()
    scale = np.abs(H_ref).max()
    np.testing.assert_allclose(H, H_ref, atol=2e-5 * scale, rtol=1e-4)
    np.testing.assert_allclose(c, c_ref, atol=2e-5 * np.abs(c_ref).max(), rtol=1e-4)
    # per-task error / Jacobian
    from oracle import kinematics as okin, tasks as otk

    fk = okin.forward_kinematics(sc.table, sc.q64)
    for k, (ot, d) in enumerate(zip(sc.oracle_tasks, descs)):
        e, J = hs.task_terms(prob, k, d["k"], sc.q32, targets)
        e_ref, J_ref = otk.task_error_jacobian(sc.table, sc.q64, fk, ot)
        np.testing.assert_allclose(e, e_ref, atol=5e-6, rtol=1e-5)
        np.testing.assert_allclose(J, np.broadcast_to(J_ref, J.shape), atol=1e-5, rtol=1e-5)
    # velocities
    v, st = hs.solve_ik(prob, sc.q32, targets)
    v_ref, st_ref = sc.oracle_solve()
    np.testing.assert_array_equal(st & 1, st_ref & 1)
    good = (st & 1) == 0
    helpers.parity_by_condition(v[good], v_ref[good], H_ref[good])


def test_forward_kinematics_and_frame_jacobian_exports():
    from oracle import kinematics as okin

    for name in ["ur5_description", "g1_description"]:
        robot, model, table = helpers.load(name)
        rng = np.random.default_rng(3)
        from pink_b200 import workloads

        q = workloads.sample_configurations(table, 16, rng)
        hs = HostSim(model)
        oMf, com = hs.forward_kinematics(q)
        fk = okin.forward_kinematics(table, q.astype(np.float32).astype(np.float64))
        for f in range(table.nframes):
            R, p = okin.frame_placement(table, fk, f)
            np.testing.assert_allclose(oMf[:, f, :, :3], R, atol=3e-6)
            np.testing.assert_allclose(oMf[:, f, :, 3], p, atol=3e-6)
        if table.mass.sum() > 0:
            np.testing.assert_allclose(com, okin.center_of_mass(table, fk), atol=3e-6)
        f = table.nframes - 1
        J = hs.frame_jacobian(f, q)
        np.testing.assert_allclose(J, okin.frame_jacobian_local(table, fk, f), atol=5e-6)


@pytest.mark.parametrize("name,kw", [
    ("draco3_description", {}),
    ("g1_description", {"with_com": True}),
    ("g1_description", {"with_com": True, "with_relative": True}),
])
def test_tree_kernel_body_matches_general_path_and_oracle(name, kw):
    """The warp-cooperative tree kernel (pk_tree.cuh), run lane by lane on the host."""
    sc = helpers.humanoid_scenario(name, 32, **kw)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v_t, st_t = hs.solve_ik(prob, sc.q32, targets, path=2)
    assert hs.used_tree
    v_g, st_g = hs.solve_ik(prob, sc.q32, targets, path=1)
    np.testing.assert_array_equal(st_t, st_g)
    np.testing.assert_allclose(v_t, v_g, atol=2e-3, rtol=2e-3)
    v_ref, st_ref = sc.oracle_solve()
    good = (st_t & 1) == 0
    ok = helpers.within_tolerance(v_t[good], v_ref[good], atol=5e-4, rtol=5e-3)
    assert ok.mean() >= 0.95


def test_tree_kernel_is_selected_for_humanoids_and_handles_limits():
    sc = helpers.humanoid_scenario("draco3_description", 16)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree and not hs.used_chain
    # out-of-limit instance with safety_break: flagged, zero velocity
    q = sc.q32.copy()
    q[3, 7 + 2] = sc.table.q_max[7 + 2] + 0.3
    sc.safety_break = True
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, q, targets)
    assert st[3] == 2 and np.abs(v[3]).max() == 0.0 and (st[np.arange(16) != 3] == 0).all()


@pytest.mark.parametrize("nj,kw", [
    (2, {}), (3, {"prismatic": (1,)}), (4, {"two_tasks": True}), (5, {"shared_target": True}),
    (7, {"two_tasks": True, "prismatic": (2,)}), (7, {}),
])
def test_chain_kernel_instantiations(nj, kw):
    """Every <NJ, NFT> instantiation of the register-resident kernel, prismatic joints,
    mid-chain frames and shared targets, against the oracle and the general path."""
    sc = helpers.chain_scenario(nj, 96, seed=nj, **kw)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_chain
    v_g, st_g = hs.solve_ik(prob, sc.q32, targets, path=1)
    np.testing.assert_array_equal(st, st_g)
    np.testing.assert_allclose(v, v_g, atol=5e-4, rtol=5e-3)
    v_ref, st_ref = sc.oracle_solve()
    np.testing.assert_array_equal(st & 3, st_ref)
    assert helpers.within_tolerance(v, v_ref, atol=5e-4, rtol=5e-3).mean() >= 0.97
    # the tree kernel body handles fixed-base chains too
    v_t, st_t = hs.solve_ik(prob, sc.q32, targets, path=2)
    np.testing.assert_allclose(v_t, v_g, atol=5e-4, rtol=5e-3)


def test_joint_velocity_and_damping_tasks():
    """JointVelocityTask / DampingTask (diagonal-only tasks)."""
    from pink_b200 import DampingTask, JointVelocityTask

    for name in ["ur5_description", "draco3_description"]:
        sc = helpers.ur5_scenario(64, "reachable") if name.startswith("ur5") else helpers.humanoid_scenario(name, 16)
        rv = 6 if sc.table.free_flyer else 0
        rng = np.random.default_rng(0)
        v_ref = rng.normal(size=sc.model.nv - rv) * 0.3
        jv = JointVelocityTask(cost=0.7)
        jv.set_target(v_ref, sc.dt)
        damp = DampingTask(cost=0.2)
        sc.tasks = sc.tasks + [jv, damp]
        sc.oracle_tasks = sc.oracle_tasks + [
            {"type": "joint_velocity", "cost": 0.7, "gain": 1.0, "lm_damping": 0.0, "target": v_ref * sc.dt},
            {"type": "joint_velocity", "cost": 0.2, "gain": 1.0, "lm_damping": 0.0, "target": np.zeros(sc.model.nv - rv)},
        ]
        hs = HostSim(sc.model)
        prob, targets, descs = sc.problem()
        v, st = hs.solve_ik(prob, sc.q32, targets)
        assert hs.used_tree  # diagonal-only extras are handled by the tree kernel body
        H, c, h4 = hs.build_ik(prob, sc.q32, targets)
        H_ref, c_ref, _, _ = sc.oracle_build()
        np.testing.assert_allclose(H, H_ref, atol=2e-5 * np.abs(H_ref).max(), rtol=1e-4)
        np.testing.assert_allclose(c, c_ref, atol=2e-5 * np.abs(c_ref).max(), rtol=1e-4)
        v_ref_o, st_ref = sc.oracle_solve()
        assert helpers.within_tolerance(v, v_ref_o, atol=5e-4, rtol=5e-3).mean() >= 0.97


@pytest.mark.parametrize("nj,free_flyer", [(40, False), (58, True)])
def test_models_beyond_the_warp_kernel_up_to_the_abi_maximum(nj, free_flyer):
    """Random joint trees with more than 32 joints (general path), up to the C-ABI
    maximum PK_MAX_JOINTS = 58 with a free-flyer (nv = PK_MAX_NV = 64)."""
    sc = helpers.tree_scenario(nj, 24, free_flyer)
    assert sc.table.nv == (nj + 6 if free_flyer else nj)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert not hs.used_tree and not hs.used_chain
    v_ref, st_ref = sc.oracle_solve()
    assert (st == 0).all() and (st_ref == 0).all()
    assert helpers.within_tolerance(v, v_ref, atol=5e-4, rtol=5e-3).all(), np.abs(v - v_ref).max()


def test_model_larger_than_the_abi_maximum_is_rejected():
    import pink_b200
    from tests import hostsim as hsmod

    rng = np.random.default_rng(0)
    model = helpers.random_tree_model(59, rng, free_flyer=True)
    with pytest.raises(RuntimeError):
        HostSim(model)


@pytest.mark.parametrize("name", ["ur5_description", "g1_description"])
def test_integration_body_matches_oracle(name):
    """q (+) v dt (Configuration.integrate, pink/configuration.py:273-283): the body of
    integrate_kernel against the oracle's SE(3) x R^n integration, small and large steps,
    in place as the closed-loop entry point uses it."""
    from oracle import kinematics as okin
    from pink_b200 import workloads

    robot, model, table = helpers.load(name)
    rng = np.random.default_rng(21)
    q = workloads.sample_configurations(table, 300, rng)
    hs = HostSim(model)
    for scale, dt in ((1.0, 5e-3), (30.0, 0.1), (1e-6, 1e-3)):
        v = rng.normal(size=(300, model.nv)) * scale
        out = hs.integrate(q, v, dt)
        ref = okin.integrate(table, q.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64) * np.float32(dt))
        if table.free_flyer:
            # quaternions up to sign
            sign = np.sign(np.sum(out[:, 3:7] * ref[:, 3:7], axis=1, keepdims=True))
            ref[:, 3:7] *= sign
            np.testing.assert_allclose(np.linalg.norm(out[:, 3:7], axis=1), 1.0, atol=3e-7)
        np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-6 * max(1.0, scale * dt))


@pytest.mark.parametrize("nj,free_flyer,seed", [(8, False, 3), (15, True, 4), (32, False, 5), (32, True, 3)])
def test_warp_kernel_on_random_trees(nj, free_flyer, seed):
    """Topologies other than the two humanoids on the warp-per-instance kernel: random joint
    trees with prismatic joints, fixed or floating base, up to its 32-joint maximum; frame,
    relative-frame, posture and CoM tasks."""
    sc = helpers.tree_scenario(nj, 24, free_flyer, seed=seed)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    assert hs.used_tree
    v_gen, st_gen = hs.solve_ik(prob, sc.q32, targets, general_path=True)
    v_ref, st_ref = sc.oracle_solve()
    assert (st == 0).all() and (st_gen == 0).all() and (st_ref == 0).all()
    assert helpers.within_tolerance(v, v_ref).all(), np.abs(v - v_ref).max()
    np.testing.assert_allclose(v, v_gen, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("name,B,kw", [
    ("draco3_description", 600, {}),
    ("g1_description", 500, {"with_com": True}),
])
def test_humanoid_parity_distribution_by_condition_number(name, B, kw):
    """Host build of the warp kernel: the binned parity statement of helpers.PARITY_BINS
    (>= 99.9 % inside 2e-4 + 2e-3 |v| for cond(H) < 1e5; the stated distribution above)."""
    sc = helpers.humanoid_scenario(name, B, **kw)
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v, st = hs.solve_ik(prob, sc.q32, targets)
    v_ref, st_ref = sc.oracle_solve()
    assert (st == 0).all() and (st_ref == 0).all()
    report = helpers.parity_by_condition(v, v_ref, sc.oracle_build()[0])
    assert sum(r["n"] for r in report) == B
