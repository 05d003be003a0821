"""pk_solve_ik_prepared_gather + pk_peer_barrier on one GPU (world size 1: the rank's own
buffer is its only peer).  The multi-rank run is scripts/peer_gather_check.py under torchrun
and the `solve_plus_allgather.fused` leg of bench.py."""

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu


def _ik(sc, B):
    from pink_b200 import BatchedIK

    return BatchedIK(sc.model, sc.tasks, sc.dt, damping=sc.damping, limits=sc.limits, device="cuda", batch_size=B)


@pytest.mark.parametrize("scenario", ["ur5", "g1"])
def test_fused_gather_equals_plain_solve(scenario):
    from pink_b200 import parallel

    if scenario == "ur5":
        sc = helpers.ur5_scenario(4096, "reachable")
    else:
        sc = helpers.humanoid_scenario("g1_description", 256, with_com=True)
    B = sc.B
    ik = _ik(sc, B)
    _, targets, _ = sc.problem()
    q = torch.as_tensor(sc.q32, device="cuda")
    t = torch.as_tensor(targets, device="cuda")
    v_ref, st_ref = ik.solve(q, t)
    peer = parallel.PeerGather(B, ik.nv, torch.device("cuda", 0), n_buffers=2)
    try:
        seen = []
        for _ in range(3):  # rotates through both buffers
            v_all, st = peer.solve(ik, q, t)
            torch.cuda.synchronize()
            assert v_all.shape == (B, ik.nv)
            assert torch.equal(v_all, v_ref)
            assert torch.equal(st, st_ref)
            seen.append(v_all.data_ptr())
        assert seen[0] != seen[1] and seen[0] == seen[2]
    finally:
        peer.close()


def test_solve_ik_all_ranks_single_rank():
    from pink_b200 import parallel

    sc = helpers.ur5_scenario(512, "reachable")
    ik = _ik(sc, 512)
    _, targets, _ = sc.problem()
    v = parallel.solve_ik_all_ranks(ik, torch.as_tensor(sc.q32), torch.as_tensor(targets))
    v_ref, _ = sc.oracle_solve()
    assert helpers.within_tolerance(v.cpu().numpy(), v_ref).all()
