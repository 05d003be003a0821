"""N > 1 plumbing on CPU: world_size 2 over gloo (shards, constant broadcast, gather)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pink_b200 import parallel
        from pink_b200.robots import load_robot_description

        model = load_robot_description("ur5_description").model if rank == 0 else None
        model = parallel.broadcast_model(model)
        assert model.nq == 6 and model.getFrameId("tool0") < len(model.frames)
        table = model.table()
        # what arrived (flat image, no pickle) is the model rank 0 holds: same tables, same frame ids
        local = load_robot_description("ur5_description").model.table()
        for key, val in vars(local).items():
            assert np.array_equal(np.asarray(val), np.asarray(getattr(table, key))), key
        B = 11
        lo, hi = parallel.shard_bounds(B)
        v_local = torch.arange(lo, hi, dtype=torch.float32).repeat_interleave(6).reshape(hi - lo, 6)
        full = parallel.gather_velocities(v_local, dst=0)
        eq = parallel.all_gather_velocities(torch.full((3, 6), float(rank)))
        if rank == 0:
            out.put((lo, hi, full.numpy(), eq.numpy(), float(table.q_max[0])))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_shard_broadcast_and_gather_world_size_2():
    from pink_b200 import parallel

    assert [parallel.shard_bounds(11, r, 2) for r in range(2)] == [(0, 6), (6, 11)]
    assert [parallel.shard_bounds(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    lo, hi, full, eq, qmax = out.get()
    assert (lo, hi) == (0, 6)
    np.testing.assert_array_equal(full[:, 0], np.arange(11, dtype=np.float32))
    np.testing.assert_array_equal(eq[:, 0], [0, 0, 0, 1, 1, 1])
    assert abs(qmax - 2 * np.pi) < 1e-12


def test_single_process_paths_are_identity():
    from pink_b200 import parallel

    v = torch.ones(4, 6)
    assert parallel.all_gather_velocities(v) is v and parallel.gather_velocities(v) is v
    assert parallel.shard_bounds(10) == (0, 10)
    assert parallel.broadcast_model("m") == "m"
