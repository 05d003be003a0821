"""Multi-GPU plumbing: one process per GPU, batch sharded across ranks.

IK instances are independent, so the data path has NO collective: rank ``r`` of
``N`` solves its contiguous shard and keeps ``v`` sharded (the consumer -
integration, the next step - is data-parallel too).  NCCL over NVLink is used
for exactly two things (BASELINE north_star):

* :func:`broadcast_model` - model constants from rank 0, once;
* :func:`gather_velocities` / :func:`all_gather_velocities` - collecting ``v``
  when a single consumer needs the whole batch.

Works with any initialised ``torch.distributed`` backend (``nccl`` on GPUs,
``gloo`` in the CPU tests).
"""

from __future__ import annotations

import io
import pickle
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous shard ``[lo, hi)`` of ``total`` instances owned by ``rank``
    (the first ``total % world`` ranks take one extra instance)."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_model(model, device=None, src: int = 0):
    """Broadcast the kinematic model (joint tree, placements, limits, frames,
    inertias: a few KB) from ``src`` to every rank; returns the model."""
    rank, world = _world()
    if world == 1:
        return model
    device = torch.device("cpu") if device is None else torch.device(device)
    if dist.get_backend() == "gloo":
        device = torch.device("cpu")
    if rank == src:
        model.__dict__.pop("_pk_engines", None)  # device handles do not travel
        payload = pickle.dumps(model)
        size = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    if rank == src:
        buf = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    else:
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    if rank != src:
        model = pickle.load(io.BytesIO(buf.cpu().numpy().tobytes()))
    return model


def all_gather_velocities(v_local: torch.Tensor) -> torch.Tensor:
    """``[B/N, nv]`` per rank -> ``[B, nv]`` on every rank (equal shards)."""
    rank, world = _world()
    if world == 1:
        return v_local
    out = torch.empty((world * v_local.shape[0],) + tuple(v_local.shape[1:]), dtype=v_local.dtype,
                      device=v_local.device)
    dist.all_gather_into_tensor(out, v_local.contiguous())
    return out


def gather_velocities(v_local: torch.Tensor, dst: int = 0) -> Optional[torch.Tensor]:
    """``[B_r, nv]`` per rank (shards may differ by one row) -> ``[B, nv]`` on
    ``dst``; ``None`` elsewhere."""
    rank, world = _world()
    if world == 1:
        return v_local
    counts = [torch.zeros(1, dtype=torch.int64, device=v_local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([v_local.shape[0]], dtype=torch.int64, device=v_local.device))
    counts = [int(c.item()) for c in counts]
    pad = max(counts)
    padded = torch.zeros((pad,) + tuple(v_local.shape[1:]), dtype=v_local.dtype, device=v_local.device)
    padded[: v_local.shape[0]] = v_local
    pieces = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, pieces, dst=dst)
    if rank != dst:
        return None
    return torch.cat([p[:n] for p, n in zip(pieces, counts)], dim=0)


def solve_ik_sharded(configuration_factory, tasks_factory, q_global, dt, **kwargs):
    """Convenience: shard ``q_global`` (host array ``[B, nq]``), solve the local
    shard and return ``(v_local, (lo, hi))``.  ``configuration_factory(q_shard)``
    and ``tasks_factory(lo, hi)`` build the per-rank objects."""
    from .solve_ik import solve_ik

    lo, hi = shard_bounds(q_global.shape[0])
    configuration = configuration_factory(q_global[lo:hi])
    v = solve_ik(configuration, tasks_factory(lo, hi), dt, **kwargs)
    return v, (lo, hi)
