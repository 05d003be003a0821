"""Multi-GPU plumbing: one process per GPU, batch sharded across ranks.

IK instances are independent, so the data path has NO collective: rank ``r`` of
``N`` solves its contiguous shard and keeps ``v`` sharded (the consumer -
integration, the next step - is data-parallel too).  NCCL over NVLink is used
for exactly two things (BASELINE north_star):

* :func:`broadcast_model` - model constants from rank 0, once;
* :func:`gather_velocities` / :func:`all_gather_velocities` - collecting ``v``
  when a single consumer needs the whole batch (plain NCCL collectives);
* :class:`PeerGather` - the same all-gather FUSED into the solve kernel: every rank's
  kernel stores its velocity rows straight into the gather buffers of all peers over
  NVLink peer memory (``pk_solve_ik_prepared_gather``); produced / released counters in peer
  memory replace the collective's synchronisation.  NCCL only carries the IPC handles, once.
* :func:`solve_ik_all_ranks` - one call that shards a batch over the box, solves and
  returns the whole ``v`` on every rank.

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
    """Broadcast the kinematic model (joint tree, placements, limits, frames, inertias: a few
    KB) from ``src`` to every rank; returns the model.  What travels is the model's flat image
    (:meth:`pink_b200.model.Model.pack`): one float64 array, one int64 array and the UTF-8 names -
    three collectives after one for the sizes, no pickled objects.  Objects without ``pack``
    (tests pass strings) fall back to a pickle broadcast."""
    rank, world = _world()
    if world == 1:
        return model
    device = torch.device("cpu") if device is None else torch.device(device)
    if dist.get_backend() == "gloo":
        device = torch.device("cpu")
    packable = torch.tensor([1 if (rank == src and hasattr(model, "pack")) else 0], dtype=torch.int64, device=device)
    dist.broadcast(packable, src)
    if not int(packable.item()):
        return _broadcast_pickled(model, device, src)
    if rank == src:
        floats, ints, text = model.pack()
        sizes = torch.tensor([floats.size, ints.size, len(text)], dtype=torch.int64, device=device)
    else:
        sizes = torch.zeros(3, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src)
    nf, ni, nt = (int(x) for x in sizes.cpu())
    if rank == src:
        bufs = [torch.as_tensor(floats, dtype=torch.float64).to(device), torch.as_tensor(ints, dtype=torch.int64).to(device),
                torch.frombuffer(bytearray(text), dtype=torch.uint8).to(device)]
    else:
        bufs = [torch.empty(nf, dtype=torch.float64, device=device), torch.empty(ni, dtype=torch.int64, device=device),
                torch.empty(nt, dtype=torch.uint8, device=device)]
    for buf in bufs:
        dist.broadcast(buf, src)
    if rank == src:
        return model
    from .model import Model

    return Model.unpack(bufs[0].cpu().numpy(), bufs[1].cpu().numpy(), bufs[2].cpu().numpy().tobytes())


def _broadcast_pickled(obj, device, src: int):
    rank, _ = _world()
    if rank == src:
        getattr(obj, "__dict__", {}).pop("_pk_engines", None)  # device handles do not travel
        payload = pickle.dumps(obj)
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
        obj = pickle.load(io.BytesIO(buf.cpu().numpy().tobytes()))
    return obj


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


class PeerGather:
    """All-gather of ``v`` through NVLink peer memory, written by the solve kernel itself.

    Every rank owns ``n_buffers`` gather buffers ``[world * B, nv]`` (rotated per call) and
    one flag block; the buffers of all ranks are mapped into every process through CUDA IPC
    handles exchanged once over the process group.  :meth:`solve` runs the IK step of this
    rank's shard with the gather fused into the kernel epilogue: the kernel stores its rows
    into slot ``k % n_buffers`` of every rank (after checking, in its first instructions, that
    every rank has released gather ``k - n_buffers``); :meth:`wait` queues a one-warp kernel that
    publishes "gather k complete", blocks its stream until gather ``k`` has arrived from every
    rank and releases the slot.  ``solve(..., wait=True)`` (default) does both on the current stream;
    for throughput, call ``solve(..., wait=False)`` on the compute stream and :meth:`wait` on
    the stream that consumes the gathered velocities - the next solve then overlaps the
    NVLink transfer of this one.  Every rank must issue the same sequence of ``solve`` calls
    and exactly one ``wait`` per ``solve``.  Equal shard sizes ``B`` on all ranks.
    """

    def __init__(self, shard_rows: int, nv: int, device, n_buffers: int = 2):
        import ctypes as C

        from . import _cabi

        rank, world = _world()
        if world > _cabi.PK_MAX_PEERS:
            raise ValueError(f"at most {_cabi.PK_MAX_PEERS} ranks")
        self.rank, self.world, self.B, self.nv = rank, world, int(shard_rows), int(nv)
        self.device = torch.device(device)
        self.lib = _cabi.load()
        self.n_buffers = n_buffers
        self._call = 0
        self._waited = 0
        dev = self.device.index
        nbytes = self.world * self.B * self.nv * 4
        # local allocations: n_buffers gather buffers + one flag block
        self._local, handles = [], []
        for k in range(n_buffers + 1):
            ptr = C.c_void_p()
            h = C.create_string_buffer(_cabi.PK_IPC_HANDLE_BYTES)
            size = nbytes if k < n_buffers else 4 * _cabi.PK_PEER_FLAG_WORDS
            _cabi.check(self.lib.pk_peer_alloc(dev, size, C.byref(ptr), h))
            self._local.append(ptr)
            handles.append(h.raw)
        # exchange the handles (64 bytes each) over the process group
        on_gpu = dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        mine = torch.tensor(list(b"".join(handles)), dtype=torch.uint8, device=self.device if on_gpu else "cpu")
        if world > 1:
            everyone = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(everyone, mine)
            everyone = bytes(everyone.cpu().numpy().tobytes())
        else:
            everyone = bytes(mine.cpu().numpy().tobytes())
        per = (n_buffers + 1) * _cabi.PK_IPC_HANDLE_BYTES
        self._opened = []
        # ptrs[k][r]: buffer k of rank r as seen from this process (k = n_buffers: flag block)
        self.ptrs = [[None] * world for _ in range(n_buffers + 1)]
        for r in range(world):
            for k in range(n_buffers + 1):
                if r == rank:
                    self.ptrs[k][r] = self._local[k]
                    continue
                h = everyone[r * per + k * _cabi.PK_IPC_HANDLE_BYTES: r * per + (k + 1) * _cabi.PK_IPC_HANDLE_BYTES]
                ptr = C.c_void_p()
                _cabi.check(self.lib.pk_peer_open(dev, h, C.byref(ptr)))
                self._opened.append(ptr)
                self.ptrs[k][r] = ptr
        self._arrays = [(C.c_void_p * world)(*[p.value for p in self.ptrs[k]]) for k in range(n_buffers + 1)]
        # torch views of the local gather buffers
        self.views = [self._view(self._local[k], (self.world * self.B, self.nv)) for k in range(n_buffers)]
        if world > 1:
            dist.barrier()  # everyone has mapped everyone before the first store

    def _view(self, ptr, shape):
        class _Holder:
            pass

        hld = _Holder()
        hld.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": "<f4", "data": (int(ptr.value), False), "version": 3, "strides": None,
        }
        with torch.cuda.device(self.device):
            return torch.as_tensor(hld, device=self.device)

    def solve(self, ik, q: torch.Tensor, targets, status=None, v_local=None, wait: bool = True):
        """IK step of this rank's shard with the fused gather.  Returns ``(v_all, status)``:
        ``v_all`` is this rank's ``[world * B, nv]`` buffer of the current rotation slot,
        complete once the matching :meth:`wait` has run (``wait=True``: on the current stream)."""
        from .engine import _addr, _stream
        from . import _cabi

        eng = ik.engine
        B = q.shape[0]
        if B != self.B:
            raise ValueError(f"shard has {B} rows, PeerGather was built for {self.B}")
        if status is None:
            status = torch.empty((B,), device=eng.device, dtype=torch.int32)
        if v_local is None:
            # the shard's own rows, also kept locally (kernels without the fused epilogue solve
            # into this buffer and scatter from it)
            if getattr(self, "_v_local", None) is None:
                self._v_local = torch.empty((B, self.nv), device=eng.device, dtype=torch.float32)
            v_local = self._v_local
        k = self._call % self.n_buffers
        self._call += 1
        with torch.cuda.device(eng.device):
            _cabi.check(self.lib.pk_solve_ik_prepared_gather(
                eng.handle, ik._handle, _addr(q), _addr(targets), _addr(v_local), _addr(status), B,
                self._arrays[k], self.world, self.rank * self.B,
                self._arrays[self.n_buffers], self.rank, self.n_buffers, _stream(eng.device)))
        if wait:
            self.wait()
        return self.views[k], status

    def wait(self, release: bool = True):
        """Queue, on the current stream, the one-warp kernel that publishes this rank's oldest
        unpublished gather, waits until that gather has arrived from every rank and (``release``)
        frees its buffer slot for the gather ``n_buffers`` calls later.  The current stream must
        be ordered after the matching :meth:`solve` (same stream, or ``wait_stream``)."""
        from .engine import _stream
        from . import _cabi

        if self._waited >= self._call:
            raise RuntimeError("PeerGather.wait without a matching solve")
        self._waited += 1
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_peer_sync(self.device.index, self._arrays[self.n_buffers], self.world, self.rank,
                                              1, 1, 1 if release else 0, _stream(self.device)))

    def timeouts(self) -> int:
        """Number of spin-wait time-outs recorded in this rank's flag block (0 in a healthy run)."""
        torch.cuda.synchronize(self.device)
        flags = self._view(self._local[self.n_buffers], (48,)).view(torch.int32)
        return int(flags[35].item())

    def close(self):
        dev = self.device.index
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier()  # nobody still writes into a buffer that is about to go
        for p in self._opened:
            self.lib.pk_peer_close(dev, p)
        for p in self._local:
            self.lib.pk_peer_free(dev, p)
        self._opened, self._local = [], []

    def __del__(self):
        try:
            if self._local:
                self.close()
        except Exception:
            pass


def solve_ik_all_ranks(ik, q_global, targets_global, gather: Optional["PeerGather"] = None):
    """Shard ``q_global [B, nq]`` / ``targets_global [B, stride]`` (host or device tensors,
    identical on every rank) over the ranks, solve, and return the full ``v [B, nv]`` on every
    rank (``B`` divisible by the world size).  ``ik`` is this rank's :class:`BatchedIK`.
    With a :class:`PeerGather` the gather is fused into the kernel; without, NCCL all-gather."""
    rank, world = _world()
    lo, hi = shard_bounds(q_global.shape[0])
    dev = ik.engine.device
    q = q_global[lo:hi].to(dev, dtype=torch.float32).contiguous()
    t = None if targets_global is None else targets_global[lo:hi].to(dev, dtype=torch.float32).contiguous()
    if gather is not None and world > 1:
        v_all, _ = gather.solve(ik, q, t)
        return v_all
    v, _ = ik.solve(q, t)
    return all_gather_velocities(v)
