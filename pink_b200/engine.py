"""Per-(model, device) handle on the CUDA library.

Everything numerical happens behind ``include/pink_b200.h``; this module only
owns the opaque ``PkModel*``, turns torch tensors into raw device addresses and
passes the current torch CUDA stream.  torch is used for device memory, streams
and (in ``parallel.py``) ``torch.distributed`` - plumbing, not arithmetic.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pink_b200 needs a CUDA device (sm_100a); it has no CPU fallback"
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"pink_b200 runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _addr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """Device-resident model constants + thin call wrappers."""

    def __init__(self, model, device: torch.device):
        self.lib = _cabi.load()
        self.device = device
        self.table = model.table()
        self.nq, self.nv = int(self.table.nq), int(self.table.nv)
        self.nframes = int(self.table.nframes)
        self.root_nq, self.root_nv = (7, 6) if self.table.free_flyer else (0, 0)
        self._holder = _cabi.ModelDescHolder(self.table)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(self.lib.pk_model_create(C.byref(self._holder.desc), device.index, C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pk_model_destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass

    # -- helpers -----------------------------------------------------------
    def _f32(self, t, cols: int) -> torch.Tensor:
        """``[B, cols]`` contiguous fp32 tensor on this device."""
        if not isinstance(t, torch.Tensor):
            t = torch.tensor(np.asarray(t))  # copy: the source may be read-only
        t = t.to(device=self.device, dtype=torch.float32)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        t = t.reshape(t.shape[0], -1)
        if t.shape[1] != cols:
            raise ValueError(f"expected {cols} columns, got {tuple(t.shape)}")
        return t.contiguous()

    # -- entry points --------------------------------------------------------
    def solve_ik(self, prob: _cabi.PkProblemDesc, q: torch.Tensor, targets: Optional[torch.Tensor],
                 v: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None):
        B = q.shape[0]
        if v is None:
            v = torch.empty((B, self.nv), device=self.device, dtype=torch.float32)
        if status is None:
            status = torch.empty((B,), device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_solve_ik_batched(
                self.handle, C.byref(prob), _addr(q), _addr(targets), _addr(v), _addr(status), B,
                _stream(self.device)))
        return v, status

    def solve_ik_host(self, prob: _cabi.PkProblemDesc, q: torch.Tensor, targets: Optional[torch.Tensor],
                      v: torch.Tensor, status: Optional[torch.Tensor]):
        """Host (ideally pinned) tensors in and out; copies run inside the library."""
        B = q.shape[0]
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_solve_ik_batched_host(
                self.handle, C.byref(prob), _addr(q), _addr(targets), _addr(v), _addr(status), B,
                _stream(self.device)))
        return v, status

    def build_ik(self, prob: _cabi.PkProblemDesc, q: torch.Tensor, targets: Optional[torch.Tensor]):
        B, nv = q.shape[0], self.nv
        H = torch.empty((B, nv, nv), device=self.device, dtype=torch.float32)
        c = torch.empty((B, nv), device=self.device, dtype=torch.float32)
        h = torch.empty((B, 4, nv), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_build_ik_batched(
                self.handle, C.byref(prob), _addr(q), _addr(targets), _addr(H), _addr(c), _addr(h), B,
                _stream(self.device)))
        return H, c, h

    def constraint_rows(self, prob: _cabi.PkProblemDesc, q: torch.Tensor, targets: Optional[torch.Tensor]):
        """Dense inequality rows ``(G, hG)``, equality rows ``(E, f)`` and the box
        ``(lo, hi)`` of every instance (``pk_constraint_rows_batched``)."""
        B, nv = q.shape[0], self.nv
        mk = lambda *shape: torch.empty(shape, device=self.device, dtype=torch.float32)
        G, hG = mk(B, _cabi.PK_MAX_INEQ_ROWS, nv), mk(B, _cabi.PK_MAX_INEQ_ROWS)
        E, f = mk(B, _cabi.PK_MAX_EQ_ROWS, nv), mk(B, _cabi.PK_MAX_EQ_ROWS)
        lo, hi = mk(B, nv), mk(B, nv)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_constraint_rows_batched(
                self.handle, C.byref(prob), _addr(q), _addr(targets), _addr(G), _addr(hG), _addr(E), _addr(f),
                _addr(lo), _addr(hi), B, _stream(self.device)))
        return G, hG, E, f, lo, hi

    def task_terms(self, prob: _cabi.PkProblemDesc, task_index: int, k: int, q: torch.Tensor,
                   targets: Optional[torch.Tensor]):
        B, nv = q.shape[0], self.nv
        e = torch.empty((B, k), device=self.device, dtype=torch.float32)
        J = torch.empty((B, k, nv), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_task_terms_batched(
                self.handle, C.byref(prob), task_index, _addr(q), _addr(targets), _addr(e), _addr(J), B,
                _stream(self.device)))
        return e, J

    def forward_kinematics(self, q: torch.Tensor, want_com: bool = False):
        B = q.shape[0]
        oMf = torch.empty((B, self.nframes, 3, 4), device=self.device, dtype=torch.float32)
        com = torch.empty((B, 3), device=self.device, dtype=torch.float32) if want_com else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_forward_kinematics_batched(
                self.handle, _addr(q), _addr(oMf), _addr(com), B, _stream(self.device)))
        return oMf, com

    def frame_jacobian(self, frame: int, q: torch.Tensor):
        B = q.shape[0]
        J = torch.empty((B, 6, self.nv), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_frame_jacobian_batched(
                self.handle, frame, _addr(q), _addr(J), B, _stream(self.device)))
        return J

    def integrate(self, q: torch.Tensor, v: torch.Tensor, dt: float, out: Optional[torch.Tensor] = None):
        B = q.shape[0]
        if out is None:
            out = torch.empty_like(q)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pk_integrate_batched(
                self.handle, _addr(q), _addr(v), float(dt), _addr(out), B, _stream(self.device)))
        return out


def get_engine(model, device=None) -> Engine:
    """Engine of ``model`` on ``device``, cached on the model object (rebuilt
    when joints / frames / inertias were added since)."""
    device = require_cuda(device)
    cache = model.__dict__.setdefault("_pk_engines", {})
    key = (device.index, getattr(model, "_version", 0))
    eng = cache.get(key)
    if eng is None:
        for k in [k for k in cache if k[0] == device.index]:
            del cache[k]
        eng = Engine(model, device)
        cache[key] = eng
    return eng
