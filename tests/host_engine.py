"""TEST HARNESS ONLY: an object with the interface of ``pink_b200.engine.Engine``
whose entry points run the host build of the kernel bodies (tests/hostsim) on CPU
tensors.  Patched in by the ``host_engine`` fixture so that the CPU suite can drive
the whole Python drop-in layer (problem description, target packing, row stacking,
exception mapping) without a GPU.  The product package never sees this module."""

import numpy as np
import pytest
import torch

from tests.hostsim import HostSim


class HostEngine:
    def __init__(self, model):
        self.hs = HostSim(model)
        self.device = torch.device("cpu")
        self.table = self.hs.table
        self.nq, self.nv, self.nframes = self.hs.nq, self.hs.nv, self.hs.nframes
        self.root_nq, self.root_nv = (7, 6) if self.table.free_flyer else (0, 0)

    @staticmethod
    def _np(t):
        return None if t is None else t.detach().cpu().numpy()

    def solve_ik(self, prob, q, targets, v=None, status=None):
        vv, st = self.hs.solve_ik(prob, self._np(q), self._np(targets))
        vv, st = torch.as_tensor(vv), torch.as_tensor(st)
        if v is not None:
            v.copy_(vv)
            vv = v
        return vv, st

    def build_ik(self, prob, q, targets):
        return tuple(torch.as_tensor(a) for a in self.hs.build_ik(prob, self._np(q), self._np(targets)))

    def constraint_rows(self, prob, q, targets):
        return tuple(torch.as_tensor(a) for a in self.hs.constraint_rows(prob, self._np(q), self._np(targets)))

    def task_terms(self, prob, task_index, k, q, targets):
        return tuple(torch.as_tensor(a) for a in self.hs.task_terms(prob, task_index, k, self._np(q), self._np(targets)))

    def forward_kinematics(self, q, want_com=False):
        oMf, com = self.hs.forward_kinematics(self._np(q))
        return torch.as_tensor(oMf), (torch.as_tensor(com) if want_com else None)

    def frame_jacobian(self, frame, q):
        return torch.as_tensor(self.hs.frame_jacobian(frame, self._np(q)))

    def _f32(self, t, cols):
        t = torch.tensor(np.asarray(t), dtype=torch.float32)
        t = t.unsqueeze(0) if t.dim() == 1 else t
        assert t.shape[1] == cols
        return t.contiguous()

    def integrate(self, q, v, dt, out=None):
        res = torch.as_tensor(self.hs.integrate(self._np(q), self._np(v), float(dt)))
        if out is not None:
            out.copy_(res)
            return out
        return res


@pytest.fixture
def host_engine(monkeypatch):
    """Route Configuration.engine to the host build for the duration of a test."""
    import pink_b200.configuration as cfgmod

    cache = {}

    def get_engine(model, device=None):
        key = (id(model), len(model.frames))
        if key not in cache:
            cache[key] = HostEngine(model)
        return cache[key]

    monkeypatch.setattr(cfgmod, "get_engine", get_engine)
    return get_engine
