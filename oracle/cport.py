"""ctypes wrapper of the oracle's C port (``oracle/c/pink_oracle.c``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): a second, independent fp64
checker (dense H, Goldfarb-Idnani with Givens updates - the reference's own
formulation) fast enough for full-size batches, and the CPU baseline of
``bench.py``.  Fixed-base trees, FrameTask + PostureTask, default limits.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_HERE, "libpink_oracle.so")
MAXJ, MAXT = 16, 4


class OcModel(C.Structure):
    _fields_ = [
        ("njoints", C.c_int),
        ("parent", C.POINTER(C.c_int)),
        ("jtype", C.POINTER(C.c_int)),
        ("joint_placement", C.POINTER(C.c_double)),
        ("axis", C.POINTER(C.c_double)),
        ("q_min", C.POINTER(C.c_double)),
        ("q_max", C.POINTER(C.c_double)),
        ("v_max", C.POINTER(C.c_double)),
    ]


class OcProblem(C.Structure):
    _fields_ = [
        ("n_frame_tasks", C.c_int),
        ("frame_body", C.c_int * MAXT),
        ("frame_placement", (C.c_double * 12) * MAXT),
        ("frame_cost", (C.c_double * 6) * MAXT),
        ("frame_gain", C.c_double * MAXT),
        ("frame_lm", C.c_double * MAXT),
        ("has_posture", C.c_int),
        ("posture_cost", C.c_double),
        ("posture_gain", C.c_double),
        ("posture_lm", C.c_double),
        ("posture_target", C.c_double * MAXJ),
        ("dt", C.c_double),
        ("damping", C.c_double),
        ("cfg_gain", C.c_double),
        ("use_cfg_limit", C.c_int),
        ("use_vel_limit", C.c_int),
        ("safety_break", C.c_int),
    ]


_lib = None


def build():
    src = os.path.join(_HERE, "pink_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class CPort:
    """Holds the arrays of one (model table, task set)."""

    def __init__(self, table, tasks, dt, damping=1e-12, limits=None, safety_break=True):
        assert not table.free_flyer and table.njoints <= MAXJ
        nj = table.njoints
        d = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self._arrays = dict(
            parent=np.ascontiguousarray(table.parent, dtype=np.int32),
            jtype=np.ascontiguousarray(table.jtype, dtype=np.int32),
            jp=d(np.concatenate([np.asarray(table.joint_R).reshape(nj, 3, 3), np.asarray(table.joint_p).reshape(nj, 3, 1)], axis=2)),
            axis=d(table.axis), q_min=d(table.q_min), q_max=d(table.q_max), v_max=d(table.v_max),
        )
        a = self._arrays
        p = lambda x, t: x.ctypes.data_as(C.POINTER(t))  # noqa: E731
        self.model = OcModel(nj, p(a["parent"], C.c_int), p(a["jtype"], C.c_int), p(a["jp"], C.c_double),
                             p(a["axis"], C.c_double), p(a["q_min"], C.c_double), p(a["q_max"], C.c_double),
                             p(a["v_max"], C.c_double))
        P = OcProblem()
        k = 0
        for t in tasks:
            if t["type"] == "frame":
                f = t["frame"]
                P.frame_body[k] = int(table.frame_body[f])
                X = np.concatenate([np.asarray(table.frame_R[f]), np.asarray(table.frame_p[f]).reshape(3, 1)], axis=1).reshape(12)
                for i in range(12):
                    P.frame_placement[k][i] = X[i]
                cost = np.broadcast_to(np.asarray(t.get("cost", 1.0), dtype=float), (6,))
                for i in range(6):
                    P.frame_cost[k][i] = cost[i]
                P.frame_gain[k] = t.get("gain", 1.0)
                P.frame_lm[k] = t.get("lm_damping", 0.0)
                k += 1
            elif t["type"] == "posture":
                P.has_posture = 1
                P.posture_cost = float(t["cost"])
                P.posture_gain = t.get("gain", 1.0)
                P.posture_lm = t.get("lm_damping", 0.0)
                for i, x in enumerate(np.asarray(t["target"], dtype=float)):
                    P.posture_target[i] = x
            else:
                raise ValueError("the C port covers frame and posture tasks only")
        P.n_frame_tasks = k
        P.dt, P.damping = dt, damping
        P.cfg_gain = 0.5
        P.use_cfg_limit = P.use_vel_limit = 1
        if limits is not None:
            P.use_cfg_limit = P.use_vel_limit = 0
            for kind, arg in limits:
                if kind == "configuration":
                    P.use_cfg_limit, P.cfg_gain = 1, arg
                elif kind == "velocity":
                    P.use_vel_limit = 1
        P.safety_break = 1 if safety_break else 0
        self.problem = P
        self.nft = k

    def solve(self, q, targets, threads=1, reuse_outputs=False):
        """``q [B, nj]``, ``targets [B, nft, 3, 4]`` (fp64) -> ``v [B, nj]``, ``status [B]``."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        B = q.shape[0]
        t = np.ascontiguousarray(np.asarray(targets, dtype=np.float64).reshape(B, 12 * self.nft))
        if reuse_outputs:
            # repeated timing calls: no 3 MB allocation + page faults per step
            if getattr(self, "_out", None) is None or self._out[0].shape != q.shape:
                self._out = (np.zeros_like(q), np.zeros(B, dtype=np.int32))
            v, st = self._out
        else:
            v = np.zeros_like(q)
            st = np.zeros(B, dtype=np.int32)
        rc = lib().oc_solve_ik_batch(C.byref(self.model), C.byref(self.problem), q.ctypes.data_as(C.c_void_p),
                                     t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                     st.ctypes.data_as(C.c_void_p), C.c_int64(B), int(threads))
        assert rc == 0
        return v, st
