"""Acceleration limit (``/root/reference/pink/limits/acceleration_limit.py``)."""

from typing import Optional

import numpy as np

from ..tasks._targets import as_vector_target
from .limit import Limit, magnitude_limited, range_limited_coordinates, select_joints, selection_matrix


class AccelerationLimit(Limit):
    r"""Finite-difference acceleration bound plus braking distance to the
    configuration limits (``acceleration_limit.py:20-200``):

    .. math::

        \Delta q_i \leq \min(a\,dt^2 + \Delta q_{prev,i},\ dt \sqrt{2 a (q_{max} \ominus q)_i}), \quad
        -\Delta q_i \leq \min(a\,dt^2 - \Delta q_{prev,i},\ dt \sqrt{2 a (q \ominus q_{min})_i})

    ``Delta_q_prev`` is per instance in the batched engine
    (:meth:`set_last_integration` accepts ``[nv]`` or ``[B, nv]`` velocities).
    """

    def __init__(self, model, acceleration_limit: np.ndarray):
        from ..exceptions import PinkError

        acceleration_limit = np.asarray(acceleration_limit, dtype=float).flatten()
        if model.nv > 0 and acceleration_limit.shape[0] != model.nv:  # acceleration_limit.py:55-56
            raise PinkError(f"{acceleration_limit.shape=} but {model.nv=}")
        joints, indices = select_joints(model, magnitude_limited(acceleration_limit), "v")
        # per selected tangent coordinate: does its joint also have a position range (the
        # braking-distance term applies only there, acceleration_limit.py:88-100)
        ranged = range_limited_coordinates(model)
        with_range = [bool(ranged[j.idx_q:j.idx_q + j.nq].all()) for j in joints for _ in range(j.nv)]
        self._delta_q_prev_full = np.zeros(model.nv)
        self.a_max = acceleration_limit[indices] if len(indices) > 0 else np.empty(0)
        self.acceleration_limit = acceleration_limit
        self.has_configuration_limit = np.array(with_range, dtype=bool)
        self.indices = indices
        self.model = model
        self.projection_matrix = selection_matrix(model.nv, indices)

    def set_last_integration(self, v_prev, dt) -> None:
        """Latest integrated velocity (``[nv]`` or ``[B, nv]``) and its timestep."""
        self._delta_q_prev_full = as_vector_target(v_prev, self.model.nv) * dt

    @property
    def Delta_q_prev(self):
        """Last displacement on the limited coordinates only, ``[len(indices)]`` (or
        ``[B, len(indices)]``), as the reference stores it (``acceleration_limit.py:116-117``);
        the engine keeps the full ``[.., nv]`` vector internally."""
        return self._delta_q_prev_full[..., self.indices]

    @Delta_q_prev.setter
    def Delta_q_prev(self, value):
        value = np.asarray(value, dtype=float) if not hasattr(value, "dim") else value
        if value.shape[-1] == self.model.nv:
            self._delta_q_prev_full = as_vector_target(value, self.model.nv)
            return
        if value.shape[-1] != len(self.indices):
            raise ValueError(f"Delta_q_prev has {value.shape[-1]} entries, expected {len(self.indices)}")
        full = np.zeros(tuple(value.shape[:-1]) + (self.model.nv,))
        full[..., self.indices] = np.asarray(value, dtype=float)
        self._delta_q_prev_full = as_vector_target(full, self.model.nv)

    def box_arrays(self):
        """``(a_max, q_lo, q_hi)`` per tangent index (inf: no row / no braking term)."""
        nv = self.model.nv
        shift = self.model.nq - nv
        a = np.full(nv, np.inf)
        qlo = np.full(nv, -np.inf)
        qhi = np.full(nv, np.inf)
        for k, i in enumerate(self.indices):
            a[i] = self.a_max[k]
            if self.has_configuration_limit[k]:
                qlo[i] = self.model.lowerPositionLimit[i + shift]
                qhi[i] = self.model.upperPositionLimit[i + shift]
        return a, qlo, qhi

    def compute_qp_inequalities(self, configuration, dt: float):
        """``(G, h)`` with ``G = [P; -P]`` (``acceleration_limit.py:119-200``),
        evaluated by the CUDA library."""
        if self.projection_matrix is None:
            return None
        from ..solve_ik import _acceleration_rows

        return _acceleration_rows(configuration, self, dt)
