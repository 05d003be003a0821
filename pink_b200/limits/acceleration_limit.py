"""Acceleration limit (``/root/reference/pink/limits/acceleration_limit.py``)."""

from typing import List, Optional

import numpy as np

from ..tasks._targets import as_vector_target
from .limit import Limit


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
        acceleration_limit = np.asarray(acceleration_limit, dtype=float).flatten()
        has_acceleration_limit = np.logical_and(acceleration_limit < 1e20, acceleration_limit > 1e-10)
        joints = [
            joint
            for joint in model.joints
            if joint.idx_v >= 0
            and has_acceleration_limit[slice(joint.idx_v, joint.idx_v + joint.nv)].all()
        ]
        has_configuration_limit = np.logical_and(
            model.hasConfigurationLimit(),
            np.logical_and(
                model.upperPositionLimit < 1e20,
                model.upperPositionLimit > model.lowerPositionLimit + 1e-10,
            ),
        )
        index_list: List[int] = []
        config_limit_list: List[bool] = []
        for joint in joints:
            index_list.extend(range(joint.idx_v, joint.idx_v + joint.nv))
            joint_has_config_limit = bool(has_configuration_limit[slice(joint.idx_q, joint.idx_q + joint.nq)].all())
            config_limit_list.extend([joint_has_config_limit] * joint.nv)
        indices = np.array(index_list, dtype=np.int64)
        indices.setflags(write=False)
        dim = len(indices)
        self.Delta_q_prev = np.zeros(model.nv)
        self.a_max = acceleration_limit[indices] if dim > 0 else np.empty(0)
        self.acceleration_limit = acceleration_limit
        self.has_configuration_limit = np.array(config_limit_list, dtype=bool)
        self.indices = indices
        self.model = model
        self.projection_matrix = np.eye(model.nv)[indices] if dim > 0 else None

    def set_last_integration(self, v_prev, dt) -> None:
        """Latest integrated velocity (``[nv]`` or ``[B, nv]``) and its timestep."""
        self.Delta_q_prev = as_vector_target(v_prev, self.model.nv) * dt

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
