"""Configuration limit (``/root/reference/pink/limits/configuration_limit.py``)."""

from typing import List, Optional, Tuple

import numpy as np

from .limit import Limit


class ConfigurationLimit(Limit):
    r"""Rows :math:`\pm P \Delta q \leq g\,(q_{lim} \ominus q)`.

    Attributes mirror the reference: ``config_limit_gain``, ``indices``,
    ``joints``, ``model``, ``projection_matrix``.
    """

    def __init__(self, model, config_limit_gain: float = 0.5):
        assert 0.0 < config_limit_gain <= 1.0
        # selection at construction time (configuration_limit.py:50-72)
        has_configuration_limit = np.logical_and(
            model.hasConfigurationLimit(),
            np.logical_and(
                model.upperPositionLimit < 1e20,
                model.upperPositionLimit > model.lowerPositionLimit + 1e-10,
            ),
        )
        joints = [
            joint
            for joint in model.joints
            if joint.idx_q >= 0
            and has_configuration_limit[slice(joint.idx_q, joint.idx_q + joint.nq)].all()
        ]
        index_list: List[int] = []
        for joint in joints:
            index_list.extend(range(joint.idx_v, joint.idx_v + joint.nv))
        indices = np.array(index_list, dtype=np.int64)
        indices.setflags(write=False)
        dim = len(indices)
        self.config_limit_gain = config_limit_gain
        self.indices = indices
        self.joints = joints
        self.model = model
        self.projection_matrix = np.eye(model.nv)[indices] if dim > 0 else None

    def box_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        """Per-tangent-index position bounds (+-inf where no row exists), read
        from the model's *current* limit vectors (configuration_limit.py:111-116)."""
        nv = self.model.nv
        lo = np.full(nv, -np.inf)
        hi = np.full(nv, np.inf)
        shift = self.model.nq - nv
        for i in self.indices:
            lo[i] = self.model.lowerPositionLimit[i + shift]
            hi[i] = self.model.upperPositionLimit[i + shift]
        return lo, hi

    def compute_qp_inequalities(self, configuration, dt: float) -> Optional[Tuple]:
        """``configuration_limit.py:82-121``, evaluated by the CUDA library."""
        if self.projection_matrix is None:
            return None
        from ..solve_ik import _limit_rows

        return _limit_rows(configuration, [self], dt)
