"""Configuration limit (``/root/reference/pink/limits/configuration_limit.py``)."""

from typing import Optional, Tuple

import numpy as np

from .limit import Limit, range_limited_coordinates, select_joints, selection_matrix


class ConfigurationLimit(Limit):
    r"""Rows :math:`\pm P \Delta q \leq g\,(q_{lim} \ominus q)`.

    Attributes mirror the reference: ``config_limit_gain``, ``indices``,
    ``joints``, ``model``, ``projection_matrix``.
    """

    def __init__(self, model, config_limit_gain: float = 0.5):
        assert 0.0 < config_limit_gain <= 1.0
        # selection at construction time (configuration_limit.py:50-72)
        self.joints, self.indices = select_joints(model, range_limited_coordinates(model), "q")
        self.config_limit_gain = config_limit_gain
        self.model = model
        self.projection_matrix = selection_matrix(model.nv, self.indices)

    def box_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        """Per-tangent-index position bounds (+-inf where no row exists), read
        from the model's *current* limit vectors (configuration_limit.py:111-116)."""
        nv = self.model.nv
        lo = np.full(nv, -np.inf)
        hi = np.full(nv, np.inf)
        if len(self.indices) > 0:
            shift = self.model.nq - nv
            lo[self.indices] = self.model.lowerPositionLimit[self.indices + shift]
            hi[self.indices] = self.model.upperPositionLimit[self.indices + shift]
        return lo, hi

    def compute_qp_inequalities(self, configuration, dt: float) -> Optional[Tuple]:
        """``configuration_limit.py:82-121``, evaluated by the CUDA library."""
        if self.projection_matrix is None:
            return None
        from ..solve_ik import _limit_rows

        return _limit_rows(configuration, [self], dt)
