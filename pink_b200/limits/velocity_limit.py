"""Velocity limit (``/root/reference/pink/limits/velocity_limit.py``)."""

from typing import Optional, Tuple

import numpy as np

from ..exceptions import PinkError
from .limit import Limit, magnitude_limited, select_joints, selection_matrix


class VelocityLimit(Limit):
    r"""Rows :math:`\pm P \Delta q \leq \mathrm{d}t\, v_{max}`."""

    def __init__(self, model, velocity_limit: Optional[np.ndarray] = None):
        if velocity_limit is None:
            velocity_limit = model.velocityLimit
        else:
            velocity_limit = np.asarray(velocity_limit, dtype=float).flatten()
            if model.nv > 0 and velocity_limit.shape[0] != model.nv:
                raise PinkError(f"{velocity_limit.shape=} but {model.nv=}")
        self.joints, self.indices = select_joints(model, magnitude_limited(velocity_limit), "v")
        self.model = model
        self.projection_matrix = selection_matrix(model.nv, self.indices)
        self.velocity_limit = velocity_limit

    def box_bounds(self) -> np.ndarray:
        """Per-tangent-index velocity bound (+inf where no row exists)."""
        v = np.full(self.model.nv, np.inf)
        if len(self.indices) > 0:
            v[self.indices] = self.velocity_limit[self.indices]
        return v

    def compute_qp_inequalities(self, configuration, dt: float) -> Optional[Tuple]:
        """``velocity_limit.py:90-121``, evaluated by the CUDA library."""
        if self.projection_matrix is None:
            return None
        if configuration is None:
            # the rows are constants of (model, dt), no configuration to evaluate
            # (tests/test_velocity_limit.py:46-55 calls it that way)
            v_max = self.velocity_limit[self.indices]
            return (np.vstack([self.projection_matrix, -self.projection_matrix]), dt * np.hstack([v_max, v_max]))
        from ..solve_ik import _limit_rows

        return _limit_rows(configuration, [self], dt)
