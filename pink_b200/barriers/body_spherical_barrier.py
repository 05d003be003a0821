"""Two-frame distance barrier
(``/root/reference/pink/barriers/body_spherical_barrier.py``)."""

from typing import Tuple, Union

import numpy as np

from .._cabi import PK_BARRIER_BODY_SPHERICAL, PK_GAINFN_SATURATING
from ..exceptions import NegativeMinimumDistance
from .barrier import Barrier


class BodySphericalBarrier(Barrier):
    r"""``h = |p_1 - p_2|^2 - d_min^2`` with the class-K function
    ``h / (1 + |h|)`` (``body_spherical_barrier.py:54-143``)."""

    gain_function_id = PK_GAINFN_SATURATING

    def __init__(self, frames: Tuple[str, str], d_min: float, gain: Union[float, np.ndarray] = 1.0,
                 safe_displacement_gain: float = 3.0):
        if d_min < 0.0:
            raise NegativeMinimumDistance("The minimum distance threshold must be non-negative.")
        super().__init__(dim=1, gain=gain, safe_displacement_gain=safe_displacement_gain)
        self.frames = frames
        self.d_min = d_min

    def _pk_describe(self, model) -> dict:
        return {
            "type": PK_BARRIER_BODY_SPHERICAL,
            "frame": model.getFrameId(self.frames[0]),
            "frame2": model.getFrameId(self.frames[1]),
            "dim": 1,
            "d_min": float(self.d_min),
            "gain": np.asarray(self.gain, dtype=float),
        }

    def compute_jacobian(self, configuration):
        """``dh/dq`` as the reference returns it for this barrier: the row itself, ``[nv]``
        (``body_spherical_barrier.py:139-143``: ``dh_dx.T @ dx_dq``); ``[B, 1, nv]`` batched."""
        J = super().compute_jacobian(configuration)
        return J if getattr(configuration, "batched", False) else J.reshape(-1)
