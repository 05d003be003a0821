"""Frame position barrier (``/root/reference/pink/barriers/position_barrier.py``)."""

from typing import List, Optional, Union

import numpy as np

from .._cabi import PK_BARRIER_POSITION
from ..exceptions import NoPositionLimitProvided
from .barrier import Barrier


class PositionBarrier(Barrier):
    r"""Keep selected world coordinates of a frame within ``[p_min, p_max]``:
    ``h = [p[idx] - p_min; p_max - p[idx]]`` (``position_barrier.py:95-153``)."""

    def __init__(self, frame: str, indices: Optional[List[int]] = None, p_min: Optional[np.ndarray] = None,
                 p_max: Optional[np.ndarray] = None, gain: Union[float, np.ndarray] = 1.0,
                 safe_displacement_gain: float = 0.0):
        indices = [0, 1, 2] if indices is None else indices
        if p_min is None and p_max is None:
            raise NoPositionLimitProvided(f"Position barrier for frame {frame} requires either p_min or p_max")
        dim = 0
        if p_min is not None:
            dim += len(indices)
        if p_max is not None:
            dim += len(indices)
        if isinstance(gain, np.ndarray) and len(gain) != dim:
            gain = np.tile(gain, 2)
        super().__init__(dim, gain=gain, safe_displacement_gain=safe_displacement_gain)
        self.indices = indices
        self.frame = frame
        self.p_min = p_min
        self.p_max = p_max

    def _pk_describe(self, model) -> dict:
        n = len(self.indices)
        return {
            "type": PK_BARRIER_POSITION,
            "frame": model.getFrameId(self.frame),
            "dim": self.dim,
            "indices": list(self.indices),
            "p_min": None if self.p_min is None else np.asarray(self.p_min, dtype=float).reshape(n),
            "p_max": None if self.p_max is None else np.asarray(self.p_max, dtype=float).reshape(n),
            "gain": np.asarray(self.gain, dtype=float),
        }
