"""Self-collision barrier on sphere pairs
(``/root/reference/pink/barriers/self_collision_barrier.py``)."""

from typing import Union

import numpy as np

from .._cabi import PK_BARRIER_SELF_COLLISION
from ..exceptions import InvalidCollisionPairs, NegativeMinimumDistance
from .barrier import Barrier


class SelfCollisionBarrier(Barrier):
    r"""``h_i = d_i - d_min`` for the ``n_collision_pairs`` closest collision
    pairs of the configuration's :class:`pink_b200.SphereCollisionModel`
    (``self_collision_barrier.py:85-224``)."""

    def __init__(self, n_collision_pairs: int, gain: Union[float, np.ndarray] = 1.0,
                 safe_displacement_gain: float = 1.0, d_min: float = 0.02):
        if d_min < 0.0:
            raise NegativeMinimumDistance("The minimum distance threshold must be non-negative.")
        if n_collision_pairs < 0:
            raise InvalidCollisionPairs("The number of collision pairs must be non-negative.")
        super().__init__(dim=n_collision_pairs, gain=gain, safe_displacement_gain=safe_displacement_gain)
        self.d_min = d_min

    def _pk_describe(self, model, collision_model=None) -> dict:
        if collision_model is None:
            raise InvalidCollisionPairs("SelfCollisionBarrier needs a configuration with a collision model")
        if len(collision_model.collisionPairs) < self.dim:
            raise InvalidCollisionPairs(
                f"The number of collision pairs ({len(collision_model.collisionPairs)}) "
                f"is less than the barrier dimension ({self.dim})."
            )
        gain = np.asarray(self.gain, dtype=float)
        if gain.size and not np.all(gain == gain.flat[0]):
            raise NotImplementedError("per-pair gains: rows follow the closest pairs, use a scalar gain")
        return {
            "type": PK_BARRIER_SELF_COLLISION,
            "dim": self.dim,
            "d_min": float(self.d_min),
            "gain": gain[:1] if gain.size else np.ones(1),
            "pairs": collision_model.pair_frames(),
            "radii": collision_model.pair_radii(),
        }
