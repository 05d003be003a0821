"""Base class of kinematic limits (``/root/reference/pink/limits/limit.py:17-45``)."""

import abc
from typing import List, Optional, Tuple

import numpy as np


def range_limited_coordinates(model) -> np.ndarray:
    """``bool[nq]``: configuration coordinates with a finite, non-degenerate position range
    (the selection rule of ``configuration_limit.py:50-56``)."""
    upper, lower = model.upperPositionLimit, model.lowerPositionLimit
    return model.hasConfigurationLimit() & (upper < 1e20) & (upper > lower + 1e-10)


def magnitude_limited(bound: np.ndarray) -> np.ndarray:
    """``bool[nv]``: tangent coordinates whose velocity / acceleration bound is finite and
    non-zero (``velocity_limit.py:57-63``)."""
    return (bound < 1e20) & (bound > 1e-10)


def select_joints(model, flags: np.ndarray, space: str) -> Tuple[List, np.ndarray]:
    """Joints all of whose coordinates are flagged (``space`` = ``"q"``: flags over the
    configuration vector, ``"v"``: over the tangent space) and the read-only array of
    their tangent indices."""
    chosen, tangent = [], []
    for joint in model.joints:
        start, width = (joint.idx_q, joint.nq) if space == "q" else (joint.idx_v, joint.nv)
        if start < 0 or not flags[start:start + width].all():
            continue
        chosen.append(joint)
        tangent.extend(range(joint.idx_v, joint.idx_v + joint.nv))
    indices = np.array(tangent, dtype=np.int64)
    indices.setflags(write=False)
    return chosen, indices


def selection_matrix(nv: int, indices: np.ndarray) -> Optional[np.ndarray]:
    """Rows of the identity picked by ``indices`` (``None`` when there are none)."""
    return np.eye(nv)[indices] if len(indices) > 0 else None


class Limit(abc.ABC):
    """Abstract base class for kinematic limits."""

    @abc.abstractmethod
    def compute_qp_inequalities(self, configuration, dt: float) -> Optional[Tuple]:
        r"""Pair :math:`(G, h)` with :math:`G \Delta q \leq h`, or ``None``.

        ``G`` is ``[m, nv]`` (numpy, the same for every instance); ``h`` is
        ``[m]`` numpy for a single configuration and a ``[B, m]`` device tensor
        for a batched one.
        """
