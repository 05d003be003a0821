"""Base class of kinematic limits (``/root/reference/pink/limits/limit.py:17-45``)."""

import abc
from typing import Optional, Tuple


class Limit(abc.ABC):
    """Abstract base class for kinematic limits."""

    @abc.abstractmethod
    def compute_qp_inequalities(self, configuration, dt: float) -> Optional[Tuple]:
        r"""Pair :math:`(G, h)` with :math:`G \Delta q \leq h`, or ``None``.

        ``G`` is ``[m, nv]`` (numpy, the same for every instance); ``h`` is
        ``[m]`` numpy for a single configuration and a ``[B, m]`` device tensor
        for a batched one.
        """
