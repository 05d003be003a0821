"""Task base class (``/root/reference/pink/tasks/task.py:23-171``).

``compute_error`` / ``compute_jacobian`` / ``compute_qp_objective`` are public
reference API and are evaluated by the CUDA library
(``pk_task_terms_batched`` / ``pk_build_ik_batched``); inside ``solve_ik`` they
are never called - the fused kernel computes them on chip.
"""

import abc
from typing import Optional, Sequence, Tuple, Union

import numpy as np


class Task(abc.ABC):
    r"""Abstract base class for kinematic tasks.

    Attributes:
        cost: cost vector with the same dimension as the error of the task.
        gain: Task gain :math:`\alpha \in [0, 1]`.
        lm_damping: Unitless scale of the Levenberg-Marquardt regularization.
    """

    cost: Optional[Union[float, Sequence[float], np.ndarray]]
    gain: float
    lm_damping: float

    def __init__(self, cost=None, gain: float = 1.0, lm_damping: float = 0.0):
        self.cost = cost
        self.gain = gain
        self.lm_damping = lm_damping

    # -- description for the C-ABI (PkTaskDesc) ---------------------------------
    @abc.abstractmethod
    def _pk_describe(self, model) -> dict:
        """``{"type", "frame", "root", "cost6", "k", "target"}``; ``target`` is a
        numpy array (shared by all instances) or a ``[B, n]`` tensor."""

    def _terms(self, configuration):
        from ..solve_ik import _task_terms

        return _task_terms(configuration, self)

    def compute_error(self, configuration):
        """Task error ``e(q)``: ``[k]`` numpy, or ``[B, k]`` tensor when batched."""
        return self._terms(configuration)[0]

    def compute_jacobian(self, configuration):
        """Task Jacobian ``J(q)``: ``[k, nv]`` numpy, or ``[B, k, nv]`` tensor."""
        return self._terms(configuration)[1]

    def compute_qp_objective(self, configuration) -> Tuple:
        r"""Pair :math:`(H, c)` of this task alone (``task.py:115-167``):
        ``H = (WJ)^T (WJ) + mu I``, ``c = -(W(-alpha e))^T (WJ)``."""
        from ..solve_ik import _task_objective

        return _task_objective(configuration, self)

    @abc.abstractmethod
    def __repr__(self):
        """Human-readable representation of the task."""
