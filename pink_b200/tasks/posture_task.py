"""Posture task (``/root/reference/pink/tasks/posture_task.py``)."""

from typing import Optional

import numpy as np

from .._cabi import PK_TASK_POSTURE
from ..exceptions import TargetNotSet, TaskDefinitionError
from ..utils import get_root_joint_dim
from ._targets import as_vector_target
from .task import Task


class PostureTask(Task):
    r"""Regulate joint angles to a desired posture.

    ``e = (q (-) q*)[root_nv:]``, ``J = I[root_nv:, :]``
    (``posture_task.py:100-129``): floating-base coordinates are not affected.
    """

    target_q: Optional[object]

    def __init__(self, cost: float, lm_damping: float = 0.0, gain: float = 1.0) -> None:
        super().__init__(cost=cost, gain=gain, lm_damping=lm_damping)
        self.target_q = None

    def set_target(self, target_q) -> None:
        """``[nq]`` for all instances or ``[B, nq]`` per instance (copied)."""
        self.target_q = as_vector_target(target_q, np.shape(target_q)[-1])

    def set_target_from_configuration(self, configuration) -> None:
        self.set_target(configuration.q)

    def _pk_describe(self, model) -> dict:
        if self.target_q is None:
            raise TargetNotSet("no posture target")
        if not isinstance(self.cost, (float, int)):
            raise TaskDefinitionError(f"Posture task cost should be a scalar, currently cost={self.cost}")
        if np.shape(self.target_q)[-1] != model.nq:
            raise TaskDefinitionError(f"posture target has {np.shape(self.target_q)[-1]} coordinates, model has nq={model.nq}")
        _, root_nv = get_root_joint_dim(model)
        cost6 = np.zeros(6)
        cost6[0] = float(self.cost)
        return {
            "type": PK_TASK_POSTURE,
            "frame": 0,
            "root": 0,
            "cost6": cost6,
            "k": model.nv - root_nv,
            "target": self.target_q,
        }

    def __repr__(self):
        return f"PostureTask(cost={self.cost}, gain={self.gain}, lm_damping={self.lm_damping})"
