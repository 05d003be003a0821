"""Joint-velocity and damping tasks
(``/root/reference/pink/tasks/joint_velocity_task.py``,
``/root/reference/pink/tasks/damping_task.py``)."""

from typing import Optional

import numpy as np

from .._cabi import PK_TASK_JOINT_VELOCITY
from ..exceptions import TargetNotSet, TaskDefinitionError
from ..utils import get_root_joint_dim
from ._targets import as_vector_target
from .task import Task


class JointVelocityTask(Task):
    r"""Track a reference joint velocity: ``e = v_ref dt``, ``J = I[root_nv:, :]``
    (``joint_velocity_task.py:59-110``), unit gain, no LM damping."""

    def __init__(self, cost: float) -> None:
        super().__init__(cost=cost, gain=1.0, lm_damping=0.0)
        self._target_Delta_q: Optional[object] = None

    def set_target(self, target_v, dt: float) -> None:
        """``target_v``: ``[nv - root_nv]`` for all instances or ``[B, nv - root_nv]``."""
        if np.ndim(target_v) not in (1, 2):
            raise TaskDefinitionError(
                f"joint velocity target should be a vector, but the provided target has shape {np.shape(target_v)}"
            )
        self._target_Delta_q = as_vector_target(target_v, np.shape(target_v)[-1]) * dt

    def _error_target(self, model):
        if self._target_Delta_q is None:
            raise TargetNotSet(repr(self))
        return self._target_Delta_q

    def _pk_describe(self, model) -> dict:
        _, root_nv = get_root_joint_dim(model)
        task_nv = model.nv - root_nv
        target = self._error_target(model)
        if np.shape(target)[-1] != task_nv:
            raise TaskDefinitionError(
                f"Target has dimension nv={np.shape(target)[-1]} but the task expects nv={task_nv} "
                f"({model.nv=}, {root_nv=})"
            )
        cost6 = np.zeros(6)
        cost6[0] = float(self.cost)
        return {"type": PK_TASK_JOINT_VELOCITY, "frame": 0, "root": 0, "cost6": cost6, "k": task_nv, "target": target}

    def __repr__(self):
        return f"JointVelocityTask(cost={self.cost})"


class DampingTask(JointVelocityTask):
    r"""Minimise joint velocities: a joint-velocity task with zero error
    (``damping_task.py:15-47``)."""

    def __init__(self, cost: float) -> None:
        super().__init__(cost=cost)

    def _error_target(self, model):
        _, root_nv = get_root_joint_dim(model)
        return np.zeros(model.nv - root_nv)

    def __repr__(self):
        return f"DampingTask(cost={self.cost})"
