"""Low-acceleration task
(``/root/reference/pink/tasks/low_acceleration_task.py``)."""

from typing import Optional

import numpy as np

from .._cabi import PK_TASK_JOINT_VELOCITY
from ..utils import get_root_joint_dim
from ._targets import as_vector_target
from .task import Task


class LowAccelerationTask(Task):
    r"""Minimise :math:`\|v - v_{prev}\|^2`: the error is :math:`-\mathrm{d}t\,v_{prev}`
    with the posture Jacobian ``I[root_nv:, :]``, unit gain and no LM damping
    (``low_acceleration_task.py:58-80``).  On the device this is the
    joint-velocity task with ``e = -dq_prev``."""

    def __init__(self, cost: float) -> None:
        super().__init__(cost=cost, gain=1.0, lm_damping=0.0)
        self.Delta_q_prev: Optional[object] = None

    def set_last_integration(self, v_prev, dt) -> None:
        """``v_prev``: ``[nv]`` for all instances or ``[B, nv]`` per instance."""
        self.Delta_q_prev = as_vector_target(v_prev, np.shape(v_prev)[-1]) * dt

    def _pk_describe(self, model) -> dict:
        _, root_nv = get_root_joint_dim(model)
        prev = np.zeros(model.nv) if self.Delta_q_prev is None else self.Delta_q_prev
        cost6 = np.zeros(6)
        cost6[0] = float(self.cost)
        return {"type": PK_TASK_JOINT_VELOCITY, "frame": 0, "root": 0, "cost6": cost6, "k": model.nv - root_nv,
                "target": -prev[..., root_nv:]}

    def __repr__(self):
        return f"LowAccelerationTask(cost={self.cost})"
