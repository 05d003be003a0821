"""Center-of-mass task (``/root/reference/pink/tasks/com_task.py``)."""

from typing import Optional, Sequence, Union

import numpy as np

from .._cabi import PK_TASK_COM
from ..exceptions import TargetNotSet, TaskDefinitionError
from ._targets import as_vector_target
from .task import Task


class ComTask(Task):
    r"""Regulate the position of the center of mass in the world frame.

    ``e = com(q) - com*``, ``J = jacobianCenterOfMass`` (``com_task.py:120-148``).
    """

    target_com: Optional[object]

    def __init__(self, cost: Union[float, Sequence[float]], lm_damping: float = 0.0, gain: float = 1.0) -> None:
        super().__init__(cost=np.ones(3), gain=gain, lm_damping=lm_damping)
        self.target_com = None
        self.set_cost(cost)

    def set_cost(self, cost) -> None:
        if isinstance(cost, float):
            assert cost >= 0.0
        else:
            assert all(c >= 0.0 for c in cost)
        if isinstance(self.cost, np.ndarray):
            self.cost[0:3] = cost
        else:
            raise TaskDefinitionError(f"CoM task cost should be a vector, currently cost={self.cost}")

    def set_target(self, target_com) -> None:
        """``[3]`` for all instances or ``[B, 3]`` per instance (copied)."""
        self.target_com = as_vector_target(target_com, 3)

    def set_target_from_configuration(self, configuration) -> None:
        self.set_target(configuration.get_center_of_mass())

    def _pk_describe(self, model) -> dict:
        if self.target_com is None:
            raise TargetNotSet("no target set for CoM")
        cost6 = np.zeros(6)
        cost6[0:3] = self.cost[0:3]
        return {
            "type": PK_TASK_COM,
            "frame": 0,
            "root": 0,
            "cost6": cost6,
            "k": 3,
            "target": self.target_com,
        }

    def __repr__(self):
        cost = self.cost if isinstance(self.cost, float) else self.cost[0:3]
        return (
            "ComTask("
            f"target_com={self.target_com}, "
            f"cost={cost}, "
            f"gain={self.gain}, "
            f"lm_damping={self.lm_damping})"
        )
