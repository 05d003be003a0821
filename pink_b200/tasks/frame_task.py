"""Frame task (``/root/reference/pink/tasks/frame_task.py``)."""

from typing import Optional, Sequence, Union

import numpy as np

from .._cabi import PK_TASK_FRAME
from ..exceptions import FrameNotFound, TargetNotSet, TaskDefinitionError
from ..spatial import SE3
from ._targets import as_se3_target
from .task import Task


class FrameTask(Task):
    r"""Regulate the pose of a robot frame in the world frame.

    ``e = log6(T_b^-1 T_t)`` (``frame_task.py:181-193``),
    ``J = -Jlog6(T_t^-1 T_b) bJ_b`` (``frame_task.py:222-227``).

    Attributes:
        frame: Frame name.
        transform_target_to_world: target pose: an :class:`SE3` shared by all
            instances, or a ``[B, 12]`` tensor (rows ``[R | p]``) per instance.
    """

    frame: str
    transform_target_to_world: Optional[object]

    def __init__(
        self,
        frame: str,
        position_cost: Union[float, Sequence[float]],
        orientation_cost: Union[float, Sequence[float]],
        lm_damping: float = 0.0,
        gain: float = 1.0,
    ) -> None:
        super().__init__(cost=np.ones(6), gain=gain, lm_damping=lm_damping)
        self.frame = frame
        self.lm_damping = lm_damping
        self.transform_target_to_world = None
        self.set_position_cost(position_cost)
        self.set_orientation_cost(orientation_cost)

    def set_position_cost(self, position_cost) -> None:
        if isinstance(position_cost, float):
            assert position_cost >= 0.0
        else:
            assert all(cost >= 0.0 for cost in position_cost)
        if isinstance(self.cost, np.ndarray):
            self.cost[0:3] = position_cost
        else:
            raise TaskDefinitionError(f"Frame task cost should be a vector, currently cost={self.cost}")

    def set_orientation_cost(self, orientation_cost) -> None:
        if isinstance(orientation_cost, float):
            assert orientation_cost >= 0.0
        else:
            assert all(cost >= 0.0 for cost in orientation_cost)
        if isinstance(self.cost, np.ndarray):
            self.cost[3:6] = orientation_cost
        else:
            raise TaskDefinitionError(f"Frame task cost should be a vector, currently cost={self.cost}")

    def set_target(self, transform_target_to_world) -> None:
        """Set the target pose (copied, ``frame_task.py:129-136``): an SE3 for
        all instances or ``[B, 3, 4]`` / ``[B, 4, 4]`` / ``[B, 12]`` per instance."""
        self.transform_target_to_world = as_se3_target(transform_target_to_world)

    def set_target_from_configuration(self, configuration) -> None:
        self.set_target(configuration.get_transform_frame_to_world(self.frame))

    def _pk_describe(self, model) -> dict:
        if self.transform_target_to_world is None:
            raise TargetNotSet(f"no target set for frame '{self.frame}'")
        if not model.existFrame(self.frame):
            raise FrameNotFound(self.frame, model.frames)
        if not isinstance(self.cost, np.ndarray):
            raise TaskDefinitionError(f"Frame task cost should be a vector, currently cost={self.cost}")
        tgt = self.transform_target_to_world
        return {
            "type": PK_TASK_FRAME,
            "frame": model.getFrameId(self.frame),
            "root": 0,
            "cost6": np.asarray(self.cost, dtype=np.float64),
            "k": 6,
            "target": tgt.as_3x4().reshape(12) if isinstance(tgt, SE3) else tgt,
        }

    @property
    def position_cost(self):
        if isinstance(self.cost, np.ndarray):
            return self.cost[0:3]
        elif isinstance(self.cost, float):
            return self.cost
        raise TaskDefinitionError(f"Frame task cost should be a vector or a scalar, currently cost={self.cost}")

    @property
    def orientation_cost(self):
        if isinstance(self.cost, np.ndarray):
            return self.cost[3:6]
        elif isinstance(self.cost, float):
            return self.cost
        raise TaskDefinitionError(f"Frame task cost should be a vector or a scalar, currently cost={self.cost}")

    def __repr__(self):
        return (
            "FrameTask("
            f"frame={self.frame}, "
            f"position_cost={self.position_cost}, "
            f"orientation_cost={self.orientation_cost}, "
            f"lm_damping={self.lm_damping}, "
            f"gain={self.gain})"
        )
