"""Relative frame task (``/root/reference/pink/tasks/relative_frame_task.py``)."""

from typing import Optional, Sequence, Union

import numpy as np

from .._cabi import PK_TASK_RELATIVE_FRAME
from ..exceptions import FrameNotFound, TargetNotSet, TaskDefinitionError
from ..spatial import SE3
from ._targets import as_se3_target
from .task import Task


class RelativeFrameTask(Task):
    r"""Regulate the pose of a frame relative to another (moving) frame.

    ``e = log6(T_rt^-1 T_rf)`` (``relative_frame_task.py:178-185``),
    ``J = Jlog6(T_tf) (fJ_f - Ad_{T_rf^-1} rJ_r)`` (``:233-246``).
    """

    frame: str
    root: str
    transform_target_to_root: Optional[object]

    def __init__(
        self,
        frame: str,
        root: str,
        position_cost: Union[float, Sequence[float]],
        orientation_cost: Union[float, Sequence[float]],
        lm_damping: float = 0.0,
        gain: float = 1.0,
    ) -> None:
        super().__init__(cost=np.ones(6), gain=gain, lm_damping=lm_damping)
        self.frame = frame
        self.root = root
        self.lm_damping = lm_damping
        self.transform_target_to_root = None
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

    def set_target(self, transform_target_to_root) -> None:
        self.transform_target_to_root = as_se3_target(transform_target_to_root)

    def set_target_from_configuration(self, configuration) -> None:
        self.set_target(configuration.get_transform(self.frame, self.root))

    def _pk_describe(self, model) -> dict:
        if self.transform_target_to_root is None:
            raise TargetNotSet(
                f"target pose of frame '{self.frame}' in frame '{self.root}' is undefined"
            )
        for name in (self.frame, self.root):
            if not model.existFrame(name):
                raise FrameNotFound(name, model.frames)
        tgt = self.transform_target_to_root
        return {
            "type": PK_TASK_RELATIVE_FRAME,
            "frame": model.getFrameId(self.frame),
            "root": model.getFrameId(self.root),
            "cost6": np.asarray(self.cost, dtype=np.float64),
            "k": 6,
            "target": tgt.as_3x4().reshape(12) if isinstance(tgt, SE3) else tgt,
        }

    @property
    def position_cost(self):
        return self.cost[0:3] if isinstance(self.cost, np.ndarray) else self.cost

    @property
    def orientation_cost(self):
        return self.cost[3:6] if isinstance(self.cost, np.ndarray) else self.cost

    def __repr__(self):
        return (
            "RelativeFrameTask("
            f"frame={self.frame}, "
            f"root={self.root}, "
            f"position_cost={self.position_cost}, "
            f"orientation_cost={self.orientation_cost}, "
            f"lm_damping={self.lm_damping}, "
            f"gain={self.gain})"
        )
