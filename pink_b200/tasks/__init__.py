"""Kinematic tasks (``/root/reference/pink/tasks/__init__.py``).

On the hot path (BASELINE north_star): :class:`FrameTask`,
:class:`PostureTask`, :class:`ComTask`, :class:`RelativeFrameTask`; from the
"next" rows: :class:`JointVelocityTask`, :class:`DampingTask`,
:class:`LowAccelerationTask`, :class:`LinearHolonomicTask`,
:class:`JointCouplingTask`.
"""

from .com_task import ComTask
from .frame_task import FrameTask
from .joint_velocity_task import DampingTask, JointVelocityTask
from .linear_holonomic_task import JointCouplingTask, LinearHolonomicTask
from .low_acceleration_task import LowAccelerationTask
from .posture_task import PostureTask
from .relative_frame_task import RelativeFrameTask
from .task import Task

__all__ = [cls.__name__ for cls in (
    ComTask, DampingTask, FrameTask, JointCouplingTask, JointVelocityTask,
    LinearHolonomicTask, LowAccelerationTask, PostureTask, RelativeFrameTask, Task,
)]
