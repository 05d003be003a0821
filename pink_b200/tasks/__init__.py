"""Kinematic tasks (``/root/reference/pink/tasks/__init__.py``).

On the hot path (BASELINE north_star): :class:`FrameTask`,
:class:`PostureTask`, :class:`ComTask`, :class:`RelativeFrameTask`.  The
constant-Jacobian tasks (``JointCouplingTask``, ``DampingTask``, ...) are
SURVEY section 8(f) "next" rows and are not provided yet.
"""

from .com_task import ComTask
from .frame_task import FrameTask
from .posture_task import PostureTask
from .relative_frame_task import RelativeFrameTask
from .task import Task

__all__ = ["ComTask", "FrameTask", "PostureTask", "RelativeFrameTask", "Task"]
