"""Module path of the reference (``pink/tasks/damping_task.py``); the class lives with
its parent in :mod:`pink_b200.tasks.joint_velocity_task`."""

from .joint_velocity_task import DampingTask

__all__ = ["DampingTask"]
