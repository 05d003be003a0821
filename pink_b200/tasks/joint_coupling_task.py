"""Module path of the reference (``pink/tasks/joint_coupling_task.py``); the class lives
with its parent in :mod:`pink_b200.tasks.linear_holonomic_task`."""

from .linear_holonomic_task import JointCouplingTask

__all__ = ["JointCouplingTask"]
