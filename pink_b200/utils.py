"""Utility functions with the reference's names (``/root/reference/pink/utils.py``)."""

from typing import Tuple

import numpy as np

from .exceptions import ConfigurationError, PinkError
from .model import neutral


def custom_configuration_vector(robot, **kwargs) -> np.ndarray:
    """Configuration where named joints take given values, others neutral
    (``pink/utils.py:16-37``)."""
    model = robot.model if hasattr(robot, "model") else robot
    q = neutral(model)
    for name, value in kwargs.items():
        joint_id = model.getJointId(name)
        if joint_id >= len(model.joints):
            raise PinkError(f"joint '{name}' not found in the model")
        joint = model.joints[joint_id]
        value = np.array(value, dtype=float).flatten()
        if value.shape[0] != joint.nq:
            raise ConfigurationError(
                f"Joint '{name}' has {joint.nq=} but is set to {value.shape=}"
            )
        q[joint.idx_q : joint.idx_q + joint.nq] = value
    return q


def get_root_joint_dim(model) -> Tuple[int, int]:
    """``(nq, nv)`` of the joint named ``root_joint``, else ``(0, 0)``
    (``pink/utils.py:40-54``)."""
    if model.existJointName("root_joint"):
        root_joint = model.joints[model.getJointId("root_joint")]
        return root_joint.nq, root_joint.nv
    return 0, 0


def get_joint_idx(model, joint_name: str) -> Tuple[int, int]:
    """``(idx_q, idx_v)`` of a joint (``pink/utils.py:57-74``)."""
    if model.existJointName(joint_name):
        joint = model.joints[model.getJointId(joint_name)]
        return joint.idx_q, joint.idx_v
    raise PinkError(f"cannot find the joint index corresponding to joint {joint_name}")


class VectorSpace:
    """Read-only ``eye / ones / zeros`` of a vector space (``pink/utils.py:77-113``)."""

    def __init__(self, dim: int):
        eye, ones, zeros = np.eye(dim), np.ones(dim), np.zeros(dim)
        for a in (eye, ones, zeros):
            a.setflags(write=False)
        self.__eye, self.__ones, self.__zeros = eye, ones, zeros

    @property
    def eye(self) -> np.ndarray:
        return self.__eye

    @property
    def ones(self) -> np.ndarray:
        return self.__ones

    @property
    def zeros(self) -> np.ndarray:
        return self.__zeros


def process_collision_pairs(model, collision_model, srdf_path: str = ""):
    """Add every collision pair of ``collision_model`` (a
    :class:`pink_b200.SphereCollisionModel`) and drop the ones an SRDF file disables
    (``pink/utils.py:116-142``).  Returns the collision data to hand to
    :class:`pink_b200.Configuration` (a :class:`pink_b200.collision.SphereCollisionData`;
    the solver itself evaluates the distances inside the kernels)."""
    from .collision import SphereCollisionData

    collision_model.add_all_collision_pairs()
    if srdf_path != "":
        collision_model.remove_collision_pairs_from_srdf(srdf_path)
    return SphereCollisionData(collision_model)
