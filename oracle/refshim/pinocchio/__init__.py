"""Stand-in for the ``pinocchio`` module NAME so that the reference's own Python can be
imported here (see ../README.md).  TEST INFRASTRUCTURE: not Pinocchio, not product code.

Every function below is the oracle's restatement (``oracle/kinematics.py``,
``oracle/lie.py``) behind the signature the reference calls; the docstrings cite the
reference call sites.  Model / Data / SE3 are the host classes of ``pink_b200.model``
(pure numpy, fp64), which already spell their attributes like ``pin.Model``.
"""
import enum

import numpy as np

from oracle import kinematics as _kin
from oracle import lie as _lie
from pink_b200.model import Data, Frame, JointModelFreeFlyer, Model, RobotWrapper  # noqa: F401
from pink_b200.model import SE3, load_urdf, model_from_urdf_string  # noqa: F401

__version__ = "0.0.0+oracle.refshim"


class ReferenceFrame(enum.Enum):
    WORLD = 0
    LOCAL = 1
    LOCAL_WORLD_ALIGNED = 2


class ArgumentPosition(enum.Enum):
    ARG0 = 0
    ARG1 = 1


ARG0 = ArgumentPosition.ARG0
ARG1 = ArgumentPosition.ARG1


class GeometryModel:  # collision geometry is out of reach of this shim (hpp-fcl / coal)
    def __init__(self, *a, **k):
        raise NotImplementedError("collision geometry is not part of the reference shim")


class GeometryData(GeometryModel):
    pass


def _table(model):
    t = model.__dict__.get("_refshim_table")
    if t is None:
        t = model.table()
        model.__dict__["_refshim_table"] = t
    return t


def _frame_index(model, frame_id):
    t = _table(model)
    return t.frame_names.index(model.frames[frame_id].name)


def _se3(R, p):
    return SE3(np.array(R, dtype=np.float64), np.array(p, dtype=np.float64))


def forwardKinematics(model, data, q):
    t = _table(model)
    data._q = np.array(q, dtype=np.float64)
    data._fk = _kin.forward_kinematics(t, data._q)
    R_root, p_root, R, p = data._fk
    first = 2 if model.free_flyer else 1
    oMi = [SE3.Identity()]
    if model.free_flyer:
        oMi.append(_se3(R_root, p_root))
    for j in range(t.njoints):
        oMi.append(_se3(R[j], p[j]))
    assert len(oMi) == first + t.njoints
    data.oMi = oMi


def computeJointJacobians(model, data, q):
    """``pink/configuration.py:163``: FK + the full model Jacobian (kept implicit here; the
    frame Jacobians are evaluated from the stored FK on demand)."""
    forwardKinematics(model, data, q)


def updateFramePlacements(model, data):
    """``pink/configuration.py:164``: ``data.oMf`` for every frame of the model."""
    t = _table(model)
    oMf = []
    for fr in model.frames:
        if fr.name in t.frame_names:
            oMf.append(_se3(*_kin.frame_placement(t, data._fk, t.frame_names.index(fr.name))))
        else:  # the universe frame
            oMf.append(SE3.Identity())
    data.oMf = oMf


def getFrameJacobian(model, data, frame_id, reference_frame):
    """``pink/configuration.py:233-235``, ``limits/floating_base_velocity_limit.py:128-133``."""
    if reference_frame != ReferenceFrame.LOCAL:
        raise NotImplementedError("the in-scope reference code asks for LOCAL Jacobians only")
    return np.array(_kin.frame_jacobian_local(_table(model), data._fk, _frame_index(model, frame_id)))


def integrate(model, q, dv):
    """``pink/configuration.py:283,292``."""
    return np.array(_kin.integrate(_table(model), np.asarray(q, dtype=np.float64), np.asarray(dv, dtype=np.float64)))


def difference(model, q0, q1):
    """``pink/tasks/posture_task.py:103``, ``limits/configuration_limit.py:111-116``."""
    return np.array(_kin.difference(_table(model), np.asarray(q0, dtype=np.float64), np.asarray(q1, dtype=np.float64)))


def dDifference(model, q0, q1, arg):
    """``pink/tasks/linear_holonomic_task.py:190``."""
    if arg != ARG1:
        raise NotImplementedError("the reference asks for ARG1 only")
    return np.array(_kin.d_difference_arg1(_table(model), np.asarray(q0, dtype=np.float64), np.asarray(q1, dtype=np.float64)))


def neutral(model):
    return np.array(_kin.neutral(_table(model)))


def centerOfMass(model, data, q, *unused):
    """``pink/tasks/com_task.py:103-105,123-125``."""
    t = _table(model)
    return np.array(_kin.center_of_mass(t, _kin.forward_kinematics(t, np.asarray(q, dtype=np.float64))))


def jacobianCenterOfMass(model, data, q, *unused):
    """``pink/tasks/com_task.py:145-147``."""
    t = _table(model)
    return np.array(_kin.com_jacobian(t, _kin.forward_kinematics(t, np.asarray(q, dtype=np.float64))))


class _Motion:
    def __init__(self, vector):
        self.vector = np.array(vector, dtype=np.float64)
        self.linear = self.vector[:3]
        self.angular = self.vector[3:]

    @property
    def np(self):
        return self.vector


def log(M):
    """``pin.log(SE3).vector`` (``pink/tasks/frame_task.py:192``)."""
    return _Motion(_lie.log6(M.rotation, M.translation))


log6 = log


def Jlog6(M):
    """``pink/tasks/frame_task.py:226``, ``relative_frame_task.py:242``."""
    return np.array(_lie.jlog6(M.rotation, M.translation))


def skew(v):
    return np.array(_lie.hat(np.asarray(v, dtype=np.float64)))


def buildModelFromXML(xml, root_joint=None):
    return model_from_urdf_string(xml, root_joint=root_joint)
