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
import unittest

from pink_b200 import model as _host
from pink_b200.model import JointModelFreeFlyer  # noqa: F401
from pink_b200.model import SE3, load_urdf, model_from_urdf_string  # noqa: F401

# Pinocchio marks "no limit" with the largest double, not with inf (the reference's tests rely on
# 0 * limit == 0); the models handed to the reference's code are converted accordingly.
HUGE = float(np.finfo(np.float64).max)


def pinocchio_like_limits(model):
    for name in ("lowerPositionLimit", "upperPositionLimit", "velocityLimit"):
        v = np.array(getattr(model, name), dtype=np.float64)
        v[np.isposinf(v)] = HUGE
        v[np.isneginf(v)] = -HUGE
        setattr(model, name, v)
    return model


class FrameType:
    OP_FRAME, JOINT, FIXED_JOINT, BODY, SENSOR = "OP_FRAME", "JOINT", "FIXED_JOINT", "BODY", "SENSOR"


def Frame(name, parent_joint, *rest):
    """``pin.Frame(name, parentJoint, [parentFrame,] placement, type)``."""
    placement, frame_type = rest[-2], rest[-1]
    return _host.Frame(name, parent_joint, placement, frame_type)


class _Unsupported:
    """Joint models outside this repo's scope (1-dof joints + free-flyer)."""


class JointModelPlanar(_Unsupported):
    pass


class JointModelSpherical(_Unsupported):
    pass


class JointModelRevoluteUnaligned:
    def __init__(self, *axis):
        self.axis = np.array(axis if len(axis) == 3 else (1.0, 0.0, 0.0), dtype=np.float64)


class Model(_host.Model):
    """``pin.Model()`` as the reference's tests build it by hand."""

    def addJoint(self, parent, joint_model, placement, name, max_effort=None, max_velocity=None, min_config=None,
                 max_config=None):
        if isinstance(joint_model, _Unsupported):
            raise unittest.SkipTest(f"{type(joint_model).__name__} is outside the scope of this repo")
        lo = -np.inf if min_config is None else float(np.asarray(min_config).reshape(-1)[0])
        hi = np.inf if max_config is None else float(np.asarray(max_config).reshape(-1)[0])
        vel = np.inf if max_velocity is None else float(np.asarray(max_velocity).reshape(-1)[0])
        jid = self.add_joint(name, parent, placement, joint_model.axis, "revolute", lo, hi, vel)
        pinocchio_like_limits(self)
        return jid


class Data(_host.Data):
    """``pin.Data(model)``: ``J`` exists (zeros) before the first ``computeJointJacobians``."""

    def __init__(self, model=None):
        super().__init__(model)
        if model is not None:
            self.J = np.zeros((6, model.nv))


class RobotWrapper(_host.RobotWrapper):
    def __init__(self, model=None, **unused):
        super().__init__(model)
        self.data = Data(model)

    @staticmethod
    def BuildFromURDF(*a, **k):
        raise unittest.SkipTest("URDF packages with meshes / collision geometry are not available offline")

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


class CollisionPair:
    def __init__(self, first, second):
        self.first, self.second = int(first), int(second)


class GeometryObject:
    def __init__(self, name, parent_joint, radius):
        self.name, self.parentJoint, self.radius = name, int(parent_joint), float(radius)


class GeometryModel:
    """Spheres only (SURVEY section 2 scopes the self-collision barrier to sphere pairs; the
    reference evaluates arbitrary meshes through hpp-fcl / coal): built from a
    ``pink_b200.collision.SphereCollisionModel`` so that the reference's ``SelfCollisionBarrier``
    (``pink/barriers/self_collision_barrier.py:85-224``) can be executed on the same spheres."""

    def __init__(self, spheres=None):
        self.geometryObjects, self.collisionPairs, self._sphere_frames = [], [], []
        if spheres is not None:
            # radii as the C ABI carries them (fp32): the scenarios' oracle records hold these values
            self.geometryObjects = [GeometryObject(n, j, float(np.float32(r))) for n, j, r in zip(spheres.names, spheres.parents, spheres.radii)]
            self.collisionPairs = [CollisionPair(i, j) for i, j in spheres.collisionPairs]
            self._sphere_frames = [spheres.model.frames[f].name for f in spheres.frames]


class DistanceResult:
    def __init__(self, min_distance, w1, w2):
        self.min_distance, self._w1, self._w2 = float(min_distance), np.array(w1), np.array(w2)

    def getNearestPoint1(self):
        return self._w1

    def getNearestPoint2(self):
        return self._w2


class GeometryData:
    def __init__(self, collision_model):
        self.collision_model = collision_model
        self.distanceResults = []


def computeCollisions(model, data, collision_model, collision_data, q, stop_at_first):
    """``pink/configuration.py:147-154``: nothing to do for analytic sphere pairs."""


def updateGeometryPlacements(*unused):
    pass


def computeDistances(model, data, collision_model, collision_data, q):
    """``pink/configuration.py:155-161``: signed distance and nearest points of every pair."""
    from oracle import barriers as _bar

    t = _table(model)
    fk = _kin.forward_kinematics(t, np.asarray(q, dtype=np.float64))
    pairs = []
    for cp in collision_model.collisionPairs:
        fa = t.frame_names.index(collision_model._sphere_frames[cp.first])
        fb = t.frame_names.index(collision_model._sphere_frames[cp.second])
        pairs.append((fa, fb, collision_model.geometryObjects[cp.first].radius, collision_model.geometryObjects[cp.second].radius))
    collision_data.distanceResults = [DistanceResult(d, w1, w2) for d, w1, w2 in _bar.sphere_pair_distances(t, fk, pairs)]


def getJointJacobian(model, data, joint_id, reference_frame):
    """``pink/barriers/self_collision_barrier.py:205-214`` (LOCAL_WORLD_ALIGNED): the LOCAL
    Jacobian of the joint's own frame, rotated into world axes."""
    if reference_frame != ReferenceFrame.LOCAL_WORLD_ALIGNED:
        raise NotImplementedError("the in-scope reference code asks for LOCAL_WORLD_ALIGNED joint Jacobians only")
    J = getFrameJacobian(model, data, model.getFrameId(model.joints[joint_id].name), ReferenceFrame.LOCAL)
    R = data.oMi[joint_id].rotation
    return np.vstack([R @ J[:3], R @ J[3:]])


def _table(model):
    key = (getattr(model, "_version", 0), len(model.frames), len(model.joints),
           tuple(np.asarray(model.upperPositionLimit).tolist()), tuple(np.asarray(model.velocityLimit).tolist()))
    cached = model.__dict__.get("_refshim_table")
    if cached is None or cached[0] != key:
        cached = (key, model.table())
        model.__dict__["_refshim_table"] = cached
    return cached[1]


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
    # data.J: joint Jacobians as spatial velocities in the world frame (columns oMi.act(S))
    J = np.zeros((6, t.nv))
    rv = 6 if model.free_flyer else 0
    if model.free_flyer:
        J[:, :6] = _lie.action(R_root, p_root)
    for j in range(t.njoints):
        a = R[j] @ np.asarray(t.axis[j], dtype=np.float64)
        if int(t.jtype[j]) == 0:
            J[:3, rv + j], J[3:, rv + j] = np.cross(p[j], a), a
        else:
            J[:3, rv + j] = a
    data.J = J


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


def centerOfMass(model, data, q=None, *unused):
    """``pink/tasks/com_task.py:103-105,123-125``."""
    t = _table(model)
    q = data._q if q is None or isinstance(q, bool) else np.asarray(q, dtype=np.float64)
    return np.array(_kin.center_of_mass(t, _kin.forward_kinematics(t, q)))


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
    if isinstance(root_joint, _Unsupported):
        raise unittest.SkipTest(f"{type(root_joint).__name__} is outside the scope of this repo")
    return pinocchio_like_limits(model_from_urdf_string(xml, root_joint=root_joint))
