"""Kinematic model container and URDF loader (host side, set-up time only).

The reference reads its model through Pinocchio's ``pin.Model`` /
``pin.Data``; the attributes and methods Pink touches on them are listed in
SURVEY.md section 8(b) (``model.nq/nv/joints[i].idx_q/idx_v/nq/nv``,
``upperPositionLimit/lowerPositionLimit/velocityLimit``,
``hasConfigurationLimit()``, ``existFrame/getFrameId/frames``,
``existJointName/getJointId`` with the magic name ``"root_joint"``;
``/root/reference/pink/utils.py:40-54``,
``/root/reference/pink/limits/configuration_limit.py:50-72``).  :class:`Model`
exposes the same names so that the Python layer reads like the reference's,
and :meth:`Model.table` flattens it into the plain arrays that both the C-ABI
(``PkModelDesc`` in ``include/pink_b200.h``) and the test oracle consume.

Supported: trees of 1-dof revolute / prismatic joints (URDF ``revolute``,
``continuous``, ``prismatic``), ``fixed`` joints (merged into the parent body,
kept as frames), optional free-flyer root named ``root_joint``.
Deviation from Pinocchio, stated once: a URDF ``continuous`` joint is kept as
a 1-coordinate unbounded revolute joint instead of Pinocchio's (cos, sin) pair.
"""

from __future__ import annotations

import types
import xml.etree.ElementTree as ET
from typing import List, Optional, Sequence

import numpy as np

from .spatial import SE3, rpy_to_matrix

REVOLUTE, PRISMATIC = 0, 1
_INF = np.inf


class JointModelFreeFlyer:
    """Marker with the same spelling as ``pin.JointModelFreeFlyer()``."""


class JointModel:
    """Index bookkeeping of one joint (``pin.JointModel`` subset)."""

    def __init__(self, jid, name, kind, idx_q, idx_v, nq, nv):
        self.id = jid
        self.name = name
        self.kind = kind  # "universe" | "free_flyer" | "revolute" | "prismatic"
        self.idx_q = idx_q
        self.idx_v = idx_v
        self.nq = nq
        self.nv = nv

    def shortname(self):
        return {
            "universe": "JointModelUniverse",
            "free_flyer": "JointModelFreeFlyer",
            "revolute": "JointModelRevoluteUnaligned",
            "prismatic": "JointModelPrismaticUnaligned",
        }[self.kind]

    def __repr__(self):
        return f"JointModel({self.name}, {self.kind}, idx_q={self.idx_q}, idx_v={self.idx_v})"


class Frame:
    """Operational frame attached to a joint's body (``pin.Frame`` subset)."""

    def __init__(self, name, parent_joint, placement, frame_type="OP_FRAME"):
        self.name = name
        self.parentJoint = parent_joint  # Pinocchio joint id (0 = universe)
        self.parent = parent_joint
        self.placement = placement
        self.type = frame_type

    def __repr__(self):
        return f"Frame({self.name}, parent={self.parentJoint})"


class Data:
    """Per-configuration results (``pin.Data`` subset: ``oMi``, ``oMf``, ``J``).

    Filled by :class:`pink_b200.Configuration` from the GPU for single
    configurations; carries nothing for batches (the batched results live in
    device tensors on the configuration)."""

    def __init__(self, model=None):
        self.oMi: List[SE3] = []
        self.oMf: List[SE3] = []
        self.J = None

    def copy(self):
        d = Data()
        d.oMi = [T.copy() for T in self.oMi]
        d.oMf = [T.copy() for T in self.oMf]
        d.J = None if self.J is None else self.J.copy()
        return d


class Model:
    """Joint tree + frames + limits + inertias."""

    def __init__(self, name: str = "model", free_flyer: bool = False):
        self.name = name
        self.free_flyer = bool(free_flyer)
        self.names: List[str] = ["universe"]
        self.joints: List[JointModel] = [JointModel(0, "universe", "universe", -1, -1, 0, 0)]
        self.parents: List[int] = [0]  # Pinocchio-style parent joint ids
        self.jointPlacements: List[SE3] = [SE3()]
        self.axes: List[np.ndarray] = [np.zeros(3)]
        self.frames: List[Frame] = [Frame("universe", 0, SE3(), "FIXED_JOINT")]
        self.masses: List[float] = [0.0]  # per Pinocchio joint id (body of that joint)
        self.coms: List[np.ndarray] = [np.zeros(3)]
        self.nq = 0
        self.nv = 0
        self._q_min: List[float] = []
        self._q_max: List[float] = []
        self._v_max: List[float] = []
        self._has_cfg_limit: List[bool] = []
        self._version = 0
        if self.free_flyer:
            self._add_joint_record("root_joint", "free_flyer", 0, SE3(), np.zeros(3), 7, 6)
            self._q_min += [-_INF] * 7
            self._q_max += [_INF] * 7
            self._v_max += [_INF] * 6
            self._has_cfg_limit += [True, True, True, False, False, False, False]
            self.frames.append(Frame("root_joint", 1, SE3(), "JOINT"))

    # -- construction ------------------------------------------------------
    def _add_joint_record(self, name, kind, parent, placement, axis, nq, nv):
        jid = len(self.joints)
        self.joints.append(JointModel(jid, name, kind, self.nq, self.nv, nq, nv))
        self.names.append(name)
        self.parents.append(parent)
        self.jointPlacements.append(placement.copy())
        self.axes.append(np.asarray(axis, dtype=np.float64))
        self.masses.append(0.0)
        self.coms.append(np.zeros(3))
        self.nq += nq
        self.nv += nv
        self._version += 1
        return jid

    @property
    def root_body(self) -> int:
        """Pinocchio joint id of the body fixed joints of the root attach to."""
        return 1 if self.free_flyer else 0

    def add_joint(
        self,
        name: str,
        parent: int,
        placement: SE3,
        axis: Sequence[float],
        kind: str = "revolute",
        lower: float = -_INF,
        upper: float = _INF,
        velocity: float = _INF,
    ) -> int:
        """Append a 1-dof joint under Pinocchio joint id ``parent``; returns its id."""
        if kind not in ("revolute", "prismatic"):
            raise ValueError(f"unsupported joint kind {kind!r}")
        if name in self.names:
            raise ValueError(f"duplicate joint name {name!r}")
        if not 0 <= parent < len(self.joints):
            raise ValueError("parent joint must already exist")
        axis = np.asarray(axis, dtype=np.float64)
        nrm = np.linalg.norm(axis)
        if nrm == 0.0:
            raise ValueError("zero joint axis")
        jid = self._add_joint_record(name, kind, parent, placement, axis / nrm, 1, 1)
        self._q_min.append(float(lower))
        self._q_max.append(float(upper))
        self._v_max.append(float(velocity))
        self._has_cfg_limit.append(True)
        self.frames.append(Frame(name, jid, SE3(), "JOINT"))
        return jid

    def add_frame(self, name: str, parent_joint: int, placement: SE3, frame_type="OP_FRAME") -> int:
        self.frames.append(Frame(name, parent_joint, placement.copy(), frame_type))
        self._version += 1
        return len(self.frames) - 1

    def addFrame(self, frame: Frame) -> int:  # Pinocchio spelling
        return self.add_frame(frame.name, frame.parentJoint, frame.placement, frame.type)

    def append_inertia(self, joint_id: int, mass: float, com_in_joint: Sequence[float]) -> None:
        """Add a point-mass-equivalent (mass, CoM) to the body of ``joint_id``."""
        mass = float(mass)
        if mass <= 0.0:
            return
        m0 = self.masses[joint_id]
        c0 = self.coms[joint_id]
        c1 = np.asarray(com_in_joint, dtype=np.float64)
        self.masses[joint_id] = m0 + mass
        self.coms[joint_id] = (m0 * c0 + mass * c1) / (m0 + mass)
        self._version += 1

    # -- flat image (what travels between ranks) -----------------------------
    def pack(self):
        """The model constants as three flat arrays: ``(floats, ints, text)`` - float64 numbers
        (placements, axes, limits, inertias), int64 structure (parents, kinds, sizes) and a UTF-8
        blob of the names.  :meth:`unpack` rebuilds an equal model; this is the payload of
        ``parallel.broadcast_model`` (no pickle: only numbers and names cross the wire)."""
        import json

        kinds = {"revolute": 0, "prismatic": 1}
        first = 2 if self.free_flyer else 1
        jf, ji = [], []
        for j in range(first, len(self.joints)):
            jt = self.joints[j]
            P = self.jointPlacements[j]
            jf += list(P.rotation.reshape(9)) + list(P.translation) + list(self.axes[j])
            jf += [self._q_min[jt.idx_q], self._q_max[jt.idx_q], self._v_max[jt.idx_v]]
            ji += [kinds[jt.kind], self.parents[j], int(self._has_cfg_limit[jt.idx_q])]
        # every frame, in order (frame ids are part of the model's identity)
        ff, fi, fnames = [], [], []
        for f in self.frames:
            ff += list(f.placement.rotation.reshape(9)) + list(f.placement.translation)
            fi += [f.parentJoint]
            fnames.append([f.name, f.type])
        mf = []
        for j in range(len(self.joints)):
            mf += [self.masses[j]] + list(self.coms[j])
        floats = np.array(jf + ff + mf, dtype=np.float64)
        ints = np.array([int(self.free_flyer), len(self.joints) - first, len(fnames)] + ji + fi, dtype=np.int64)
        text = json.dumps({"name": self.name, "joints": self.names[first:], "frames": fnames}).encode("utf-8")
        return floats, ints, text

    @staticmethod
    def unpack(floats, ints, text) -> "Model":
        import json

        floats = np.asarray(floats, dtype=np.float64)
        ints = [int(x) for x in np.asarray(ints).reshape(-1)]
        meta = json.loads(bytes(text).decode("utf-8"))
        free_flyer, nj, nf = bool(ints[0]), ints[1], ints[2]
        m = Model(meta["name"], free_flyer=free_flyer)
        kinds = ["revolute", "prismatic"]
        fpos, ipos = 0, 3
        for k in range(nj):
            rec = floats[fpos:fpos + 18]
            fpos += 18
            kind, parent, has_cfg = ints[ipos:ipos + 3]
            ipos += 3
            jid = m.add_joint(meta["joints"][k], parent, SE3(rec[:9].reshape(3, 3), rec[9:12]), rec[12:15], kinds[kind],
                              lower=rec[15], upper=rec[16], velocity=rec[17])
            m._has_cfg_limit[m.joints[jid].idx_q] = bool(has_cfg)
        frames = []
        for k in range(nf):
            rec = floats[fpos:fpos + 12]
            fpos += 12
            frames.append(Frame(meta["frames"][k][0], ints[ipos], SE3(rec[:9].reshape(3, 3), rec[9:12]), meta["frames"][k][1]))
            ipos += 1
        m.frames = frames  # replaces the frames the constructor and add_joint created, same content, original order
        for j in range(len(m.joints)):
            m.masses[j] = float(floats[fpos])
            m.coms[j] = np.array(floats[fpos + 1:fpos + 4])
            fpos += 4
        m._version += 1
        return m

    # -- Pinocchio-style queries ------------------------------------------
    @property
    def njoints(self) -> int:
        return len(self.joints)

    @property
    def nframes(self) -> int:
        return len(self.frames)

    @property
    def lowerPositionLimit(self) -> np.ndarray:
        if not hasattr(self, "_lower_arr") or len(self._lower_arr) != self.nq:
            self._lower_arr = np.array(self._q_min, dtype=np.float64)
        return self._lower_arr

    @lowerPositionLimit.setter
    def lowerPositionLimit(self, value):
        self._lower_arr = np.array(value, dtype=np.float64)

    @property
    def upperPositionLimit(self) -> np.ndarray:
        if not hasattr(self, "_upper_arr") or len(self._upper_arr) != self.nq:
            self._upper_arr = np.array(self._q_max, dtype=np.float64)
        return self._upper_arr

    @upperPositionLimit.setter
    def upperPositionLimit(self, value):
        self._upper_arr = np.array(value, dtype=np.float64)

    @property
    def velocityLimit(self) -> np.ndarray:
        if not hasattr(self, "_vel_arr") or len(self._vel_arr) != self.nv:
            self._vel_arr = np.array(self._v_max, dtype=np.float64)
        return self._vel_arr

    @velocityLimit.setter
    def velocityLimit(self, value):
        self._vel_arr = np.array(value, dtype=np.float64)

    def hasConfigurationLimit(self) -> np.ndarray:
        return np.array(self._has_cfg_limit, dtype=bool)

    def _frame_index(self) -> dict:
        """name -> first frame id, rebuilt when frames were added."""
        cache = self.__dict__.get("_frame_ids")
        if cache is None or cache[0] != len(self.frames):
            ids: dict = {}
            for i, f in enumerate(self.frames):
                ids.setdefault(f.name, i)
            cache = (len(self.frames), ids)
            self.__dict__["_frame_ids"] = cache
        return cache[1]

    def existFrame(self, name: str) -> bool:
        return name in self._frame_index()

    def getFrameId(self, name: str) -> int:
        return self._frame_index().get(name, len(self.frames))  # Pinocchio returns nframes when absent

    def existJointName(self, name: str) -> bool:
        return name in self.names

    def getJointId(self, name: str) -> int:
        return self.names.index(name) if name in self.names else len(self.names)

    def createData(self) -> Data:
        return Data(self)

    # -- flat table for the C-ABI and the oracle --------------------------
    def table(self) -> types.SimpleNamespace:
        """Plain arrays in "1-dof joint" indexing (the free-flyer is implicit).

        Bodies: ``-2`` universe, ``-1`` root body, ``j >= 0`` body of joint j;
        ``mass``/``com`` are indexed by ``body + 1``.  Limits are read from the
        *current* ``upper/lowerPositionLimit`` / ``velocityLimit`` arrays.
        """
        first = 2 if self.free_flyer else 1  # Pinocchio id of 1-dof joint 0
        nj = len(self.joints) - first

        def body_of(pin_id: int) -> int:
            if pin_id >= first:
                return pin_id - first
            if self.free_flyer:
                return -1 if pin_id == 1 else -2
            return -1

        t = types.SimpleNamespace()
        t.name = self.name
        t.njoints = nj
        t.free_flyer = self.free_flyer
        t.nq, t.nv = self.nq, self.nv
        t.parent = np.array([body_of(self.parents[first + j]) for j in range(nj)], dtype=np.int32)
        if nj and (t.parent >= np.arange(nj)).any():
            raise ValueError("joints must be ordered parents-first")
        t.parent = np.maximum(t.parent, -1)  # 1-dof joints hang off the root body at most
        t.jtype = np.array(
            [REVOLUTE if self.joints[first + j].kind == "revolute" else PRISMATIC for j in range(nj)],
            dtype=np.int32,
        )
        t.joint_R = np.array([self.jointPlacements[first + j].rotation for j in range(nj)]).reshape(nj, 3, 3)
        t.joint_p = np.array([self.jointPlacements[first + j].translation for j in range(nj)]).reshape(nj, 3)
        t.axis = np.array([self.axes[first + j] for j in range(nj)]).reshape(nj, 3)
        t.joint_names = [self.names[first + j] for j in range(nj)]
        t.q_min = np.array(self.lowerPositionLimit, dtype=np.float64)
        t.q_max = np.array(self.upperPositionLimit, dtype=np.float64)
        t.v_max = np.array(self.velocityLimit, dtype=np.float64)
        nf = len(self.frames)
        t.nframes = nf
        t.frame_names = [f.name for f in self.frames]
        t.frame_body = np.array([body_of(f.parentJoint) for f in self.frames], dtype=np.int32)
        t.frame_R = np.array([f.placement.rotation for f in self.frames]).reshape(nf, 3, 3)
        t.frame_p = np.array([f.placement.translation for f in self.frames]).reshape(nf, 3)
        # inertias: universe-attached mass never moves and is ignored by the
        # CoM, exactly as Pinocchio ignores the universe's inertia
        t.mass = np.array([self.masses[first - 1] if self.free_flyer else 0.0] + [self.masses[first + j] for j in range(nj)])
        t.com = np.array([self.coms[first - 1] if self.free_flyer else np.zeros(3)] + [self.coms[first + j] for j in range(nj)]).reshape(nj + 1, 3)
        return t


class RobotWrapper:
    """``pin.RobotWrapper`` look-alike: ``.model``, ``.data``, ``.q0``."""

    def __init__(self, model: Model):
        self.model = model
        self.data = model.createData()
        self.q0 = neutral(model)

    @property
    def nq(self):
        return self.model.nq

    @property
    def nv(self):
        return self.model.nv


def neutral(model: Model) -> np.ndarray:
    """``pin.neutral``: zeros, identity quaternion for a free-flyer."""
    q = np.zeros(model.nq)
    if model.free_flyer:
        q[6] = 1.0
    return q


# ---------------------------------------------------------------------------
# URDF
# ---------------------------------------------------------------------------


def _origin(elem) -> SE3:
    if elem is None:
        return SE3()
    xyz = [float(s) for s in elem.get("xyz", "0 0 0").split()]
    rpy = [float(s) for s in elem.get("rpy", "0 0 0").split()]
    return SE3(rpy_to_matrix(*rpy), xyz)


def model_from_urdf_string(xml: str, root_joint=None, name: Optional[str] = None) -> Model:
    """Build a :class:`Model` from URDF text.

    Traversal is depth-first with siblings in alphabetical order of joint
    name (the order Pinocchio's URDF parser produces).  ``root_joint`` is
    ``None`` (fixed base) or a :class:`JointModelFreeFlyer` / ``"free_flyer"``.
    """
    robot = ET.fromstring(xml)
    free_flyer = root_joint is not None
    model = Model(name or robot.get("name", "robot"), free_flyer=free_flyer)

    links = {l.get("name"): l for l in robot.findall("link")}
    joints = robot.findall("joint")
    children = {}
    child_links = set()
    for j in joints:
        for end in ("parent", "child"):
            el = j.find(end)
            if el is None or el.get("link") not in links:
                raise ValueError(f"URDF joint {j.get('name')!r}: {end} link "
                                 f"{None if el is None else el.get('link')!r} is not declared")
        if j.find("child").get("link") in child_links:
            raise ValueError(f"URDF link {j.find('child').get('link')!r} has two parent joints (kinematic loop)")
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_links.add(j.find("child").get("link"))
    roots = [n for n in links if n not in child_links]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")

    def add_link(link_name: str, joint_id: int, placement: SE3) -> None:
        model.add_frame(link_name, joint_id, placement, "BODY")
        inertial = links[link_name].find("inertial")
        if inertial is not None and inertial.find("mass") is not None:
            mass = float(inertial.find("mass").get("value", "0"))
            com = placement * _origin(inertial.find("origin")).translation
            model.append_inertia(joint_id, mass, com)

    def visit(link_name: str, joint_id: int, placement: SE3) -> None:
        for j in sorted(children.get(link_name, []), key=lambda e: e.get("name")):
            jtype = j.get("type")
            jname = j.get("name")
            child = j.find("child").get("link")
            T = placement * _origin(j.find("origin"))
            if jtype == "fixed":
                model.add_frame(jname, joint_id, T, "FIXED_JOINT")
                add_link(child, joint_id, T)
                visit(child, joint_id, T)
                continue
            if jtype not in ("revolute", "continuous", "prismatic"):
                raise NotImplementedError(f"URDF joint type {jtype!r} ({jname})")
            axis_el = j.find("axis")
            axis = [1.0, 0.0, 0.0] if axis_el is None else [float(s) for s in axis_el.get("xyz").split()]
            lim = j.find("limit")
            lower, upper, vel = -_INF, _INF, _INF
            if lim is not None:
                vel = float(lim.get("velocity", "inf"))
                if jtype != "continuous":
                    lower = float(lim.get("lower", "0"))
                    upper = float(lim.get("upper", "0"))
            elif jtype != "continuous":
                lower = upper = 0.0
            jid = model.add_joint(
                jname, joint_id, T, axis,
                kind="prismatic" if jtype == "prismatic" else "revolute",
                lower=lower, upper=upper, velocity=vel,
            )
            add_link(child, jid, SE3())
            visit(child, jid, SE3())

    root = roots[0]
    add_link(root, model.root_body, SE3())
    visit(root, model.root_body, SE3())
    return model


def load_urdf(path: str, root_joint=None) -> RobotWrapper:
    """``pin.RobotWrapper.BuildFromURDF`` look-alike
    (``/root/reference/examples/load_custom_urdf.py``)."""
    with open(path, "r", encoding="utf-8") as fh:
        return RobotWrapper(model_from_urdf_string(fh.read(), root_joint=root_joint))
