"""Sphere collision model for :class:`pink_b200.barriers.SelfCollisionBarrier`.

The reference evaluates collision pairs of arbitrary geometry through coal
(``pin.GeometryModel`` / ``pin.computeDistances``,
``/root/reference/pink/configuration.py:145-161``).  The CUDA engine covers the
sphere--sphere case (the reference's own self-collision
test and examples use sphere-decomposed URDFs): every sphere is a frame at its
centre plus a radius, every collision pair a pair of spheres.
"""

import xml.etree.ElementTree as ET
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .spatial import SE3


class SphereCollisionModel:
    """Spheres rigidly attached to joints of a :class:`pink_b200.Model`.

    Adding a sphere registers a frame named ``sphere:<name>`` at its centre in
    ``model`` (before any :class:`Configuration` of that model is created)."""

    def __init__(self, model):
        self.model = model
        self.names: List[str] = []
        self.frames: List[int] = []
        self.radii: List[float] = []
        self.parents: List[int] = []
        self.links: List[str] = []  # URDF link each sphere belongs to ("" if unknown)
        self.collisionPairs: List[Tuple[int, int]] = []

    @classmethod
    def from_urdf_string(cls, model, xml: str) -> "SphereCollisionModel":
        """Spheres of the ``<collision><geometry><sphere radius=.../>`` elements of a
        sphere-decomposed URDF (such as ``iiwa14_spheres_collision.urdf`` used by
        ``tests/test_self_collision_barrier.py:26-31`` of the reference), attached to the
        joints of ``model`` (built from the same URDF by
        :func:`pink_b200.model.model_from_urdf_string`).  Collision elements with another
        geometry are skipped."""
        self = cls(model)
        for link in ET.fromstring(xml).findall("link"):
            lname = link.get("name")
            if not model.existFrame(lname):
                continue
            frame = model.frames[model.getFrameId(lname)]
            k = 0
            for col in link.findall("collision"):
                geom = col.find("geometry")
                sphere = None if geom is None else geom.find("sphere")
                if sphere is None:
                    continue
                origin = col.find("origin")
                xyz = [0.0, 0.0, 0.0] if origin is None else [float(t) for t in origin.get("xyz", "0 0 0").split()]
                center = frame.placement * np.asarray(xyz, dtype=float)
                self.add_sphere(f"{lname}_{k}", frame.parentJoint, center, float(sphere.get("radius")), link=lname)
                k += 1
        return self

    @classmethod
    def from_urdf(cls, model, path: str) -> "SphereCollisionModel":
        with open(path, "r", encoding="utf-8") as fh:
            return cls.from_urdf_string(model, fh.read())

    def add_sphere(self, name: str, parent_joint: int, center: Sequence[float], radius: float,
                   link: str = "") -> int:
        """Attach a sphere to joint ``parent_joint`` (Pinocchio joint id), centre in
        the joint frame; returns the sphere index."""
        if radius < 0.0:
            raise ValueError("sphere radius must be non-negative")
        frame = self.model.add_frame(f"sphere:{name}", parent_joint, SE3(np.eye(3), np.asarray(center, dtype=float)))
        self.names.append(name)
        self.frames.append(frame)
        self.radii.append(float(radius))
        self.parents.append(int(parent_joint))
        self.links.append(link)
        return len(self.names) - 1

    def add_collision_pair(self, first: int, second: int) -> None:
        n = len(self.names)
        if not (0 <= first < n and 0 <= second < n) or first == second:
            raise ValueError(f"invalid collision pair ({first}, {second})")
        self.collisionPairs.append((int(first), int(second)))

    def add_all_collision_pairs(self, skip_same_joint: bool = True) -> None:
        """``GeometryModel.addAllCollisionPairs``: every pair of spheres, except
        (by default) two spheres carried by the same joint."""
        for i in range(len(self.names)):
            for j in range(i + 1, len(self.names)):
                if skip_same_joint and self.parents[i] == self.parents[j]:
                    continue
                self.collisionPairs.append((i, j))

    def remove_collision_pairs(self, excluded: Sequence[Tuple[str, str]]) -> None:
        """Drop the pairs whose sphere names or link names match an excluded pair, the job
        of an SRDF ``disable_collisions`` list (``pin.removeCollisionPairs``)."""
        ex = {frozenset(p) for p in excluded}
        self.collisionPairs = [
            (i, j) for (i, j) in self.collisionPairs
            if frozenset((self.names[i], self.names[j])) not in ex
            and frozenset((self.links[i], self.links[j])) not in ex
        ]

    def remove_collision_pairs_from_srdf(self, path: str) -> None:
        """``<disable_collisions link1="a" link2="b"/>`` entries of an SRDF file."""
        with open(path, "r", encoding="utf-8") as fh:
            root = ET.fromstring(fh.read())
        self.remove_collision_pairs([(e.get("link1"), e.get("link2")) for e in root.iter("disable_collisions")])

    @property
    def geometryObjects(self) -> List["SphereObject"]:
        """``collision_model.geometryObjects[i].parentJoint`` of the reference's barrier
        (``pink/barriers/self_collision_barrier.py:181-191``)."""
        return [SphereObject(n, j, r) for n, j, r in zip(self.names, self.parents, self.radii)]

    # arrays for the C-ABI ------------------------------------------------------
    def pair_frames(self) -> np.ndarray:
        return np.array([[self.frames[i], self.frames[j]] for i, j in self.collisionPairs], dtype=np.int32).reshape(-1, 2)

    def pair_radii(self) -> np.ndarray:
        return np.array([[self.radii[i], self.radii[j]] for i, j in self.collisionPairs], dtype=np.float32).reshape(-1, 2)


class SphereObject:
    """One entry of :attr:`SphereCollisionModel.geometryObjects`."""

    def __init__(self, name: str, parent_joint: int, radius: float):
        self.name = name
        self.parentJoint = parent_joint
        self.radius = radius


class DistanceResult:
    """``collision_data.distanceResults[k]``: signed distance of one sphere pair and its
    nearest points in the world frame."""

    def __init__(self, min_distance: float, p1: np.ndarray, p2: np.ndarray):
        self.min_distance = min_distance
        self._p1, self._p2 = p1, p2

    def getNearestPoint1(self) -> np.ndarray:
        return self._p1

    def getNearestPoint2(self) -> np.ndarray:
        return self._p2


class SphereCollisionData:
    """Counterpart of ``pin.GeometryData`` for a :class:`SphereCollisionModel`, returned by
    :func:`pink_b200.utils.process_collision_pairs`.

    The solver never reads it (the kernels evaluate the pair distances themselves); it
    serves user code that inspects ``distanceResults`` as the reference's tests do
    (``tests/test_self_collision_barrier.py:139-159``).  The results are evaluated on
    access from the sphere-centre frames of the configuration bound last (unbatched
    configurations only)."""

    def __init__(self, collision_model: SphereCollisionModel):
        self.collision_model = collision_model
        self.enable_contact = True
        self._configuration = None

    def bind(self, configuration) -> None:
        self._configuration = configuration

    @property
    def distanceResults(self) -> List[DistanceResult]:
        cm, cfg = self.collision_model, self._configuration
        if cfg is None or cfg.batched:
            return [DistanceResult(float("nan"), np.full(3, np.nan), np.full(3, np.nan)) for _ in cm.collisionPairs]
        cfg._ensure_fk()
        out = []
        for i, j in cm.collisionPairs:
            ca = cfg.data.oMf[cm.frames[i]].translation
            cb = cfg.data.oMf[cm.frames[j]].translation
            gap = np.linalg.norm(cb - ca)
            n = (cb - ca) / gap if gap > 0.0 else np.array([1.0, 0.0, 0.0])
            out.append(DistanceResult(float(gap - cm.radii[i] - cm.radii[j]), ca + cm.radii[i] * n, cb - cm.radii[j] * n))
        return out
