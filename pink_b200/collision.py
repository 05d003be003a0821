"""Sphere collision model for :class:`pink_b200.barriers.SelfCollisionBarrier`.

The reference evaluates collision pairs of arbitrary geometry through coal
(``pin.GeometryModel`` / ``pin.computeDistances``,
``/root/reference/pink/configuration.py:145-161``).  The CUDA engine covers the
sphere--sphere case (the reference's own self-collision
test and examples use sphere-decomposed URDFs): every sphere is a frame at its
centre plus a radius, every collision pair a pair of spheres.
"""

from typing import List, Sequence, Tuple

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
        self.collisionPairs: List[Tuple[int, int]] = []

    def add_sphere(self, name: str, parent_joint: int, center: Sequence[float], radius: float) -> int:
        """Attach a sphere to joint ``parent_joint`` (Pinocchio joint id), centre in
        the joint frame; returns the sphere index."""
        if radius < 0.0:
            raise ValueError("sphere radius must be non-negative")
        frame = self.model.add_frame(f"sphere:{name}", parent_joint, SE3(np.eye(3), np.asarray(center, dtype=float)))
        self.names.append(name)
        self.frames.append(frame)
        self.radii.append(float(radius))
        self.parents.append(int(parent_joint))
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
        """Drop the pairs whose sphere names (or name prefixes up to the first '_')
        match an excluded pair, the job of an SRDF ``disable_collisions`` list."""
        ex = {frozenset(p) for p in excluded}
        self.collisionPairs = [
            (i, j) for (i, j) in self.collisionPairs if frozenset((self.names[i], self.names[j])) not in ex
        ]

    # arrays for the C-ABI ------------------------------------------------------
    def pair_frames(self) -> np.ndarray:
        return np.array([[self.frames[i], self.frames[j]] for i, j in self.collisionPairs], dtype=np.int32).reshape(-1, 2)

    def pair_radii(self) -> np.ndarray:
        return np.array([[self.radii[i], self.radii[j]] for i, j in self.collisionPairs], dtype=np.float32).reshape(-1, 2)
