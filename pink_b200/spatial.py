"""Host-side SE(3) value type with the subset of ``pinocchio.SE3`` that Pink's
public API exposes to users (targets passed to ``FrameTask.set_target``,
values returned by ``Configuration.get_transform_frame_to_world``;
``/root/reference/pink/tasks/frame_task.py:129-146``,
``/root/reference/pink/configuration.py:238-271``).

This is set-up/marshalling code (fp64 numpy, a handful of 3x3 products); the
per-instance arithmetic of the IK path runs in the CUDA kernels only.
"""

from __future__ import annotations

import numpy as np


def _skew(p):
    return np.array(
        [[0.0, -p[2], p[1]], [p[2], 0.0, -p[0]], [-p[1], p[0], 0.0]]
    )


def rpy_to_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """``pin.utils.rpyToMatrix``: ``Rz(yaw) Ry(pitch) Rx(roll)`` (URDF convention)."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


class SE3:
    """Rigid transform ``(rotation, translation)``; ``T_AB`` maps B to A."""

    __slots__ = ("rotation", "translation")

    def __init__(self, rotation=None, translation=None):
        if rotation is not None and translation is None:
            M = np.asarray(rotation, dtype=np.float64)
            if M.shape == (4, 4):
                rotation, translation = M[:3, :3], M[:3, 3]
            elif M.shape == (3, 4):
                rotation, translation = M[:, :3], M[:, 3]
            else:
                raise ValueError("SE3 expects (R, p), a 4x4 or a 3x4 matrix")
        self.rotation = (
            np.eye(3) if rotation is None else np.array(rotation, dtype=np.float64).reshape(3, 3)
        )
        self.translation = (
            np.zeros(3) if translation is None else np.array(translation, dtype=np.float64).reshape(3)
        )

    # -- constructors ------------------------------------------------------
    @staticmethod
    def Identity() -> "SE3":
        return SE3()

    @staticmethod
    def Random(rng=None) -> "SE3":
        rng = np.random.default_rng() if rng is None else rng
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        x, y, z, w = q
        R = np.array(
            [
                [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
            ]
        )
        return SE3(R, rng.uniform(-1.0, 1.0, size=3))

    # -- group operations --------------------------------------------------
    def copy(self) -> "SE3":
        return SE3(self.rotation.copy(), self.translation.copy())

    def inverse(self) -> "SE3":
        Rt = self.rotation.T
        return SE3(Rt, -Rt @ self.translation)

    def __mul__(self, other):
        if isinstance(other, SE3):
            return SE3(
                self.rotation @ other.rotation,
                self.rotation @ other.translation + self.translation,
            )
        v = np.asarray(other, dtype=np.float64)
        return self.rotation @ v + self.translation

    def act(self, other):
        return self * other

    def actInv(self, other):
        """``self^-1 * other`` (``frame_task.py:181-183``)."""
        if isinstance(other, SE3):
            Rt = self.rotation.T
            return SE3(Rt @ other.rotation, Rt @ (other.translation - self.translation))
        v = np.asarray(other, dtype=np.float64)
        return self.rotation.T @ (v - self.translation)

    @property
    def action(self) -> np.ndarray:
        """``Ad_T`` on ``[linear; angular]`` twists."""
        A = np.zeros((6, 6))
        A[:3, :3] = self.rotation
        A[:3, 3:] = _skew(self.translation) @ self.rotation
        A[3:, 3:] = self.rotation
        return A

    @property
    def actionInverse(self) -> np.ndarray:
        return self.inverse().action

    # -- views -------------------------------------------------------------
    @property
    def homogeneous(self) -> np.ndarray:
        M = np.eye(4)
        M[:3, :3] = self.rotation
        M[:3, 3] = self.translation
        return M

    @property
    def np(self) -> np.ndarray:
        return self.homogeneous

    def as_3x4(self) -> np.ndarray:
        """Row-major ``[R | p]`` (the 12-float target layout of the C-ABI)."""
        return np.hstack([self.rotation, self.translation[:, None]])

    def __array__(self, dtype=None, copy=None):
        M = self.homogeneous
        return M if dtype is None else M.astype(dtype)

    def isApprox(self, other: "SE3", prec: float = 1e-12) -> bool:
        return bool(
            np.allclose(self.rotation, other.rotation, atol=prec)
            and np.allclose(self.translation, other.translation, atol=prec)
        )

    def __repr__(self):
        return f"SE3(R=\n{self.rotation},\n  p={self.translation})"
