"""Floating-base velocity limit
(``/root/reference/pink/limits/floating_base_velocity_limit.py``)."""

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from .limit import Limit


def _as_velocity_vector(values: Union[Sequence[float], float], label: str) -> np.ndarray:
    array = np.asarray(values, dtype=float)
    if array.ndim == 0:
        array = np.repeat(array, 3)
    array = array.flatten()
    if array.shape != (3,):
        raise ValueError(f"{label} must be a scalar or an iterable of length 3, got shape {array.shape}")
    return array


def _find_base_frame(model, base_frame: Optional[str]) -> Tuple[str, int]:
    if base_frame is not None:
        if not model.existFrame(base_frame):
            raise ValueError(f"Frame '{base_frame}' does not exist in the model.")
        return base_frame, model.getFrameId(base_frame)
    root_joint_id = model.getJointId("root_joint")
    for frame in model.frames:
        if frame.parentJoint == root_joint_id:
            return frame.name, model.getFrameId(frame.name)
    raise ValueError("Model does not expose a frame attached to 'root_joint'.")


class FloatingBaseVelocityLimit(Limit):
    r"""Bound the twist of a frame attached to the floating base:
    :math:`\pm J_{base}[:, root]\,\Delta q \leq \mathrm{d}t\,[v_{max}; \omega_{max}]`, rows with an
    infinite bound dropped (``floating_base_velocity_limit.py:59-148``)."""

    def __init__(self, model, base_frame: Optional[str], max_linear_velocity: Union[Sequence[float], float],
                 max_angular_velocity: Union[Sequence[float], float]):
        self.model = model
        self.linear_max = _as_velocity_vector(max_linear_velocity, "max_linear_velocity")
        self.angular_max = _as_velocity_vector(max_angular_velocity, "max_angular_velocity")
        self.twist_max = np.hstack([self.linear_max, self.angular_max])
        if not model.existJointName("root_joint"):
            raise ValueError("FloatingBaseVelocityLimit requires a floating-base root joint.")
        self.root_joint_id = model.getJointId("root_joint")
        root_joint = model.joints[self.root_joint_id]
        self.root_idx_v = root_joint.idx_v
        self.root_nv = root_joint.nv
        self.base_frame, self.frame_id = _find_base_frame(model, base_frame)
        if model.frames[self.frame_id].parentJoint != self.root_joint_id:
            raise ValueError(f"Frame '{self.base_frame}' is not attached to the root joint.")

    def compute_qp_inequalities(self, configuration, dt: float):
        """Evaluated by the CUDA library; ``G`` depends on the instance only through
        nothing at all (the frame is fixed in the base), but is returned per
        instance like barrier rows: ``G [B, m, nv]``, ``h [B, m]``."""
        if not np.isfinite(self.twist_max).any():
            return None
        from ..solve_ik import _dense_limit_rows

        return _dense_limit_rows(configuration, self, dt)
