"""Task errors, Jacobians and QP objectives of the oracle (fp64 numpy, batched).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

A task is a plain dict (so that tests can build them without product classes):

``{"type": "frame", "frame": f, "cost": [6], "gain": a, "lm_damping": l,
   "target": (R[...,3,3], p[...,3])}``
``{"type": "relative_frame", "frame": f, "root": r, ... "target": (R, p)}``
``{"type": "posture", "cost": w, "gain": a, "lm_damping": l, "target": q*[..., nq]}``
``{"type": "com", "cost": [3], ..., "target": c*[..., 3]}``
``{"type": "joint_velocity", "cost": w, ..., "target": dq_ref[..., nv - root_nv]}``
``{"type": "linear", "A": [p, nv], "b": [p], "q0": [nq] | None, "cost": [p], ...}``
"""

import numpy as np

from . import kinematics as kin
from . import lie


def frame_task_error(m, fk, task):
    """``e = log6(T_b^-1 T_t)`` (``pink/tasks/frame_task.py:178-193``)."""
    Rf, pf = kin.frame_placement(m, fk, task["frame"])
    Rt, pt = task["target"]
    Rbt, pbt = lie.se3_act_inv(Rf, pf, np.asarray(Rt, dtype=np.float64), np.asarray(pt, dtype=np.float64))
    return lie.log6(Rbt, pbt)


def frame_task_jacobian(m, fk, task):
    """``J = -Jlog6(T_t^-1 T_b) bJ_b`` (``pink/tasks/frame_task.py:216-227``)."""
    Rf, pf = kin.frame_placement(m, fk, task["frame"])
    Rt, pt = task["target"]
    Rtb, ptb = lie.se3_act_inv(np.asarray(Rt, dtype=np.float64), np.asarray(pt, dtype=np.float64), Rf, pf)
    Jf = kin.frame_jacobian_local(m, fk, task["frame"])
    return -lie.jlog6(Rtb, ptb) @ Jf


def _relative_transforms(m, fk, task):
    Rf, pf = kin.frame_placement(m, fk, task["frame"])
    Rr, pr = kin.frame_placement(m, fk, task["root"])
    # transform_frame_to_root = get_transform(frame, root) = oM_r^-1 oM_f
    # (pink/configuration.py:256-271)
    Rrf, prf = lie.se3_act_inv(Rr, pr, Rf, pf)
    Rt, pt = task["target"]
    Rtf, ptf = lie.se3_act_inv(np.asarray(Rt, dtype=np.float64), np.asarray(pt, dtype=np.float64), Rrf, prf)
    return (Rrf, prf), (Rtf, ptf)


def relative_frame_task_error(m, fk, task):
    """``e = log6(T_rt^-1 T_rf)`` (``pink/tasks/relative_frame_task.py:178-185``)."""
    _, (Rtf, ptf) = _relative_transforms(m, fk, task)
    return lie.log6(Rtf, ptf)


def relative_frame_task_jacobian(m, fk, task):
    """``J = Jlog6(T_tf) (fJ_f - Ad_{T_rf^-1} rJ_r)``
    (``pink/tasks/relative_frame_task.py:233-246``)."""
    (Rrf, prf), (Rtf, ptf) = _relative_transforms(m, fk, task)
    Jf = kin.frame_jacobian_local(m, fk, task["frame"])
    Jr = kin.frame_jacobian_local(m, fk, task["root"])
    return lie.jlog6(Rtf, ptf) @ (Jf - lie.action_inverse(Rrf, prf) @ Jr)


def posture_task_error(m, q, task):
    """``e = (q (-) q*)[root_nv:]`` (``pink/tasks/posture_task.py:100-107``)."""
    _, rv = kin.root_dims(m)
    return kin.difference(m, task["target"], q)[..., rv:]


def posture_task_jacobian(m, q, task):
    """``J = I[root_nv:, :]`` (``pink/tasks/posture_task.py:128-129``)."""
    _, rv = kin.root_dims(m)
    return np.eye(m.nv)[rv:, :]


def com_task_error(m, fk, task):
    """``e = com(q) - com*`` (``pink/tasks/com_task.py:120-127``)."""
    return kin.center_of_mass(m, fk) - np.asarray(task["target"], dtype=np.float64)


def com_task_jacobian(m, fk, task):
    """``J = jacobianCenterOfMass`` (``pink/tasks/com_task.py:142-148``)."""
    return kin.com_jacobian(m, fk)


def task_error_jacobian(m, q, fk, task):
    """Dispatch on the task type; returns ``(e[..., k], J[..., k, nv])``."""
    t = task["type"]
    if t == "frame":
        return frame_task_error(m, fk, task), frame_task_jacobian(m, fk, task)
    if t == "relative_frame":
        return (
            relative_frame_task_error(m, fk, task),
            relative_frame_task_jacobian(m, fk, task),
        )
    if t == "posture":
        e = posture_task_error(m, q, task)
        J = np.broadcast_to(posture_task_jacobian(m, q, task), e.shape[:-1] + (e.shape[-1], m.nv))
        return e, J
    if t == "com":
        return com_task_error(m, fk, task), com_task_jacobian(m, fk, task)
    if t == "joint_velocity":
        # e = dq_ref (zeros for the damping task), J = I[root_nv:, :]
        # (pink/tasks/joint_velocity_task.py:59-110, damping_task.py:33-43)
        _, rv = kin.root_dims(m)
        batch = np.asarray(q).shape[:-1]
        e = np.broadcast_to(np.asarray(task["target"], dtype=np.float64), batch + (m.nv - rv,))
        J = np.broadcast_to(np.eye(m.nv)[rv:, :], batch + (m.nv - rv, m.nv))
        return e, J
    if t == "linear":
        # e = A (q (-) q_0) - b, J = A dDifference(q_0, q, ARG1)
        # (pink/tasks/linear_holonomic_task.py:148-192; JointCouplingTask is the
        # one-row case with q_0 = neutral, joint_coupling_task.py:82-100)
        A = np.asarray(task["A"], dtype=np.float64)
        q0 = kin.neutral(m) if task.get("q0") is None else np.asarray(task["q0"], dtype=np.float64)
        e = np.einsum("pn,...n->...p", A, kin.difference(m, q0, q)) - np.asarray(task["b"], dtype=np.float64)
        J = A @ kin.d_difference_arg1(m, q0, q)
        return e, J
    raise ValueError(f"unknown task type {t!r}")


def task_weight(task, k):
    """Diagonal of ``W`` (``pink/tasks/task.py:148-157``): ``None`` -> ones,
    float -> repeated, otherwise the vector itself."""
    cost = task.get("cost")
    if cost is None:
        return np.ones(k)
    if isinstance(cost, float):
        return np.full(k, cost)
    return np.asarray(cost, dtype=np.float64)


def task_qp_objective(m, q, fk, task):
    """``(H, c)`` of one task, line by line as ``pink/tasks/task.py:145-166``."""
    e, J = task_error_jacobian(m, q, fk, task)
    minus_gain_error = -task.get("gain", 1.0) * e
    w = task_weight(task, J.shape[-2])
    weighted_jacobian = w[:, None] * J
    weighted_error = w * minus_gain_error
    mu = task.get("lm_damping", 0.0) * np.sum(weighted_error * weighted_error, axis=-1)
    H = np.swapaxes(weighted_jacobian, -1, -2) @ weighted_jacobian + mu[..., None, None] * np.eye(m.nv)
    c = -np.einsum("...k,...kn->...n", weighted_error, weighted_jacobian)
    return H, c
