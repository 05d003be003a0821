"""SO(3)/SE(3) maps of the oracle (fp64 numpy, leading batch dimensions allowed).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Conventions are Pinocchio's, as used by the reference:

* an SE(3) element is a pair ``(R, p)``; ``T_AB`` maps B-coordinates to A;
* twists and Jacobian rows are ordered ``[linear(3); angular(3)]``
  (``pink/tasks/frame_task.py:152-155``);
* ``a.actInv(b) = a^-1 b`` (``pink/tasks/frame_task.py:181-183``);
* ``T.action`` is ``Ad_T``, ``T.actionInverse`` is ``Ad_{T^-1}``
  (``pink/tasks/relative_frame_task.py:239``).

The closed forms are those of SURVEY.md section 9 (Pinocchio ``log3``,
``log6``, ``Jlog3``, ``Jlog6``); ``tests/test_oracle_lie.py`` checks them
against ``scipy.linalg.logm/expm`` and central differences.
"""

import numpy as np

_SMALL = 1e-4  # below this angle the Taylor branches are used (fp64)
_NEAR_PI = 1e-2  # log3 switches to the diagonal formula within this of pi


def hat(w):
    """Skew matrix ``[w]x`` with ``[w]x v = w x v``; ``w`` is ``(..., 3)``."""
    w = np.asarray(w, dtype=np.float64)
    out = np.zeros(w.shape[:-1] + (3, 3))
    out[..., 0, 1] = -w[..., 2]
    out[..., 0, 2] = w[..., 1]
    out[..., 1, 0] = w[..., 2]
    out[..., 1, 2] = -w[..., 0]
    out[..., 2, 0] = -w[..., 1]
    out[..., 2, 1] = w[..., 0]
    return out


def exp3(w):
    """Rodrigues formula ``exp([w]x)``."""
    w = np.asarray(w, dtype=np.float64)
    t2 = np.sum(w * w, axis=-1)
    t = np.sqrt(t2)
    small = t < _SMALL
    ts = np.where(small, 1.0, t)
    a = np.where(small, 1.0 - t2 / 6.0 + t2 * t2 / 120.0, np.sin(ts) / ts)
    b = np.where(
        small, 0.5 - t2 / 24.0 + t2 * t2 / 720.0, (1.0 - np.cos(ts)) / (ts * ts)
    )
    W = hat(w)
    return (
        np.eye(3)
        + a[..., None, None] * W
        + b[..., None, None] * (W @ W)
    )


def log3(R):
    """Rotation vector of ``R`` and its angle.

    Mirrors Pinocchio ``log3`` (called through ``pin.log``,
    ``pink/tasks/frame_task.py:192``): ``theta = acos((tr R - 1)/2)``,
    ``w = theta/(2 sin theta) vee(R - R^T)``, with a diagonal-based formula
    near ``pi`` where ``R - R^T`` vanishes.
    """
    R = np.asarray(R, dtype=np.float64)
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    ct = np.clip(0.5 * (tr - 1.0), -1.0, 1.0)
    theta = np.arccos(ct)
    vee = np.stack(
        [
            R[..., 2, 1] - R[..., 1, 2],
            R[..., 0, 2] - R[..., 2, 0],
            R[..., 1, 0] - R[..., 0, 1],
        ],
        axis=-1,
    )
    # generic branch (also fine near zero with the Taylor factor)
    small = theta < _SMALL
    st = np.sin(theta)
    fac = np.where(
        small,
        0.5 * (1.0 + theta**2 / 6.0 + 7.0 * theta**4 / 360.0),
        0.5 * theta / np.where(small, 1.0, st),
    )
    w_gen = fac[..., None] * vee
    # near-pi branch: w_k = sign_k * theta * sqrt((R_kk - cos)/(1 - cos))
    diag = np.stack([R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]], axis=-1)
    one_m_c = np.maximum(1.0 - ct, 1e-300)
    mag = theta[..., None] * np.sqrt(
        np.maximum((diag - ct[..., None]) / one_m_c[..., None], 0.0)
    )
    sgn = np.where(vee >= 0.0, 1.0, -1.0)
    w_pi = sgn * mag
    near_pi = theta >= np.pi - _NEAR_PI
    w = np.where(near_pi[..., None], w_pi, w_gen)
    return w, theta


def _alpha_beta(theta):
    """Coefficients of the translational part of ``log6``."""
    t2 = theta * theta
    small = theta < _SMALL
    ts = np.where(small, 1.0, theta)
    st, ct = np.sin(ts), np.cos(ts)
    alpha = np.where(
        small, 1.0 - t2 / 12.0 - t2 * t2 / 720.0, ts * st / (2.0 * (1.0 - ct))
    )
    beta = np.where(
        small,
        1.0 / 12.0 + t2 / 720.0,
        1.0 / (ts * ts) - st / (2.0 * ts * (1.0 - ct)),
    )
    return alpha, beta


def log6(R, p):
    """Body twist ``[v; w]`` with ``exp6([v; w]) = (R, p)`` (``pin.log``)."""
    p = np.asarray(p, dtype=np.float64)
    w, theta = log3(R)
    alpha, beta = _alpha_beta(theta)
    wp = np.sum(w * p, axis=-1)
    v = (
        alpha[..., None] * p
        - 0.5 * np.cross(w, p)
        + (beta * wp)[..., None] * w
    )
    return np.concatenate([v, w], axis=-1)


def exp6(xi):
    """``(R, p)`` of the twist ``xi = [v; w]`` (``pin.exp6``)."""
    xi = np.asarray(xi, dtype=np.float64)
    v, w = xi[..., :3], xi[..., 3:]
    t2 = np.sum(w * w, axis=-1)
    t = np.sqrt(t2)
    small = t < _SMALL
    ts = np.where(small, 1.0, t)
    b = np.where(
        small, 0.5 - t2 / 24.0 + t2 * t2 / 720.0, (1.0 - np.cos(ts)) / (ts * ts)
    )
    c = np.where(
        small, 1.0 / 6.0 - t2 / 120.0 + t2 * t2 / 5040.0, (ts - np.sin(ts)) / ts**3
    )
    W = hat(w)
    V = np.eye(3) + b[..., None, None] * W + c[..., None, None] * (W @ W)
    return exp3(w), np.einsum("...ij,...j->...i", V, v)


def jlog3(w, theta):
    """``Jlog3``: ``log3(R exp3(d)) ~ log3(R) + Jlog3 d`` (right Jacobian inverse)."""
    _, a = _alpha_beta(theta)  # a = 1/t^2 - sin/(2 t (1-cos))
    t2 = theta * theta
    d = 1.0 - t2 * a
    return (
        a[..., None, None] * (w[..., :, None] * w[..., None, :])
        + d[..., None, None] * np.eye(3)
        + 0.5 * hat(w)
    )


def jlog6(R, p):
    """``pin.Jlog6``: ``log6(T exp6(d)) ~ log6(T) + Jlog6(T) d``.

    Block form ``[[A, B], [0, A]]`` with ``A = Jlog3`` and ``B = C A``
    (SURVEY.md section 9); used at ``pink/tasks/frame_task.py:226`` and
    ``pink/tasks/relative_frame_task.py:242``.
    """
    p = np.asarray(p, dtype=np.float64)
    w, theta = log3(R)
    t2 = theta * theta
    A = jlog3(w, theta)
    _, beta = _alpha_beta(theta)
    small = theta < _SMALL
    ts = np.where(small, 1.0, theta)
    st, ct = np.sin(ts), np.cos(ts)
    beta_dot = np.where(
        small,
        1.0 / 360.0 + t2 / 7560.0,
        -2.0 / ts**4 + (1.0 + st / ts) / (ts * ts * 2.0 * (1.0 - ct)),
    )
    wp = np.sum(w * p, axis=-1)
    v3 = (beta_dot * wp)[..., None] * w - (t2 * beta_dot + 2.0 * beta)[..., None] * p
    C = (
        v3[..., :, None] * w[..., None, :]
        + beta[..., None, None] * (w[..., :, None] * p[..., None, :])
        + (beta * wp)[..., None, None] * np.eye(3)
        + 0.5 * hat(p)
    )
    B = C @ A
    out = np.zeros(A.shape[:-2] + (6, 6))
    out[..., :3, :3] = A
    out[..., :3, 3:] = B
    out[..., 3:, 3:] = A
    return out


# ---- SE3 group operations on (R, p) pairs --------------------------------


def se3_mul(Ra, pa, Rb, pb):
    """``T_a T_b``."""
    return Ra @ Rb, np.einsum("...ij,...j->...i", Ra, pb) + pa


def se3_inv(R, p):
    """``T^-1``."""
    Rt = np.swapaxes(R, -1, -2)
    return Rt, -np.einsum("...ij,...j->...i", Rt, p)


def se3_act_inv(Ra, pa, Rb, pb):
    """``a.actInv(b) = a^-1 b``."""
    Rt = np.swapaxes(Ra, -1, -2)
    return Rt @ Rb, np.einsum("...ij,...j->...i", Rt, pb - pa)


def action(R, p):
    """``Ad_T = [[R, [p]x R], [0, R]]`` acting on ``[linear; angular]`` twists."""
    out = np.zeros(R.shape[:-2] + (6, 6))
    out[..., :3, :3] = R
    out[..., :3, 3:] = hat(p) @ R
    out[..., 3:, 3:] = R
    return out


def action_inverse(R, p):
    """``Ad_{T^-1}`` (``SE3.actionInverse``)."""
    Ri, pi = se3_inv(R, p)
    return action(Ri, pi)


def rpy_to_matrix(roll, pitch, yaw):
    """URDF fixed-axis roll-pitch-yaw: ``R = Rz(yaw) Ry(pitch) Rx(roll)``."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_to_matrix(qxyzw):
    """Rotation of a unit quaternion stored ``[x, y, z, w]`` (Pinocchio order,
    ``pink/configuration.py:224-226``)."""
    q = np.asarray(qxyzw, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def matrix_to_quat(R):
    """Unit quaternion ``[x, y, z, w]`` (w >= 0) of a rotation matrix."""
    R = np.asarray(R, dtype=np.float64)
    w, theta = log3(R)
    half = 0.5 * theta
    small = theta < _SMALL
    ts = np.where(small, 1.0, theta)
    k = np.where(small, 0.5 - theta**2 / 48.0, np.sin(half) / ts)
    return np.concatenate([k[..., None] * w, np.cos(half)[..., None]], axis=-1)
