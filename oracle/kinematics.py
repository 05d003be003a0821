"""Rigid-body kinematics of the oracle (fp64 numpy, batched over leading dims).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Restates what ``pink.Configuration.update`` obtains from Pinocchio
(``pink/configuration.py:163-164``: ``pin.computeJointJacobians`` +
``pin.updateFramePlacements``), ``Configuration.get_frame_jacobian``
(``pink/configuration.py:203-236``, LOCAL reference frame),
``pin.centerOfMass`` / ``pin.jacobianCenterOfMass``
(``pink/tasks/com_task.py:103-105,145-147``) and ``pin.difference`` /
``pin.integrate`` (``pink/tasks/posture_task.py:103-107``,
``pink/configuration.py:283``).

The model is read from a *model table* ``m`` (duck-typed; produced by
``pink_b200.model.Model.table()`` so that oracle and GPU consume the same
numbers) with fields

``njoints, free_flyer, nq, nv, parent[j], jtype[j] (0 revolute / 1 prismatic),
joint_R[j], joint_p[j], axis[j], q_min[nq], q_max[nq], v_max[nv],
frame_body[f], frame_R[f], frame_p[f], mass[b], com[b]``

where body ``-2`` is the universe, ``-1`` the root body (the floating base if
``free_flyer`` else the universe) and ``j >= 0`` the body of joint ``j``;
``mass/com`` are indexed by ``body + 1``.  Free-flyer configuration layout is
Pinocchio's ``[x y z qx qy qz qw | joints]`` with a body-frame twist
``[v(3) w(3) | joint rates]`` (``pink/configuration.py:220-228``).
"""

import numpy as np

from . import lie


def root_dims(m):
    """``(root_nq, root_nv)`` as ``pink.utils.get_root_joint_dim``
    (``pink/utils.py:40-54``)."""
    return (7, 6) if m.free_flyer else (0, 0)


def forward_kinematics(m, q):
    """World placements of the root body and every joint.

    Returns ``(R_root, p_root, R[..., j, 3, 3], p[..., j, 3])`` following
    ``oMi[j] = oMi[parent] X_j exp(S_j q_j)`` (FK).
    """
    q = np.asarray(q, dtype=np.float64)
    batch = q.shape[:-1]
    rq, _ = root_dims(m)
    if m.free_flyer:
        p_root = q[..., 0:3]
        quat = q[..., 3:7]
        quat = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
        R_root = lie.quat_to_matrix(quat)
    else:
        p_root = np.zeros(batch + (3,))
        R_root = np.broadcast_to(np.eye(3), batch + (3, 3)).copy()
    nj = m.njoints
    R = np.zeros(batch + (nj, 3, 3))
    p = np.zeros(batch + (nj, 3))
    for j in range(nj):
        par = int(m.parent[j])
        Rp, pp = (R_root, p_root) if par < 0 else (R[..., par, :, :], p[..., par, :])
        qj = q[..., rq + j]
        axis = np.asarray(m.axis[j], dtype=np.float64)
        if int(m.jtype[j]) == 0:
            Rm = lie.exp3(qj[..., None] * axis)
            pm = np.zeros(batch + (3,))
        else:
            Rm = np.broadcast_to(np.eye(3), batch + (3, 3))
            pm = qj[..., None] * axis
        Rl, pl = lie.se3_mul(np.asarray(m.joint_R[j]), np.asarray(m.joint_p[j]), Rm, pm)
        Rj, pj = lie.se3_mul(Rp, pp, Rl, pl)
        R[..., j, :, :] = Rj
        p[..., j, :] = pj
    return R_root, p_root, R, p


def _body_placement(m, fk, body):
    R_root, p_root, R, p = fk
    if body == -2 or (body == -1 and not m.free_flyer):
        batch = p_root.shape[:-1]
        return np.broadcast_to(np.eye(3), batch + (3, 3)), np.zeros(batch + (3,))
    if body == -1:
        return R_root, p_root
    return R[..., body, :, :], p[..., body, :]


def frame_placement(m, fk, f):
    """``oMf = oMi[parent] X_f`` (``pin.updateFramePlacements``;
    ``pink/configuration.py:238-254``)."""
    Rb, pb = _body_placement(m, fk, int(m.frame_body[f]))
    return lie.se3_mul(Rb, pb, np.asarray(m.frame_R[f]), np.asarray(m.frame_p[f]))


def supports(m, joint, body):
    """True if ``joint`` lies on the path from the root body to ``body``."""
    b = body
    while b >= 0:
        if b == joint:
            return True
        b = int(m.parent[b])
    return False


def frame_jacobian_local(m, fk, f):
    """``pin.getFrameJacobian(model, data, f, LOCAL)``
    (``pink/configuration.py:233-235``): ``(..., 6, nv)`` with rows
    ``[linear; angular]`` expressed in the frame's own basis."""
    R_root, p_root, R, p = fk
    Rf, pf = frame_placement(m, fk, f)
    batch = pf.shape[:-1]
    J = np.zeros(batch + (6, m.nv))
    body = int(m.frame_body[f])
    _, rv = root_dims(m)
    Rft = np.swapaxes(Rf, -1, -2)
    if m.free_flyer and body != -2:
        # columns of the floating base: Ad_{(oM_root^-1 oMf)^-1}
        Rrf, prf = lie.se3_act_inv(R_root, p_root, Rf, pf)
        J[..., :, 0:6] = lie.action_inverse(Rrf, prf)
    for j in range(m.njoints):
        if not supports(m, j, body):
            continue
        axis_w = np.einsum("...ij,j->...i", R[..., j, :, :], np.asarray(m.axis[j], dtype=np.float64))
        if int(m.jtype[j]) == 0:
            lin = np.cross(axis_w, pf - p[..., j, :])
            ang = axis_w
        else:
            lin = axis_w
            ang = np.zeros_like(axis_w)
        J[..., 0:3, rv + j] = np.einsum("...ij,...j->...i", Rft, lin)
        J[..., 3:6, rv + j] = np.einsum("...ij,...j->...i", Rft, ang)
    return J


def center_of_mass(m, fk):
    """``pin.centerOfMass`` in the world frame (``pink/tasks/com_task.py:123-126``)."""
    mass = np.asarray(m.mass, dtype=np.float64)
    total = mass.sum()
    acc = 0.0
    for b in range(-1, m.njoints):
        if mass[b + 1] == 0.0:
            continue
        Rb, pb = _body_placement(m, fk, b)
        acc = acc + mass[b + 1] * (
            pb + np.einsum("...ij,j->...i", Rb, np.asarray(m.com[b + 1], dtype=np.float64))
        )
    return acc / total


def com_jacobian(m, fk):
    """``pin.jacobianCenterOfMass``: ``(..., 3, nv)``, world frame
    (``pink/tasks/com_task.py:145-148``).  Column of joint ``j`` uses the mass
    and CoM of the subtree rooted at ``j`` (ComTask)."""
    R_root, p_root, R, p = fk
    mass = np.asarray(m.mass, dtype=np.float64)
    total = mass.sum()
    batch = p_root.shape[:-1]
    nj = m.njoints
    # world CoM (times mass) of each body, then accumulate subtrees leaf -> root
    sub_m = mass[1:].copy()
    sub_mc = np.zeros(batch + (nj, 3))
    for j in range(nj):
        sub_mc[..., j, :] = mass[j + 1] * (
            p[..., j, :]
            + np.einsum("...ij,j->...i", R[..., j, :, :], np.asarray(m.com[j + 1], dtype=np.float64))
        )
    for j in range(nj - 1, -1, -1):
        par = int(m.parent[j])
        if par >= 0:
            sub_m[par] += sub_m[j]
            sub_mc[..., par, :] += sub_mc[..., j, :]
    _, rv = root_dims(m)
    J = np.zeros(batch + (3, m.nv))
    if m.free_flyer:
        com = center_of_mass(m, fk)
        J[..., :, 0:3] = R_root
        r_local = np.einsum("...ji,...j->...i", R_root, com - p_root)
        J[..., :, 3:6] = -R_root @ lie.hat(r_local)
    for j in range(nj):
        if sub_m[j] == 0.0:
            continue
        axis_w = np.einsum("...ij,j->...i", R[..., j, :, :], np.asarray(m.axis[j], dtype=np.float64))
        if int(m.jtype[j]) == 0:
            c_sub = sub_mc[..., j, :] / sub_m[j]
            col = np.cross(axis_w, c_sub - p[..., j, :])
        else:
            col = axis_w
        J[..., :, rv + j] = (sub_m[j] / total) * col
    return J


def difference(m, q0, q1):
    """``pin.difference(model, q0, q1) = q1 (-) q0`` in the tangent space at q0
    (``pink/tasks/posture_task.py:103-107``,
    ``pink/limits/configuration_limit.py:111-116``)."""
    q0 = np.asarray(q0, dtype=np.float64)
    q1 = np.asarray(q1, dtype=np.float64)
    rq, rv = root_dims(m)
    batch = np.broadcast_shapes(q0.shape[:-1], q1.shape[:-1])
    out = np.zeros(batch + (m.nv,))
    if m.free_flyer:
        R0 = lie.quat_to_matrix(q0[..., 3:7])
        R1 = lie.quat_to_matrix(q1[..., 3:7])
        Rd, pd = lie.se3_act_inv(R0, q0[..., 0:3], R1, q1[..., 0:3])
        out[..., 0:6] = lie.log6(Rd, pd)
    out[..., rv:] = q1[..., rq:] - q0[..., rq:]
    return out


def integrate(m, q, dv):
    """``pin.integrate(model, q, dv) = q (+) dv`` (``pink/configuration.py:283``)."""
    q = np.asarray(q, dtype=np.float64)
    dv = np.asarray(dv, dtype=np.float64)
    rq, rv = root_dims(m)
    out = np.array(np.broadcast_to(q, np.broadcast_shapes(q.shape[:-1], dv.shape[:-1]) + (m.nq,)))
    if m.free_flyer:
        # Pinocchio's SE(3) integration on [p, quaternion]: p += R V(w) v,
        # quat <- quat * exp_quat(w), renormalised
        quat0 = q[..., 3:7] / np.linalg.norm(q[..., 3:7], axis=-1, keepdims=True)
        R0 = lie.quat_to_matrix(quat0)
        _, pe = lie.exp6(dv[..., 0:6])
        out[..., 0:3] = q[..., 0:3] + np.einsum("...ij,...j->...i", R0, pe)
        w = dv[..., 3:6]
        th = np.linalg.norm(w, axis=-1)
        small = th < 1e-8
        k = np.where(small, 0.5, np.sin(0.5 * th) / np.where(small, 1.0, th))
        dq = np.concatenate([k[..., None] * w, np.cos(0.5 * th)[..., None]], axis=-1)
        ax, ay, az, aw = (quat0[..., i] for i in range(4))
        bx, by, bz, bw = (dq[..., i] for i in range(4))
        quat = np.stack(
            [
                aw * bx + ax * bw + ay * bz - az * by,
                aw * by - ax * bz + ay * bw + az * bx,
                aw * bz + ax * by - ay * bx + az * bw,
                aw * bw - ax * bx - ay * by - az * bz,
            ],
            axis=-1,
        )
        out[..., 3:7] = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
    out[..., rq:] = q[..., rq:] + dv[..., rv:]
    return out


def neutral(m):
    """``pin.neutral(model)``."""
    q = np.zeros(m.nq)
    if m.free_flyer:
        q[6] = 1.0
    return q


def d_difference_arg1(m, q0, q1):
    """``pin.dDifference(model, q0, q1, ARG1)``: Jacobian of ``q1 (-) q0`` with
    respect to a LOCAL perturbation of ``q1``
    (``pink/tasks/linear_holonomic_task.py:186-192``): identity on 1-dof joints,
    ``Jlog6(M0^-1 M1)`` on the free-flyer block."""
    q0 = np.asarray(q0, dtype=np.float64)
    q1 = np.asarray(q1, dtype=np.float64)
    batch = np.broadcast_shapes(q0.shape[:-1], q1.shape[:-1])
    D = np.broadcast_to(np.eye(m.nv), batch + (m.nv, m.nv)).copy()
    if m.free_flyer:
        R0 = lie.quat_to_matrix(q0[..., 3:7])
        R1 = lie.quat_to_matrix(q1[..., 3:7])
        Rd, pd = lie.se3_act_inv(R0, q0[..., 0:3], R1, q1[..., 0:3])
        D[..., 0:6, 0:6] = lie.jlog6(Rd, pd)
    return D


def point_jacobian_world(m, fk, body, x):
    """Linear velocity (world axes) of the point ``x`` (world coordinates) rigidly
    attached to ``body``, as a ``(..., 3, nv)`` Jacobian.  This is
    ``J_p + skew(x - p_joint)^T J_w`` of ``pin.getJointJacobian(...,
    LOCAL_WORLD_ALIGNED)`` (``pink/barriers/self_collision_barrier.py:205-217``)."""
    R_root, p_root, R, p = fk
    x = np.asarray(x, dtype=np.float64)
    batch = x.shape[:-1]
    J = np.zeros(batch + (3, m.nv))
    _, rv = root_dims(m)
    if m.free_flyer and body != -2:
        # base twist is expressed in the base frame: v_world = R (v + w x (R^T (x - p)))
        xl = np.einsum("...ji,...j->...i", R_root, x - p_root)
        for k in range(3):
            ek = np.zeros(3)
            ek[k] = 1.0
            J[..., :, k] = R_root[..., :, k]
            J[..., :, 3 + k] = np.einsum("...ij,...j->...i", R_root, np.cross(ek, xl))
    for j in range(m.njoints):
        if body < 0 or not supports(m, j, body):
            continue
        axis_w = np.einsum("...ij,j->...i", R[..., j, :, :], np.asarray(m.axis[j], dtype=np.float64))
        if int(m.jtype[j]) == 0:
            J[..., :, rv + j] = np.cross(axis_w, x - p[..., j, :])
        else:
            J[..., :, rv + j] = axis_w
    return J
