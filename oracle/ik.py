"""``build_ik`` / ``solve_ik`` of the oracle (fp64 numpy).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows ``pink/solve_ik.py`` line by line: ``H = damping I + sum H_t``,
``c = sum c_t`` (``:55-60``); default limits = configuration then velocity
(``:94-105``), ``limits=[]`` disables them, rows stacked in list order
(``:109-122``); solve, ``NoSolutionFound`` when the QP has no solution
(``:270-273``); ``v = dq / dt`` (``:274``).

``limits`` is ``None`` (model defaults) or a list of
``("configuration", gain)`` / ``("velocity", v_max_or_None)`` tuples.
"""

import numpy as np

from . import barriers as bar
from . import kinematics as kin
from . import limits as lim
from . import qp
from . import tasks as tk


def qp_objective(m, q, fk, tasks, damping):
    """``__compute_qp_objective`` (``pink/solve_ik.py:20-67``), batched."""
    q = np.asarray(q, dtype=np.float64)
    batch = q.shape[:-1]
    H = np.broadcast_to(damping * np.eye(m.nv), batch + (m.nv, m.nv)).copy()
    c = np.zeros(batch + (m.nv,))
    for task in tasks:
        H_t, c_t = tk.task_qp_objective(m, q, fk, task)
        H = H + H_t
        c = c + c_t
    return H, c


def qp_inequalities(m, q, limits, dt):
    """``__compute_qp_inequalities`` (``pink/solve_ik.py:70-122``), batched.

    Returns ``(G[m, nv], h[..., m])`` or ``(None, None)``; ``G`` does not
    depend on the instance for the two default limits."""
    if limits is None:
        limits = [("configuration", 0.5), ("velocity", None)]
    q = np.asarray(q, dtype=np.float64)
    batch = q.shape[:-1]
    G_list, h_list = [], []
    for kind, arg in limits:
        if kind == "configuration":
            rows = lim.configuration_limit_rows(m, q, arg)
        elif kind == "velocity":
            rows = lim.velocity_limit_rows(m, dt, arg)
        else:
            raise ValueError(kind)
        if rows is None:
            continue
        G_list.append(rows[0])
        h_list.append(np.broadcast_to(rows[1], batch + (rows[1].shape[-1],)))
    if not G_list:
        return None, None
    return np.vstack(G_list), np.concatenate(h_list, axis=-1)


def build_ik(m, q, tasks, dt, damping=1e-12, limits=None):
    """``pink.build_ik`` (``pink/solve_ik.py:152-203``): ``(H, c, G, h)``."""
    fk = kin.forward_kinematics(m, q)
    H, c = qp_objective(m, q, fk, tasks, damping)
    G, h = qp_inequalities(m, q, limits, dt)
    return H, c, G, h


def _slice_task(task, i):
    """Task with the per-instance target of row ``i`` (targets may be shared)."""
    tgt = task.get("target")
    t = dict(task)
    if task["type"] == "linear":
        return t
    if task["type"] in ("frame", "relative_frame"):
        R, p = np.asarray(tgt[0]), np.asarray(tgt[1])
        t["target"] = (R[i] if R.ndim == 3 else R, p[i] if p.ndim == 2 else p)
    else:
        a = np.asarray(tgt)
        t["target"] = a[i] if a.ndim == 2 else a
    return t


def _slice_task_range(task, lo, hi):
    """Task restricted to instances ``[lo, hi)`` (shared targets untouched)."""
    tgt = task.get("target")
    t = dict(task)
    if task["type"] == "linear":
        return t
    if task["type"] in ("frame", "relative_frame"):
        R, p = np.asarray(tgt[0]), np.asarray(tgt[1])
        t["target"] = (R[lo:hi] if R.ndim == 3 else R, p[lo:hi] if p.ndim == 2 else p)
    else:
        a = np.asarray(tgt)
        t["target"] = a[lo:hi] if a.ndim == 2 else a
    return t


def assemble(m, q, tasks, dt, damping=1e-12, limits=None, barriers=None, constraints=None):
    """Full QP of ONE instance, ``(H, c, G, h, A, b)``, in the reference's order
    (``pink/solve_ik.py:20-149``): tasks then barriers in the objective; limits
    then barriers in the inequalities; ``constraints`` (tasks) as equalities
    ``J dq = -gain e``.

    ``limits`` entries: ``("configuration", gain)``, ``("velocity", v_max)``,
    ``("floating_base", frame, twist_max)``, ``("acceleration", a_max, dq_prev)``.
    """
    q = np.asarray(q, dtype=np.float64)
    fk = kin.forward_kinematics(m, q)
    H, c = qp_objective(m, q, fk, tasks, damping)
    for b_ in barriers or []:
        H_b, c_b = bar.barrier_qp_objective(m, q, fk, b_)
        H = H + H_b
        c = c + c_b
    if limits is None:
        limits = [("configuration", 0.5), ("velocity", None)]
    G_list, h_list = [], []
    for entry in limits:
        kind = entry[0]
        if kind == "configuration":
            rows = lim.configuration_limit_rows(m, q, entry[1])
        elif kind == "velocity":
            rows = lim.velocity_limit_rows(m, dt, entry[1])
        elif kind == "floating_base":
            rows = lim.floating_base_velocity_rows(m, fk, entry[1], entry[2], dt)
        elif kind == "acceleration":
            rows = lim.acceleration_limit_rows(m, q, entry[1], entry[2], dt)
        else:
            raise ValueError(kind)
        if rows is not None:
            G_list.append(rows[0])
            h_list.append(rows[1])
    for b_ in barriers or []:
        G_b, h_b = bar.barrier_qp_inequalities(m, q, fk, b_, dt)
        G_list.append(G_b)
        h_list.append(h_b)
    G = np.vstack(G_list) if G_list else None
    h = np.concatenate(h_list) if h_list else None
    A_list, b_list = [], []
    for task in constraints or []:
        e, J = tk.task_error_jacobian(m, q, fk, task)
        A_list.append(J)
        b_list.append(-task.get("gain", 1.0) * e)
    A = np.vstack(A_list) if A_list else None
    b = np.concatenate(b_list) if b_list else None
    return H, c, G, h, A, b


def _slice_limits(limits, i):
    if limits is None:
        return None
    out = []
    for entry in limits:
        if entry[0] == "acceleration" and entry[2] is not None and np.asarray(entry[2]).ndim == 2:
            out.append((entry[0], entry[1], np.asarray(entry[2])[i]))
        else:
            out.append(entry)
    return out


def solve_ik(m, q, tasks, dt, damping=1e-12, limits=None, safety_break=True, barriers=None, constraints=None):
    """One IK step of one instance: ``(v, status)``.

    ``status``: 0 ok, 1 no QP solution (``NoSolutionFound``,
    ``pink/solve_ik.py:271-273``), 2 outside limits with ``safety_break``
    (``pink/configuration.py:186-194``; the step is then not solved)."""
    q = np.asarray(q, dtype=np.float64)
    if safety_break and bool(lim.check_limits(m, q)):
        return np.zeros(m.nv), 2
    H, c, G, h, A, b = assemble(m, q, tasks, dt, damping, limits, barriers, constraints)
    if h is not None and not np.all(np.isfinite(h)):
        return np.zeros(m.nv), 1
    res = qp.solve_qp(H, c, G, h, A, b)
    if not res.found:
        return np.zeros(m.nv), 1
    return res.x / dt, 0


def solve_ik_batch(m, q, tasks, dt, damping=1e-12, limits=None, safety_break=True, barriers=None, constraints=None):
    """Loop of ``solve_ik`` over the leading dimension of ``q``."""
    q = np.asarray(q, dtype=np.float64)
    B = q.shape[0]
    v = np.zeros((B, m.nv))
    status = np.zeros(B, dtype=np.int32)
    for i in range(B):
        v[i], status[i] = solve_ik(
            m, q[i], [_slice_task(t, i) for t in tasks], dt, damping, _slice_limits(limits, i), safety_break,
            barriers, [_slice_task(t, i) for t in constraints or []],
        )
    return v, status


def kkt_check_batch(H, c, G, h, x, bound_tol=0.0):
    """Vectorised optimality certificate of ``x[B, nv]`` for the batch of QPs
    ``(H[B], c[B], G, h[B])`` whose rows are all ``+-e_i`` (box rows, as
    produced by the two default limits).

    ``bound_tol``: extra absolute slack when deciding that a coordinate sits on a
    bound (fp32 evaluates ``gain * (q_lim - q)`` with ~1e-7 absolute rounding).

    Returns per-instance ``(stationarity, primal_violation)`` where
    stationarity is the largest KKT violation given the best admissible
    multipliers: with ``g = H x + c``, a coordinate may have ``g_i < 0`` only
    if it sits on an upper bound and ``g_i > 0`` only on a lower bound.
    """
    g = np.einsum("bij,bj->bi", H, x) + c
    nv = x.shape[-1]
    hi = np.full(x.shape, np.inf)
    lo = np.full(x.shape, -np.inf)
    if G is not None:
        for r in range(G.shape[0]):
            nz = np.nonzero(G[r])[0]
            assert nz.size == 1 and abs(abs(G[r, nz[0]]) - 1.0) < 1e-15
            i = nz[0]
            if G[r, i] > 0:
                hi[:, i] = np.minimum(hi[:, i], h[:, r])
            else:
                lo[:, i] = np.maximum(lo[:, i], -h[:, r])
    prim = np.maximum(np.maximum(x - hi, lo - x), 0.0).max(axis=-1)
    scale = 1e-9 + 1e-6 * np.abs(x) + bound_tol
    at_hi = x >= hi - scale
    at_lo = x <= lo + scale
    # admissible: g<=0 at upper bound, g>=0 at lower bound, g=0 when free
    viol = np.abs(g)
    viol = np.where(at_hi & (g <= 0), 0.0, viol)
    viol = np.where(at_lo & (g >= 0), 0.0, viol)
    return viol.max(axis=-1), prim, lo, hi
