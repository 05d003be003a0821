"""Build and solve the batched differential-IK problem.

Mirrors ``/root/reference/pink/solve_ik.py``: ``build_ik`` (``:152-203``) and
``solve_ik`` (``:206-275``) keep their names, argument order and defaults.  The
Python here only *describes* the problem (``PkProblemDesc``); forward
kinematics, task Jacobians, the ``H``/``c`` assembly, the limit rows and the QP
solve all run inside one CUDA kernel launch per call.
"""

from __future__ import annotations

import ctypes as C
import logging
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .exceptions import NoSolutionFound, NotWithinConfigurationLimits, PinkError
from .limits import AccelerationLimit, ConfigurationLimit, FloatingBaseVelocityLimit, Limit, VelocityLimit
from .utils import get_root_joint_dim

# The only QP back-end is the in-kernel active-set solver; the `solver`
# argument is kept for signature compatibility (the parity target is
# solver="quadprog": the QP is strictly convex, so every exact solver agrees).
_ACCEPTED_SOLVERS = None  # any string is accepted


class Problem:
    """Quadratic program ``(P, q, G, h, A, b)`` with the attribute names of
    ``qpsolvers.Problem`` (what ``pink.build_ik`` returns, ``solve_ik.py:202``)."""

    def __init__(self, P, q, G=None, h=None, A=None, b=None):
        self.P, self.q, self.G, self.h, self.A, self.b = P, q, G, h, A, b
        self.lb = None
        self.ub = None

    def unpack(self):
        return self.P, self.q, self.G, self.h, self.A, self.b, self.lb, self.ub


# ---------------------------------------------------------------------------
# problem description
# ---------------------------------------------------------------------------


def _default_limits(configuration, limits):
    """``limits=None`` -> model defaults, ``[]`` -> none (``solve_ik.py:94-105``)."""
    if limits is None:
        model = configuration.model
        limits = [model.configuration_limit, model.velocity_limit]
        floating_base_limit = getattr(model, "floating_base_velocity_limit", None)
        if floating_base_limit is not None:
            limits.append(floating_base_limit)
    return list(limits)


def _fill_limits(prob: _cabi.PkProblemDesc, model, limits: Sequence[Limit], safety_break: bool,
                 check_tol: float = 1e-6):
    """Box limits (configuration / velocity / acceleration) and the floating-base
    rows.  Returns the AccelerationLimit (its ``Delta_q_prev`` is a per-problem
    input placed by the caller) or ``None``."""
    nv = model.nv
    cfg = [l for l in limits if isinstance(l, ConfigurationLimit)]
    vel = [l for l in limits if isinstance(l, VelocityLimit)]
    acc = [l for l in limits if isinstance(l, AccelerationLimit)]
    fb = [l for l in limits if isinstance(l, FloatingBaseVelocityLimit)]
    other = [l for l in limits if l is not None
             and not isinstance(l, (ConfigurationLimit, VelocityLimit, AccelerationLimit, FloatingBaseVelocityLimit))]
    if other:
        raise NotImplementedError(
            f"limits of type {[type(l).__name__ for l in other]} are not supported by the CUDA engine"
        )
    if len(cfg) > 1 or len(vel) > 1 or len(acc) > 1 or len(fb) > 1:
        raise NotImplementedError("at most one limit of each kind per solve")
    # numpy views onto the fixed-size arrays of the descriptor (filled in bulk: this runs on
    # every solve_ik call)
    def view(field):
        return np.frombuffer(field, dtype=np.float32)

    cfg_lo, cfg_hi, vel_v = view(prob.cfg_lo), view(prob.cfg_hi), view(prob.vel)
    chk_lo, chk_hi = view(prob.chk_lo), view(prob.chk_hi)
    acc_max, acc_qlo, acc_qhi = view(prob.acc_max), view(prob.acc_qlo), view(prob.acc_qhi)
    cfg_lo[:], cfg_hi[:], vel_v[:] = -np.inf, np.inf, np.inf
    chk_lo[:], chk_hi[:] = -np.inf, np.inf
    acc_max[:], acc_qlo[:], acc_qhi[:] = np.inf, -np.inf, np.inf
    prob.cfg_gain = 0.5
    if cfg:
        lo, hi = cfg[0].box_bounds()
        prob.cfg_gain = float(cfg[0].config_limit_gain)
        cfg_lo[:nv], cfg_hi[:nv] = lo, hi
    if vel:
        vel_v[:nv] = vel[0].box_bounds()
    prob.acc_enabled = 0
    prob.acc_prev_offset = -1
    if acc and acc[0].projection_matrix is not None:
        a, qlo, qhi = acc[0].box_arrays()
        prob.acc_enabled = 1
        acc_max[:nv], acc_qlo[:nv], acc_qhi[:nv] = a, qlo, qhi
    prob.fb_enabled = 0
    if fb:
        prob.fb_enabled = 1
        prob.fb_frame = int(fb[0].frame_id)
        view(prob.fb_max)[:] = fb[0].twist_max
    # Configuration.check_limits (configuration.py:181-201)
    root_nq, _ = get_root_joint_dim(model)
    shift = model.nq - nv
    q_max, q_min = model.upperPositionLimit[root_nq:], model.lowerPositionLimit[root_nq:]
    ranged = q_max > q_min + check_tol
    chk_lo[root_nq - shift:nv] = np.where(ranged, q_min - check_tol, -np.inf)
    chk_hi[root_nq - shift:nv] = np.where(ranged, q_max + check_tol, np.inf)
    prob.safety_break = 1 if safety_break else 0
    return acc[0] if prob.acc_enabled else None


class _TargetLayout:
    """Places targets either in ``PkProblemDesc.shared`` (one value for all
    instances) or in the per-instance ``targets`` row."""

    def __init__(self, prob, batch_size: int):
        self.prob, self.B = prob, batch_size
        self.shared_off = 0
        self.inst_off = 0
        self.inst_parts: List[torch.Tensor] = []

    def place(self, tgt, owner) -> Tuple[int, int]:
        """-> (offset, shared flag)."""
        if isinstance(tgt, torch.Tensor):
            if tgt.shape[0] != self.B:
                raise PinkError(
                    f"{owner!r} has {tgt.shape[0]} per-instance targets but the configuration holds {self.B} instances"
                )
            off = self.inst_off
            self.inst_parts.append(tgt)
            self.inst_off += tgt.shape[1]
            return off, 0
        flat = np.asarray(tgt, dtype=np.float64).reshape(-1)
        if self.shared_off + flat.size > _cabi.PK_MAX_SHARED:
            raise PinkError("too many shared task targets for one solve")
        off = self.shared_off
        for i, x in enumerate(flat):
            self.prob.shared[off + i] = float(x)
        self.shared_off += flat.size
        return off, 1


def _fill_task(td, task, d, layout: _TargetLayout, extra: List[float]) -> None:
    td.type, td.frame, td.root = d["type"], d["frame"], d["root"]
    for i in range(6):
        c = float(d["cost6"][i])
        if c < 0.0:
            raise PinkError(f"negative cost in {task!r}")
        td.cost[i] = c
    td.gain = float(task.gain)
    td.lm_damping = float(task.lm_damping)
    td.rows = int(d.get("rows", 0))
    td.data_offset = 0
    if "data" in d:
        td.data_offset = len(extra)
        extra.extend(float(x) for x in d["data"])
    tgt = d["target"]
    if isinstance(tgt, torch.Tensor) or np.size(tgt) > 0:
        td.target_offset, td.target_shared = layout.place(tgt, task)
    else:
        td.target_offset, td.target_shared = 0, 1


def _fill_barrier(bd, barrier, d, dt_gain: bool, extra: List[float], pairs: List[int]) -> None:
    bd.type = d["type"]
    bd.frame = int(d.get("frame", 0))
    bd.frame2 = int(d.get("frame2", 0))
    bd.dim = int(d["dim"])
    bd.d_min = float(d.get("d_min", 0.0))
    gain = np.asarray(d["gain"], dtype=np.float64).reshape(-1)
    if bd.type == _cabi.PK_BARRIER_POSITION:
        idx = d["indices"]
        bd.nidx = len(idx)
        for k, i in enumerate(idx):
            bd.indices[k] = int(i)
        bd.has_min = 0 if d["p_min"] is None else 1
        bd.has_max = 0 if d["p_max"] is None else 1
        for k in range(len(idx)):
            bd.p_min[k] = 0.0 if d["p_min"] is None else float(d["p_min"][k])
            bd.p_max[k] = 0.0 if d["p_max"] is None else float(d["p_max"][k])
        if gain.size != bd.dim:
            raise PinkError(f"{barrier!r}: gain has {gain.size} entries for {bd.dim} rows")
        for k in range(bd.dim):
            bd.gain[k] = float(gain[k]) if dt_gain else 1.0
    else:
        bd.gain[0] = float(gain[0]) if dt_gain else 1.0
    bd.safe_displacement_gain = float(barrier.safe_displacement_gain) if dt_gain else 0.0
    bd.gain_function = int(barrier.gain_function_id) if dt_gain else _cabi.PK_GAINFN_IDENTITY
    if bd.type == _cabi.PK_BARRIER_SELF_COLLISION:
        pr = np.asarray(d["pairs"], dtype=np.int32).reshape(-1, 2)
        rd = np.asarray(d["radii"], dtype=np.float64).reshape(-1, 2)
        if pr.shape[0] > _cabi.PK_MAX_PAIRS:
            raise PinkError(f"at most {_cabi.PK_MAX_PAIRS} collision pairs, got {pr.shape[0]}")
        bd.npairs = pr.shape[0]
        bd.pair_offset = len(pairs) // 2
        bd.data_offset = len(extra)
        pairs.extend(int(x) for x in pr.reshape(-1))
        extra.extend(float(x) for x in rd.reshape(-1))


def describe_problem(model, batch_size: int, tasks: Iterable, dt: float, damping: float, limits,
                     safety_break: bool, barriers=None, constraints=None, collision_model=None,
                     raw_barriers: bool = False):
    """Fill a ``PkProblemDesc`` from reference-style task / limit / barrier objects.

    Pure host code (no device access).  Returns ``(prob, parts, descs)`` where
    ``parts`` lists the per-instance target tensors in the order they must be
    concatenated along dim 1 to form the ``targets`` argument of the C-ABI.
    ``limits`` must already be a list (see :func:`_default_limits`).
    ``raw_barriers`` describes barriers with unit gains, the identity class-K
    function and no objective term (used to export ``h`` and ``J_h`` themselves)."""
    B = batch_size
    tasks = list(tasks)
    barriers = list(barriers) if barriers else []
    constraints = list(constraints) if constraints else []
    if len(tasks) > _cabi.PK_MAX_TASKS:
        raise PinkError(f"at most {_cabi.PK_MAX_TASKS} tasks per solve, got {len(tasks)}")
    if len(barriers) > _cabi.PK_MAX_BARRIERS:
        raise PinkError(f"at most {_cabi.PK_MAX_BARRIERS} barriers per solve, got {len(barriers)}")
    if len(constraints) > _cabi.PK_MAX_CONSTRAINTS:
        raise PinkError(f"at most {_cabi.PK_MAX_CONSTRAINTS} equality-constraint tasks per solve, got {len(constraints)}")
    prob = _cabi.PkProblemDesc()
    prob.ntasks = len(tasks)
    prob.dt = float(dt)
    prob.damping = float(damping)
    layout = _TargetLayout(prob, B)
    extra: List[float] = []
    pairs: List[int] = []
    descs = [t._pk_describe(model) for t in tasks]
    for k, (task, d) in enumerate(zip(tasks, descs)):
        _fill_task(prob.tasks[k], task, d, layout, extra)
    prob.nconstraints = len(constraints)
    for k, task in enumerate(constraints):
        d = task._pk_describe(model)
        if d["type"] in (_cabi.PK_TASK_POSTURE, _cabi.PK_TASK_JOINT_VELOCITY):
            raise NotImplementedError(f"{task!r} cannot be used as an equality constraint on the CUDA engine")
        _fill_task(prob.constraints[k], task, d, layout, extra)
    prob.nbarriers = len(barriers)
    for k, barrier in enumerate(barriers):
        from .barriers import SelfCollisionBarrier

        d = (barrier._pk_describe(model, collision_model) if isinstance(barrier, SelfCollisionBarrier)
             else barrier._pk_describe(model))
        _fill_barrier(prob.barriers[k], barrier, d, not raw_barriers, extra, pairs)
    acc = _fill_limits(prob, model, limits, safety_break)
    if acc is not None:
        prev = acc._delta_q_prev_full
        if isinstance(prev, torch.Tensor) or np.any(np.asarray(prev) != 0.0):
            prob.acc_prev_offset, prob.acc_prev_shared = layout.place(prev, acc)
    prob.target_stride = layout.inst_off
    # constant data referenced by pointer: keep the arrays alive with the descriptor
    extra_np = np.ascontiguousarray(extra, dtype=np.float32)
    pairs_np = np.ascontiguousarray(pairs, dtype=np.int32)
    prob._keepalive = (extra_np, pairs_np)
    prob.n_extra = int(extra_np.size)
    prob.n_pairs = int(pairs_np.size // 2)
    prob.extra = extra_np.ctypes.data_as(C.POINTER(C.c_float)) if extra_np.size else None
    prob.pairs = pairs_np.ctypes.data_as(C.POINTER(C.c_int32)) if pairs_np.size else None
    return prob, layout.inst_parts, descs


def _pack_problem(configuration, tasks: Iterable, dt: float, damping: float, limits, safety_break: bool,
                  barriers=None, constraints=None, raw_barriers: bool = False):
    """-> (PkProblemDesc, per-instance targets tensor on the device or None, task descriptions)."""
    prob, inst_parts, descs = describe_problem(
        configuration.model, configuration.batch_size, tasks, dt, damping,
        _default_limits(configuration, limits), safety_break, barriers, constraints,
        getattr(configuration, "collision_model", None), raw_barriers,
    )
    targets = None
    if inst_parts:
        device = configuration.engine.device
        parts = [p.to(device=device, dtype=torch.float32) for p in inst_parts]
        # a single per-instance target is used in place (no copy, no extra traffic)
        targets = parts[0].contiguous() if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()
    return prob, targets, descs


# ---------------------------------------------------------------------------
# helpers behind Task.compute_* and Limit.compute_qp_inequalities
# ---------------------------------------------------------------------------


def _unbatch(configuration, t: torch.Tensor):
    return t if configuration.batched else t[0].cpu().numpy().astype(np.float64)


def _task_terms(configuration, task):
    prob, targets, descs = _pack_problem(configuration, [task], 1.0, 0.0, [], False)
    e, J = configuration.engine.task_terms(prob, 0, descs[0]["k"], configuration.q_device, targets)
    return _unbatch(configuration, e), _unbatch(configuration, J)


def _task_objective(configuration, task):
    prob, targets, _ = _pack_problem(configuration, [task], 1.0, 0.0, [], False)
    H, c, _ = configuration.engine.build_ik(prob, configuration.q_device, targets)
    return _unbatch(configuration, H), _unbatch(configuration, c)


def _acc_rows_for(configuration, limit: AccelerationLimit, dt: float):
    """``h`` of one AccelerationLimit alone: its box, read back from the kernel."""
    prob, targets, _ = _pack_problem(configuration, [], dt, 0.0, [limit], False)
    _, _, _, _, lo, hi = configuration.engine.constraint_rows(prob, configuration.q_device, targets)
    idx = torch.tensor(np.asarray(limit.indices), device=lo.device, dtype=torch.long)
    return torch.cat([hi.index_select(1, idx), -lo.index_select(1, idx)], dim=1)


def _stack_inequalities(configuration, limits: Sequence[Limit], barriers, dt: float, h4: torch.Tensor, rows):
    """``(G, h)`` in the reference's order (``solve_ik.py:109-122``): limits in
    list order, then barriers.  ``G`` is ``[m, nv]`` numpy when every row is the
    same for all instances (box limits only), else a ``[B, m, nv]`` tensor."""
    G_const, h_list, dense = [], [], []
    Gd, hd = (rows[0], rows[1]) if rows is not None else (None, None)
    cursor = 0
    nv = configuration.model.nv
    for limit in limits:
        if limit is None:
            continue
        if isinstance(limit, FloatingBaseVelocityLimit):
            n = 2 * int(np.isfinite(limit.twist_max).sum())
            if n == 0:
                continue
            dense.append((len(G_const), Gd[:, cursor:cursor + n]))
            G_const.append(None)
            h_list.append(hd[:, cursor:cursor + n])
            cursor += n
            continue
        if limit.projection_matrix is None:
            continue
        idx = torch.tensor(np.asarray(limit.indices), device=h4.device, dtype=torch.long)
        G_const.append(np.vstack([limit.projection_matrix, -limit.projection_matrix]))
        if isinstance(limit, AccelerationLimit):
            h_list.append(_acc_rows_for(configuration, limit, dt))
        else:
            base = 0 if isinstance(limit, ConfigurationLimit) else 2
            h_list.append(torch.cat([h4[:, base].index_select(1, idx), h4[:, base + 1].index_select(1, idx)], dim=1))
    for barrier in barriers or []:
        n = barrier.dim
        dense.append((len(G_const), Gd[:, cursor:cursor + n]))
        G_const.append(None)
        h_list.append(hd[:, cursor:cursor + n])
        cursor += n
    if not G_const:
        return None, None
    h = torch.cat(h_list, dim=1)
    if not dense:
        return np.vstack(G_const), h
    B = h.shape[0]
    blocks = []
    for k, g in enumerate(G_const):
        if g is None:
            blocks.append(dict(dense)[k])
        else:
            blocks.append(torch.as_tensor(g, device=h.device, dtype=torch.float32).unsqueeze(0).expand(B, -1, -1))
    return torch.cat(blocks, dim=1), h


def _rows_from_h(configuration, limits: Sequence[Limit], h4: torch.Tensor):
    """Stack ``(G, h)`` of box limits in list order from the kernel's per-coordinate rows."""
    return _stack_inequalities(configuration, limits, None, 1.0, h4, None)


def _limit_rows(configuration, limits: Sequence[Limit], dt: float):
    prob, targets, _ = _pack_problem(configuration, [], dt, 0.0, list(limits), False)
    _, _, h4 = configuration.engine.build_ik(prob, configuration.q_device, targets)
    G, h = _rows_from_h(configuration, limits, h4)
    if G is None:
        return None
    return G, _unbatch(configuration, h)


def _acceleration_rows(configuration, limit: AccelerationLimit, dt: float):
    G = np.vstack([limit.projection_matrix, -limit.projection_matrix])
    return G, _unbatch(configuration, _acc_rows_for(configuration, limit, dt))


def _dense_limit_rows(configuration, limit: FloatingBaseVelocityLimit, dt: float):
    prob, targets, _ = _pack_problem(configuration, [], dt, 0.0, [limit], False)
    G, hG, _, _, _, _ = configuration.engine.constraint_rows(prob, configuration.q_device, targets)
    n = 2 * int(np.isfinite(limit.twist_max).sum())
    return _unbatch(configuration, G[:, :n]), _unbatch(configuration, hG[:, :n])


def _barrier_rows(configuration, barrier, raw: bool, dt: float = 1.0):
    """``(G, h)`` of one barrier; ``raw``: ``(-J_h, h(q))`` (unit gains, dt = 1)."""
    prob, targets, _ = _pack_problem(configuration, [], 1.0 if raw else dt, 0.0, [], False, [barrier], None, raw)
    G, hG, _, _, _, _ = configuration.engine.constraint_rows(prob, configuration.q_device, targets)
    return _unbatch(configuration, G[:, :barrier.dim]), _unbatch(configuration, hG[:, :barrier.dim])


def _barrier_objective(configuration, barrier):
    prob, targets, _ = _pack_problem(configuration, [], 1.0, 0.0, [], False, [barrier])
    H, c, _ = configuration.engine.build_ik(prob, configuration.q_device, targets)
    return _unbatch(configuration, H), _unbatch(configuration, c)


# ---------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------


def build_ik(
    configuration,
    tasks: Iterable,
    dt: float,
    damping: float = 1e-12,
    limits: Optional[Iterable[Limit]] = None,
    barriers=None,
    constraints=None,
) -> Problem:
    r"""Build the quadratic program of every instance (``solve_ik.py:152-203``).

    Returns ``Problem(P, q, G, h, A, b)``: ``P [B, nv, nv]``, ``q [B, nv]`` and
    ``h [B, m]`` are device tensors (numpy without the batch dimension for a
    single configuration).  ``G [m, nv]`` is one numpy matrix when all rows are
    box rows (the same for every instance) and a ``[B, m, nv]`` tensor as soon as
    barriers or a floating-base limit add instance-dependent rows.  ``A, b`` are the
    equality rows of ``constraints`` (``None`` without).  ``G`` and ``h`` are
    ``None`` when there is no inequality (``:120-121``).
    """
    lims = _default_limits(configuration, limits)
    barriers = list(barriers) if barriers else []
    constraints = list(constraints) if constraints else []
    prob, targets, _ = _pack_problem(configuration, tasks, dt, damping, lims, False, barriers, constraints)
    H, c, h4 = configuration.engine.build_ik(prob, configuration.q_device, targets)
    rows = None
    A = b = None
    if barriers or constraints or any(isinstance(l, FloatingBaseVelocityLimit) for l in lims):
        rows = configuration.engine.constraint_rows(prob, configuration.q_device, targets)
        if constraints:
            meq = sum(t._pk_describe(configuration.model)["k"] for t in constraints)
            A = _unbatch(configuration, rows[2][:, :meq])
            b = _unbatch(configuration, rows[3][:, :meq])
    G, h = _stack_inequalities(configuration, lims, barriers, dt, h4, rows)
    if isinstance(G, torch.Tensor):
        G = _unbatch(configuration, G)
    return Problem(
        _unbatch(configuration, H),
        _unbatch(configuration, c),
        G,
        None if h is None else _unbatch(configuration, h),
        A,
        b,
    )


def _raise_from_status(configuration, status: torch.Tensor, safety_break: bool) -> None:
    bits = int(torch.bitwise_or(status, 0).max().item()) if status.numel() else 0
    if bits == 0:
        return
    st = status.cpu().numpy()
    if safety_break and (st & _cabi.PK_STATUS_OUT_OF_LIMITS).any():
        configuration.check_limits(safety_break=True)  # raises with joint / value / bounds
        raise NotWithinConfigurationLimits(-1, float("nan"), float("nan"), float("nan"))
    if (st & _cabi.PK_STATUS_OUT_OF_LIMITS).any():
        configuration.check_limits(safety_break=False)  # logs the warning of the reference
    bad = np.nonzero(st & (_cabi.PK_STATUS_NO_SOLUTION | _cabi.PK_STATUS_NOT_POSDEF))[0]
    if bad.size:
        raise NoSolutionFound(None, None, instances=bad if configuration.batched else None)
    if (st & _cabi.PK_STATUS_ITER_LIMIT).any():
        logging.warning(
            "active-set iteration cap reached on %d instance(s); velocities are feasible but may be sub-optimal",
            int(((st & _cabi.PK_STATUS_ITER_LIMIT) != 0).sum()),
        )


def solve_ik(
    configuration,
    tasks: Iterable,
    dt: float,
    solver: str = "quadprog",
    damping: float = 1e-12,
    limits: Optional[Iterable[Limit]] = None,
    barriers=None,
    constraints=None,
    safety_break: bool = True,
    return_status: bool = False,
    out: Optional[torch.Tensor] = None,
    **kwargs,
):
    r"""Compute a velocity tangent to every configuration of the batch
    (``solve_ik.py:206-275``).

    Args:
        configuration: :class:`pink_b200.Configuration` holding ``B`` instances.
        tasks, dt, damping, limits, barriers, constraints, safety_break: as in the
            reference (``barriers``: :mod:`pink_b200.barriers`; ``constraints``:
            tasks enforced as equalities ``J dq = -gain e``).
        solver: accepted for signature compatibility; the QP is solved inside
            the CUDA kernel (parity target: ``"quadprog"``).
        return_status: batched extension. When true, return ``(v, status)``
            without any host synchronisation; ``status[B]`` holds
            ``PK_STATUS_*`` bits.  When false (default) the call checks the
            status and raises :class:`NoSolutionFound` /
            :class:`NotWithinConfigurationLimits` like the reference.
        out: optional ``[B, nv]`` fp32 device tensor to write into.

    Returns:
        Velocity ``v = dq / dt``: ``[B, nv]`` device tensor, or ``[nv]`` numpy for
        a single (1-D) configuration.
    """
    prob, targets, _ = _pack_problem(configuration, tasks, dt, damping, limits, safety_break, barriers, constraints)
    v, status = configuration.engine.solve_ik(prob, configuration.q_device, targets, v=out)
    if return_status:
        return (v, status) if configuration.batched else (v[0], status[0])
    _raise_from_status(configuration, status, safety_break)
    return _unbatch(configuration, v)
