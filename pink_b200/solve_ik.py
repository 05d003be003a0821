"""Build and solve the batched differential-IK problem.

Mirrors ``/root/reference/pink/solve_ik.py``: ``build_ik`` (``:152-203``) and
``solve_ik`` (``:206-275``) keep their names, argument order and defaults.  The
Python here only *describes* the problem (``PkProblemDesc``); forward
kinematics, task Jacobians, the ``H``/``c`` assembly, the limit rows and the QP
solve all run inside one CUDA kernel launch per call.
"""

from __future__ import annotations

import logging
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .exceptions import NoSolutionFound, NotWithinConfigurationLimits, PinkError
from .limits import ConfigurationLimit, Limit, VelocityLimit
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
                 check_tol: float = 1e-6) -> Tuple[Optional[ConfigurationLimit], Optional[VelocityLimit]]:
    nv = model.nv
    cfg = [l for l in limits if isinstance(l, ConfigurationLimit)]
    vel = [l for l in limits if isinstance(l, VelocityLimit)]
    other = [l for l in limits if not isinstance(l, (ConfigurationLimit, VelocityLimit))]
    if other:
        raise NotImplementedError(
            f"limits of type {[type(l).__name__ for l in other]} are not supported by the CUDA engine yet"
        )
    if len(cfg) > 1 or len(vel) > 1:
        raise NotImplementedError("at most one ConfigurationLimit and one VelocityLimit per solve")
    inf = float("inf")
    for i in range(_cabi.PK_MAX_NV):
        prob.cfg_lo[i], prob.cfg_hi[i], prob.vel[i] = -inf, inf, inf
        prob.chk_lo[i], prob.chk_hi[i] = -inf, inf
    prob.cfg_gain = 0.5
    if cfg:
        lo, hi = cfg[0].box_bounds()
        prob.cfg_gain = float(cfg[0].config_limit_gain)
        for i in range(nv):
            prob.cfg_lo[i], prob.cfg_hi[i] = float(lo[i]), float(hi[i])
    if vel:
        v = vel[0].box_bounds()
        for i in range(nv):
            prob.vel[i] = float(v[i])
    # Configuration.check_limits (configuration.py:181-201)
    root_nq, _ = get_root_joint_dim(model)
    shift = model.nq - nv
    q_max, q_min = model.upperPositionLimit, model.lowerPositionLimit
    for iq in range(root_nq, model.nq):
        if q_max[iq] <= q_min[iq] + check_tol:
            continue
        prob.chk_lo[iq - shift] = float(q_min[iq] - check_tol)
        prob.chk_hi[iq - shift] = float(q_max[iq] + check_tol)
    prob.safety_break = 1 if safety_break else 0
    return (cfg[0] if cfg else None), (vel[0] if vel else None)


def describe_problem(model, batch_size: int, tasks: Iterable, dt: float, damping: float, limits,
                     safety_break: bool):
    """Fill a ``PkProblemDesc`` from reference-style task / limit objects.

    Pure host code (no device access).  Returns ``(prob, parts, descs)`` where
    ``parts`` lists the per-instance target tensors in the order they must be
    concatenated along dim 1 to form the ``targets`` argument of the C-ABI.
    ``limits`` must already be a list (see :func:`_default_limits`)."""
    B = batch_size
    tasks = list(tasks)
    if len(tasks) > _cabi.PK_MAX_TASKS:
        raise PinkError(f"at most {_cabi.PK_MAX_TASKS} tasks per solve, got {len(tasks)}")
    prob = _cabi.PkProblemDesc()
    prob.ntasks = len(tasks)
    prob.dt = float(dt)
    prob.damping = float(damping)
    descs = [t._pk_describe(model) for t in tasks]
    shared_off = 0
    inst_parts: List[torch.Tensor] = []
    inst_off = 0
    for k, (task, d) in enumerate(zip(tasks, descs)):
        td = prob.tasks[k]
        td.type, td.frame, td.root = d["type"], d["frame"], d["root"]
        for i in range(6):
            c = float(d["cost6"][i])
            if c < 0.0:
                raise PinkError(f"negative cost in {task!r}")
            td.cost[i] = c
        td.gain = float(task.gain)
        td.lm_damping = float(task.lm_damping)
        tgt = d["target"]
        if isinstance(tgt, torch.Tensor):
            if tgt.shape[0] != B:
                raise PinkError(
                    f"{task!r} has {tgt.shape[0]} per-instance targets but the configuration holds {B} instances"
                )
            td.target_shared = 0
            td.target_offset = inst_off
            inst_parts.append(tgt)
            inst_off += tgt.shape[1]
        else:
            flat = np.asarray(tgt, dtype=np.float64).reshape(-1)
            if shared_off + flat.size > _cabi.PK_MAX_SHARED:
                raise PinkError("too many shared task targets for one solve")
            td.target_shared = 1
            td.target_offset = shared_off
            for i, x in enumerate(flat):
                prob.shared[shared_off + i] = float(x)
            shared_off += flat.size
    prob.target_stride = inst_off
    _fill_limits(prob, model, limits, safety_break)
    return prob, inst_parts, descs


def _pack_problem(configuration, tasks: Iterable, dt: float, damping: float, limits, safety_break: bool):
    """-> (PkProblemDesc, per-instance targets tensor on the device or None, task descriptions)."""
    prob, inst_parts, descs = describe_problem(
        configuration.model, configuration.batch_size, tasks, dt, damping,
        _default_limits(configuration, limits), safety_break,
    )
    targets = None
    if inst_parts:
        device = configuration.engine.device
        parts = [p.to(device=device, dtype=torch.float32) for p in inst_parts]
        # a single per-instance target is used in place (no copy, no extra traffic)
        targets = parts[0].contiguous() if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()
    return prob, targets, descs


def _reject_unsupported(barriers, constraints) -> None:
    if barriers:
        raise NotImplementedError("barriers are outside the scope of this engine (SURVEY section 8: config 4 only)")
    if constraints:
        raise NotImplementedError("equality constraints are a SURVEY section 8(f) 'next' row, not provided yet")


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


def _rows_from_h(configuration, limits: Sequence[Limit], h4: torch.Tensor):
    """Stack ``(G, h)`` in list order from the kernel's per-coordinate rows."""
    G_list, h_list = [], []
    for limit in limits:
        if limit.projection_matrix is None:
            continue
        idx = torch.as_tensor(np.asarray(limit.indices), device=h4.device, dtype=torch.long)
        base = 0 if isinstance(limit, ConfigurationLimit) else 2
        G_list.append(np.vstack([limit.projection_matrix, -limit.projection_matrix]))
        h_list.append(torch.cat([h4[:, base].index_select(1, idx), h4[:, base + 1].index_select(1, idx)], dim=1))
    if not G_list:
        return None, None
    return np.vstack(G_list), torch.cat(h_list, dim=1)


def _limit_rows(configuration, limits: Sequence[Limit], dt: float):
    prob, targets, _ = _pack_problem(configuration, [], dt, 0.0, list(limits), False)
    _, _, h4 = configuration.engine.build_ik(prob, configuration.q_device, targets)
    G, h = _rows_from_h(configuration, limits, h4)
    if G is None:
        return None
    return G, _unbatch(configuration, h)


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

    Returns ``Problem(P, q, G, h, None, None)``: ``P [B, nv, nv]``, ``q [B, nv]``
    and ``h [B, m]`` are device tensors (numpy without the batch dimension for a
    single configuration); ``G [m, nv]`` is the same for all instances.
    ``G`` and ``h`` are ``None`` when there is no inequality (``:120-121``).
    """
    _reject_unsupported(barriers, constraints)
    lims = _default_limits(configuration, limits)
    prob, targets, _ = _pack_problem(configuration, tasks, dt, damping, lims, False)
    H, c, h4 = configuration.engine.build_ik(prob, configuration.q_device, targets)
    G, h = _rows_from_h(configuration, lims, h4)
    return Problem(
        _unbatch(configuration, H),
        _unbatch(configuration, c),
        G,
        None if h is None else _unbatch(configuration, h),
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
        tasks, dt, damping, limits, safety_break: as in the reference.
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
    _reject_unsupported(barriers, constraints)
    prob, targets, _ = _pack_problem(configuration, tasks, dt, damping, limits, safety_break)
    v, status = configuration.engine.solve_ik(prob, configuration.q_device, targets, v=out)
    if return_status:
        return (v, status) if configuration.batched else (v[0], status[0])
    _raise_from_status(configuration, status, safety_break)
    return _unbatch(configuration, v)
