"""Inequality rows of the default limits (fp64 numpy, batched).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
"""

import numpy as np

from . import kinematics as kin


def configuration_limit_indices(m):
    """Tangent indices bounded by ``ConfigurationLimit.__init__``
    (``pink/limits/configuration_limit.py:50-72``): joints whose every
    coordinate has ``upper < 1e20`` and ``upper > lower + 1e-10``.  The
    free-flyer never qualifies (``hasConfigurationLimit`` is false on its
    quaternion and its translation bounds are infinite)."""
    rq, rv = kin.root_dims(m)
    q_min = np.asarray(m.q_min, dtype=np.float64)
    q_max = np.asarray(m.q_max, dtype=np.float64)
    idx = [
        rv + j
        for j in range(m.njoints)
        if q_max[rq + j] < 1e20 and q_max[rq + j] > q_min[rq + j] + 1e-10
    ]
    return np.array(idx, dtype=np.int64)


def velocity_limit_indices(m, v_max=None):
    """Tangent indices bounded by ``VelocityLimit.__init__``
    (``pink/limits/velocity_limit.py:61-78``): ``1e-10 < v_max < 1e20`` on every
    coordinate of the joint (so a free-flyer is bounded only if all six are)."""
    v_max = np.asarray(m.v_max if v_max is None else v_max, dtype=np.float64)
    ok = np.logical_and(v_max < 1e20, v_max > 1e-10)
    _, rv = kin.root_dims(m)
    idx = []
    if m.free_flyer and ok[0:6].all():
        idx.extend(range(6))
    idx.extend(rv + j for j in range(m.njoints) if ok[rv + j])
    return np.array(idx, dtype=np.int64)


def configuration_limit_rows(m, q, gain=0.5):
    """``G = [P; -P]``, ``h = [g (q_max (-) q)[idx]; -g (q_min (-) q)[idx]]``
    (``pink/limits/configuration_limit.py:108-121``); ``None`` if no joint is
    bounded."""
    idx = configuration_limit_indices(m)
    if idx.size == 0:
        return None
    rq, rv = kin.root_dims(m)
    q = np.asarray(q, dtype=np.float64)
    # pin.difference on 1-dof joints is a subtraction; root coordinates are
    # never selected, so they are skipped here.
    dq_max = np.asarray(m.q_max, dtype=np.float64)[rq:] - q[..., rq:]
    dq_min = np.asarray(m.q_min, dtype=np.float64)[rq:] - q[..., rq:]
    P = np.eye(m.nv)[idx]
    G = np.vstack([P, -P])
    h = np.concatenate([gain * dq_max[..., idx - rv], -gain * dq_min[..., idx - rv]], axis=-1)
    return G, h


def velocity_limit_rows(m, dt, v_max=None):
    """``G = [P; -P]``, ``h = dt [v_max; v_max]``
    (``pink/limits/velocity_limit.py:115-121``)."""
    idx = velocity_limit_indices(m, v_max)
    if idx.size == 0:
        return None
    v = np.asarray(m.v_max if v_max is None else v_max, dtype=np.float64)[idx]
    P = np.eye(m.nv)[idx]
    return np.vstack([P, -P]), np.concatenate([dt * v, dt * v])


def check_limits(m, q, tol=1e-6):
    """Boolean mask of instances outside limits
    (``pink/configuration.py:181-201``)."""
    rq, _ = kin.root_dims(m)
    q = np.asarray(q, dtype=np.float64)
    q_min = np.asarray(m.q_min, dtype=np.float64)
    q_max = np.asarray(m.q_max, dtype=np.float64)
    bad = np.zeros(q.shape[:-1], dtype=bool)
    for i in range(rq, m.nq):
        if q_max[i] <= q_min[i] + tol:
            continue
        bad |= (q[..., i] < q_min[i] - tol) | (q[..., i] > q_max[i] + tol)
    return bad


def floating_base_velocity_rows(m, fk, frame, twist_max, dt):
    """``FloatingBaseVelocityLimit.compute_qp_inequalities``
    (``pink/limits/floating_base_velocity_limit.py:118-148``): rows of the LOCAL
    Jacobian of a frame attached to the root joint, root columns only, for the
    finite entries of ``twist_max = [linear_max; angular_max]``."""
    twist_max = np.asarray(twist_max, dtype=np.float64)
    finite = np.isfinite(twist_max)
    if not finite.any():
        return None
    J = kin.frame_jacobian_local(m, fk, frame).copy()
    J[..., :, 6:] = 0.0
    rows = J[finite, :]
    bounds = dt * twist_max[finite]
    return np.vstack([rows, -rows]), np.concatenate([bounds, bounds])


def acceleration_limit_indices(m, a_max):
    """Tangent indices with ``1e-10 < a_max < 1e20``
    (``pink/limits/acceleration_limit.py:60-72``); the free-flyer qualifies only
    if all of its six coordinates do."""
    a_max = np.asarray(a_max, dtype=np.float64)
    ok = np.logical_and(a_max < 1e20, a_max > 1e-10)
    _, rv = kin.root_dims(m)
    idx = []
    if m.free_flyer and ok[0:6].all():
        idx.extend(range(6))
    idx.extend(rv + j for j in range(m.njoints) if ok[rv + j])
    return np.array(idx, dtype=np.int64)


def acceleration_limit_rows(m, q, a_max, dq_prev, dt):
    """``AccelerationLimit.compute_qp_inequalities``
    (``pink/limits/acceleration_limit.py:119-200``): finite-difference
    acceleration bound and braking distance to the configuration limits."""
    idx = acceleration_limit_indices(m, a_max)
    if idx.size == 0:
        return None
    rq, rv = kin.root_dims(m)
    q = np.asarray(q, dtype=np.float64)
    a = np.asarray(a_max, dtype=np.float64)[idx]
    has_cfg = np.isin(idx, configuration_limit_indices(m))
    q_min = np.asarray(m.q_min, dtype=np.float64)
    q_max = np.asarray(m.q_max, dtype=np.float64)
    dq_max = np.full(idx.shape, np.inf)
    dq_min = np.full(idx.shape, np.inf)
    for k, i in enumerate(idx):
        if has_cfg[k]:
            dq_max[k] = q_max[rq + i - rv] - q[rq + i - rv]
            dq_min[k] = q[rq + i - rv] - q_min[rq + i - rv]
    dq_prev = np.zeros(m.nv) if dq_prev is None else np.asarray(dq_prev, dtype=np.float64)
    prev = dq_prev[idx]
    P = np.eye(m.nv)[idx]
    with np.errstate(invalid="ignore"):
        h = np.concatenate([
            np.minimum(a * dt * dt + prev, dt * np.sqrt(2.0 * a * dq_max)),
            np.minimum(a * dt * dt - prev, dt * np.sqrt(2.0 * a * dq_min)),
        ])
    return np.vstack([P, -P]), h
