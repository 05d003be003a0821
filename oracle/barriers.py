"""Control barrier functions of the oracle (fp64 numpy, one instance at a time).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

A barrier is a plain dict:

``{"type": "position", "frame": f, "indices": [0, 1, 2], "p_min": [..] | None,
   "p_max": [..] | None, "gain": g, "safe_displacement_gain": r}``
``{"type": "body_spherical", "frames": (f1, f2), "d_min": d, "gain": g,
   "safe_displacement_gain": r}``
``{"type": "self_collision", "pairs": [(fa, fb, ra, rb), ...], "n_pairs": dim,
   "d_min": d, "gain": g, "safe_displacement_gain": r}``

where a collision "pair" is two spheres given by the frames of their centres
and their radii (the sphere--sphere restriction of SURVEY.md section 2; the
reference evaluates arbitrary geometry through coal).
"""

import numpy as np

from . import kinematics as kin


def _gains(barrier, dim):
    g = barrier.get("gain", 1.0)
    g = np.asarray(g, dtype=np.float64)
    if g.ndim == 0:
        return np.full(dim, float(g))
    if g.shape[0] != dim:
        g = np.tile(g, 2)  # pink/barriers/position_barrier.py:84-85
    return g


def _world_position_jacobian(m, fk, f):
    """``R_f J_f[:3]`` (``pink/barriers/position_barrier.py:139-146``)."""
    Rf, _ = kin.frame_placement(m, fk, f)
    return Rf @ kin.frame_jacobian_local(m, fk, f)[..., 0:3, :]


def sphere_pair_distances(m, fk, pairs):
    """Signed distance, nearest points and centres of every sphere pair."""
    out = []
    for fa, fb, ra, rb in pairs:
        _, ca = kin.frame_placement(m, fk, int(fa))
        _, cb = kin.frame_placement(m, fk, int(fb))
        gap = np.linalg.norm(ca - cb)
        u = (ca - cb) / gap if gap > 0 else np.zeros(3)
        out.append((gap - ra - rb, ca - ra * u, cb + rb * u))
    return out


def _closest(distances, dim):
    # pink/barriers/self_collision_barrier.py:123-126
    d = np.asarray(distances)
    return np.argpartition(-d, -dim)[-dim:] if dim > 0 else np.zeros(0, dtype=int)


def barrier_value(m, q, fk, barrier):
    """``Barrier.compute_barrier``: ``h(q)``."""
    t = barrier["type"]
    if t == "position":
        _, p = kin.frame_placement(m, fk, barrier["frame"])
        idx = list(barrier.get("indices") or [0, 1, 2])
        parts = []
        if barrier.get("p_min") is not None:
            parts.append(p[idx] - np.asarray(barrier["p_min"], dtype=np.float64))
        if barrier.get("p_max") is not None:
            parts.append(np.asarray(barrier["p_max"], dtype=np.float64) - p[idx])
        return np.concatenate(parts)
    if t == "body_spherical":
        # pink/barriers/body_spherical_barrier.py:95-104
        _, p1 = kin.frame_placement(m, fk, barrier["frames"][0])
        _, p2 = kin.frame_placement(m, fk, barrier["frames"][1])
        return np.array([(p1 - p2) @ (p1 - p2) - barrier["d_min"] ** 2])
    if t == "self_collision":
        # pink/barriers/self_collision_barrier.py:108-127
        dist = np.array([d for d, _, _ in sphere_pair_distances(m, fk, barrier["pairs"])]) - barrier["d_min"]
        return dist[_closest(dist, barrier["n_pairs"])]
    raise ValueError(t)


def barrier_jacobian(m, q, fk, barrier):
    """``Barrier.compute_jacobian``: ``dh/dq`` as ``(dim, nv)``."""
    t = barrier["type"]
    if t == "position":
        idx = list(barrier.get("indices") or [0, 1, 2])
        Jw = _world_position_jacobian(m, fk, barrier["frame"])[idx]
        parts = []
        if barrier.get("p_min") is not None:
            parts.append(Jw.copy())
        if barrier.get("p_max") is not None:
            parts.append(-Jw.copy())
        return np.vstack(parts)
    if t == "body_spherical":
        # pink/barriers/body_spherical_barrier.py:133-137
        _, p1 = kin.frame_placement(m, fk, barrier["frames"][0])
        _, p2 = kin.frame_placement(m, fk, barrier["frames"][1])
        J1 = _world_position_jacobian(m, fk, barrier["frames"][0])
        J2 = _world_position_jacobian(m, fk, barrier["frames"][1])
        return (2.0 * (p1 - p2) @ (J1 - J2))[None, :]
    if t == "self_collision":
        # pink/barriers/self_collision_barrier.py:169-224
        res = sphere_pair_distances(m, fk, barrier["pairs"])
        dist = np.array([d for d, _, _ in res])
        J = np.zeros((barrier["n_pairs"], m.nv))
        for row, k in enumerate(_closest(dist, barrier["n_pairs"])):
            fa, fb, _, _ = barrier["pairs"][int(k)]
            _, w1, w2 = res[int(k)]
            if np.allclose(w1, w2):
                continue
            n = (w1 - w2) / np.linalg.norm(w1 - w2)
            J1 = kin.point_jacobian_world(m, fk, int(m.frame_body[int(fa)]), w1)
            J2 = kin.point_jacobian_world(m, fk, int(m.frame_body[int(fb)]), w2)
            J[row] = n @ J1 - n @ J2
        return np.nan_to_num(J)
    raise ValueError(t)


def gain_function(barrier, h):
    """Class-K function of the barrier (identity unless the class overrides it:
    ``pink/barriers/body_spherical_barrier.py:65``)."""
    if barrier["type"] == "body_spherical":
        return h / (1.0 + np.abs(h))
    return h


def barrier_qp_objective(m, q, fk, barrier):
    """``Barrier.compute_qp_objective`` (``pink/barriers/barrier.py:151-204``) with
    the default zero safe displacement."""
    H = np.zeros((m.nv, m.nv))
    c = np.zeros(m.nv)
    r = barrier.get("safe_displacement_gain", 0.0)
    if r > 1e-6:
        J = barrier_jacobian(m, q, fk, barrier)
        H += r / np.linalg.norm(J) ** 2 * np.eye(m.nv)
    return H, c


def barrier_qp_inequalities(m, q, fk, barrier, dt):
    """``Barrier.compute_qp_inequalities`` (``pink/barriers/barrier.py:206-254``)."""
    J = barrier_jacobian(m, q, fk, barrier)
    h = barrier_value(m, q, fk, barrier)
    g = _gains(barrier, h.shape[0])
    return -J / dt, g * gain_function(barrier, h)
