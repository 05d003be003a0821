"""Dense strictly-convex QP of the oracle: Goldfarb-Idnani dual active set.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

This is the algorithm ``qpsolvers.solve_problem(problem, solver="quadprog")``
runs at ``pink/solve_ik.py:270`` (quadprog = Goldfarb & Idnani, "A numerically
stable dual method for solving strictly convex quadratic programs", Math.
Prog. 27, 1983).  quadprog itself is a third-party C library absent from
``/root/reference`` and from this image, so the published algorithm is
restated here in fp64 with explicit projectors recomputed by QR at every
change of the active set (no Givens up/down-dating: clarity over speed).

    minimise  1/2 x^T P x + q^T x   s.t.  A x = b,  G x <= h

Because ``P`` is positive definite the minimiser is unique, so any exact
method must agree with quadprog up to rounding;
``tests/test_oracle_qp.py`` cross-checks against brute-force enumeration of
active sets and against ``scipy.optimize.minimize(method="SLSQP")``.
"""

import numpy as np


class QPResult:
    """Solution record, shaped like ``qpsolvers.Solution`` where it matters
    (``.x``, ``.found``; ``pink/solve_ik.py:271-273``)."""

    def __init__(self, x, found, z=None, y=None, active=None, iterations=0):
        self.x = x
        self.found = found
        self.z = z  # multipliers of G x <= h (>= 0)
        self.y = y  # multipliers of A x = b
        self.active = active if active is not None else []
        self.iterations = iterations


def _projectors(Linv, N):
    """``Q1, R`` of ``L^-1 N`` (columns of N are the active normals)."""
    if N.shape[1] == 0:
        return np.zeros((Linv.shape[0], 0)), np.zeros((0, 0))
    Q1, R = np.linalg.qr(Linv @ N)
    return Q1, R


def solve_qp(P, q, G=None, h=None, A=None, b=None, max_iter=None):
    """Goldfarb-Idnani dual active-set method.

    Constraint normals follow the paper's convention ``n^T x >= b``: an
    inequality row ``G_i x <= h_i`` has normal ``n_i = -G_i`` and slack
    ``s_i(x) = h_i - G_i x``; an equality ``A_i x = b_i`` has normal ``A_i``
    and is activated first without a sign restriction on its multiplier.
    """
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    n = q.shape[0]
    G = np.zeros((0, n)) if G is None else np.asarray(G, dtype=np.float64).reshape(-1, n)
    h = np.zeros(0) if h is None else np.asarray(h, dtype=np.float64).reshape(-1)
    A = np.zeros((0, n)) if A is None else np.asarray(A, dtype=np.float64).reshape(-1, n)
    b = np.zeros(0) if b is None else np.asarray(b, dtype=np.float64).reshape(-1)
    meq, m = A.shape[0], G.shape[0]
    if max_iter is None:
        max_iter = 10 * (m + meq + n) + 50

    try:
        L = np.linalg.cholesky(P)
    except np.linalg.LinAlgError:
        return QPResult(None, False)
    Linv = np.linalg.inv(L)
    x = -Linv.T @ (Linv @ q)  # unconstrained minimiser

    # all constraints in "n^T x - b0 >= 0" form; equalities first
    normals = np.vstack([A, -G])  # rows are n_i^T
    offsets = np.concatenate([b, -h])  # n_i^T x >= offsets_i
    norms = np.linalg.norm(normals, axis=1)
    norms = np.where(norms > 0.0, norms, 1.0)

    active = []  # indices into `normals`, in order of activation
    u = np.zeros(0)  # multipliers of the active constraints
    iterations = 0

    def slack(i):
        return normals[i] @ x - offsets[i]

    next_eq = 0
    while True:
        iterations += 1
        if iterations > max_iter:
            return QPResult(None, False, iterations=iterations)
        # Step 1: choose a violated constraint
        if next_eq < meq:
            p = next_eq
            next_eq += 1
        else:
            inactive = [i for i in range(meq, meq + m) if i not in active]
            if not inactive:
                break
            viol = np.array([slack(i) / norms[i] for i in inactive])
            k = int(np.argmin(viol))
            if viol[k] >= -1e-10:
                break
            p = inactive[k]
        n_plus = normals[p]
        u_plus = 0.0
        is_eq = p < meq

        # Step 2: move until constraint p is satisfied (or proved infeasible)
        while True:
            N = normals[active].T if active else np.zeros((n, 0))
            Q1, R = _projectors(Linv, N)
            d = Linv @ n_plus
            d_perp = d - Q1 @ (Q1.T @ d)
            z = Linv.T @ d_perp  # primal step direction  H n+
            r = np.linalg.solve(R, Q1.T @ d) if active else np.zeros(0)  # N* n+
            zn = float(d_perp @ d_perp)  # = z^T n+ >= 0
            z_is_zero = zn <= 1e-13 * max(float(d @ d), 1e-300)

            s_p = slack(p)
            # sign handling for equalities: step towards s_p = 0 from either side
            sgn = 1.0
            if is_eq and s_p > 0.0:
                sgn = -1.0  # use the normal -n_plus so that the slack is negative
            # (a) largest dual step keeping multipliers of active inequalities >= 0
            t1, l_drop = np.inf, -1
            for k_act, idx in enumerate(active):
                if idx < meq:
                    continue
                rk = sgn * r[k_act]
                if rk > 1e-14:
                    cand = u[k_act] / rk
                    if cand < t1:
                        t1, l_drop = cand, k_act
            # (b) primal step length that makes constraint p active
            t2 = np.inf if z_is_zero else -(sgn * s_p) / zn
            t = min(t1, t2)
            if not np.isfinite(t):
                return QPResult(None, False, iterations=iterations)  # infeasible
            if np.isinf(t2):
                # dual step only, then drop the blocking constraint
                u = u - t * sgn * r
                u_plus += t
                active.pop(l_drop)
                u = np.delete(u, l_drop)
                continue
            x = x + t * sgn * z
            u = u - t * sgn * r
            u_plus += t
            if t2 <= t1:
                active.append(p)
                u = np.append(u, sgn * u_plus if is_eq else u_plus)
                break
            active.pop(l_drop)
            u = np.delete(u, l_drop)

    z_mult = np.zeros(m)
    y_mult = np.zeros(meq)
    for k_act, idx in enumerate(active):
        if idx >= meq:
            z_mult[idx - meq] = u[k_act]
        else:
            y_mult[idx] = -u[k_act]
    return QPResult(
        x,
        True,
        z=z_mult,
        y=y_mult,
        active=[i - meq for i in active if i >= meq],
        iterations=iterations,
    )


def kkt_residuals(P, q, G, h, x, z):
    """``(stationarity, primal violation, dual violation, complementarity)`` of
    an inequality-only solution; all should vanish at the optimum."""
    G = np.zeros((0, len(q))) if G is None else np.asarray(G)
    h = np.zeros(0) if h is None else np.asarray(h)
    stat = np.abs(P @ x + q + G.T @ z).max()
    prim = max(0.0, float((G @ x - h).max())) if len(h) else 0.0
    dual = max(0.0, float((-z).max())) if len(h) else 0.0
    comp = float(np.abs(z * (G @ x - h)).max()) if len(h) else 0.0
    return stat, prim, dual, comp


def solve_qp_bruteforce(P, q, G, h, tol=1e-9):
    """Exact solution by enumerating active sets (tiny problems only)."""
    import itertools

    n = len(q)
    m = 0 if G is None else len(h)
    best = None
    for k in range(0, min(n, m) + 1):
        for S in itertools.combinations(range(m), k):
            S = list(S)
            if k:
                Gs = G[S]
                K = np.block([[P, Gs.T], [Gs, np.zeros((k, k))]])
                rhs = np.concatenate([-q, h[S]])
                try:
                    sol = np.linalg.solve(K, rhs)
                except np.linalg.LinAlgError:
                    continue
                x, lam = sol[:n], sol[n:]
            else:
                x, lam = np.linalg.solve(P, -q), np.zeros(0)
            if m and (G @ x - h).max() > tol:
                continue
            if k and lam.min() < -tol:
                continue
            val = 0.5 * x @ P @ x + q @ x
            if best is None or val < best[0]:
                best = (val, x)
    return None if best is None else best[1]
