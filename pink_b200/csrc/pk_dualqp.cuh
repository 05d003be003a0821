// Dual active-set QP for IK steps with GENERAL inequality rows and equalities.
//
// pink.solve_ik stacks, next to the +-e_i rows of ConfigurationLimit /
// VelocityLimit / AccelerationLimit, dense rows from barriers
// (pink/barriers/barrier.py:246-252: G = -J_h / dt, h = gain * alpha(h(q))), from
// FloatingBaseVelocityLimit (pink/limits/floating_base_velocity_limit.py:118-148)
// and equality rows from `constraints=` (pink/solve_ik.py:125-149), and hands
//
//     minimise 1/2 x^T H x + c^T x   s.t.  G x <= h,  E x = f
//
// to quadprog, i.e. to the Goldfarb-Idnani dual method.  This file is the same
// method on the square-root form of the objective (see pk_lsq.cuh):
//
//     H = R^T R  with R from the Householder QR of [diag(d); A]   (never H itself)
//     J = R^-1   (the L^-T of Goldfarb-Idnani), kept as a dense n x n matrix and
//                rotated by Givens when constraints enter / leave the active set
//
// so that fp32 sees cond(R) = sqrt(cond(H)).  Box rows are handled as constraints
// with unit normals (d = J^T n is a row of J).  After the dual iteration has found
// the active set, x is polished with one projected Newton step that uses the
// factored gradient A^T (A x + b) + d (d x + beta):
//
//     x += J1 R_a^-T (rhs_a - N_a^T x)       (back onto the active constraints)
//     x -= J2 J2^T grad f(x)                 (minimiser on that manifold)
//
// which removes the rounding accumulated over the dual steps.  A dual method needs
// no feasible start, detects infeasibility (PK_STATUS_NO_SOLUTION, where the
// reference raises NoSolutionFound) and, the QP being strictly convex, returns the
// same unique minimiser as quadprog.
#pragma once

#include "pk_math.cuh"

#include "../../include/pink_b200.h"

namespace pk {

template <int KMAX, int N, int MG, int ME>
struct DualQP {
  static constexpr int KA = KMAX > 0 ? KMAX : 1;
  static constexpr int GA = MG > 0 ? MG : 1;
  static constexpr int EA = ME > 0 ? ME : 1;
  static constexpr int NT = N * (N + 1) / 2;

  PK_HD static constexpr int ut(int k, int j) { return k * N - k * (k - 1) / 2 + (j - k); }  // j >= k, packed upper

  // constraint ids: [0, meq) equalities, [meq, meq + p) general rows, then
  // 2 i (upper bound of x_i), 2 i + 1 (lower bound of x_i)
  struct Problem {
    const float (*A)[N];
    const float* b;
    const float* d;
    const float* beta;
    const float* lo;
    const float* hi;
    const float (*G)[N];
    const float* h;
    const float (*E)[N];
    const float* f;
    int K, n, p, meq;
  };

  // value of constraint `id` at x, in the form  s(x) >= 0  (for equalities: E x - f)
  static PK_HD float value(const Problem& P, int id, const float* x) {
    if (id < P.meq) {
      float s = -P.f[id];
      for (int k = 0; k < P.n; ++k) s = fmaf(P.E[id][k], x[k], s);
      return s;
    }
    id -= P.meq;
    if (id < P.p) {
      float s = P.h[id];
      for (int k = 0; k < P.n; ++k) s = fmaf(-P.G[id][k], x[k], s);
      return s;
    }
    id -= P.p;
    const int i = id >> 1;
    return (id & 1) ? x[i] - P.lo[i] : P.hi[i] - x[i];
  }

  // d = J^T n for the normal n of constraint `id` (gradient of s), times `sgn`
  static PK_HD void normal_times_J(const Problem& P, const float (&J)[N][N], int id, float sgn, float* dvec) {
    const int n = P.n;
    if (id < P.meq + P.p) {
      const float* row = (id < P.meq) ? P.E[id] : P.G[id - P.meq];
      const float s2 = (id < P.meq) ? sgn : -sgn;
      for (int i = 0; i < n; ++i) {
        float s = 0.f;
        for (int k = 0; k < n; ++k) s = fmaf(J[k][i], row[k], s);
        dvec[i] = s2 * s;
      }
      return;
    }
    id -= P.meq + P.p;
    const int c = id >> 1;
    const float s2 = (id & 1) ? sgn : -sgn;
    for (int i = 0; i < n; ++i) dvec[i] = s2 * J[c][i];
  }

  // z . n for the normal of `id` (times sgn)
  static PK_HD float normal_dot(const Problem& P, int id, float sgn, const float* z) {
    if (id < P.meq + P.p) {
      const float* row = (id < P.meq) ? P.E[id] : P.G[id - P.meq];
      float s = 0.f;
      for (int k = 0; k < P.n; ++k) s = fmaf(row[k], z[k], s);
      return (id < P.meq) ? sgn * s : -sgn * s;
    }
    id -= P.meq + P.p;
    return ((id & 1) ? sgn : -sgn) * z[id >> 1];
  }

  static PK_HD int run(const Problem& P, float (&x)[N]) {
    const int n = P.n, K = P.K;
    int status = 0;
    // ---- R from the Householder QR of [diag(d); A], x = unconstrained minimiser ----
    float J[N][N];
    {
      float Aw[KA][N];
      float zb[KA], zt[N], Ru[NT];
      for (int r = 0; r < K; ++r) {
        zb[r] = P.b[r];
        for (int c = 0; c < n; ++c) Aw[r][c] = P.A[r][c];
      }
      for (int k = 0; k < n; ++k) {
        float sigma = 0.f;
        for (int r = 0; r < K; ++r) sigma = fmaf(Aw[r][k], Aw[r][k], sigma);
        const float alpha = P.d[k];
        zt[k] = P.beta[k];
        const float norm = sqrtf(fmaf(alpha, alpha, sigma));
        if (!(norm > 0.f)) status |= PK_STATUS_NOT_POSDEF;
        const float v0 = alpha + norm;
        const float tau = (sigma > 0.f) ? 1.f / (norm * v0) : 0.f;
        Ru[ut(k, k)] = (sigma > 0.f) ? -norm : alpha;
        for (int j = k + 1; j < n; ++j) {
          float s = 0.f;
          for (int r = 0; r < K; ++r) s = fmaf(Aw[r][k], Aw[r][j], s);
          s *= tau;
          Ru[ut(k, j)] = -s * v0;
          for (int r = 0; r < K; ++r) Aw[r][j] = fmaf(-s, Aw[r][k], Aw[r][j]);
        }
        float s = v0 * zt[k];
        for (int r = 0; r < K; ++r) s = fmaf(Aw[r][k], zb[r], s);
        s *= tau;
        zt[k] = fmaf(-s, v0, zt[k]);
        for (int r = 0; r < K; ++r) zb[r] = fmaf(-s, Aw[r][k], zb[r]);
      }
      if (status) {
        for (int i = 0; i < n; ++i) x[i] = 0.f;
        return status;
      }
      for (int kk = 0; kk < n; ++kk) {
        const int k = n - 1 - kk;
        float s = -zt[k];
        for (int j = k + 1; j < n; ++j) s = fmaf(-Ru[ut(k, j)], x[j], s);
        x[k] = s / Ru[ut(k, k)];
      }
      // J = R^-1 (upper triangular), column by column
      for (int c = 0; c < n; ++c) {
        for (int i = n - 1; i >= 0; --i) {
          if (i > c) { J[i][c] = 0.f; continue; }
          float s = (i == c) ? 1.f : 0.f;
          for (int k = i + 1; k <= c; ++k) s = fmaf(-Ru[ut(i, k)], J[k][c], s);
          J[i][c] = s / Ru[ut(i, i)];
        }
      }
    }

    // ---- Goldfarb-Idnani iteration ----
    const int m = P.meq + P.p + 2 * n;
    float Ra[NT];            // triangular factor of the active normals (J^T N = [Ra; 0])
    float u[N + 1];          // multipliers of the active constraints (+ the entering one)
    float dv[N], z[N], r[N];
    short act[N + 1];        // active constraint ids
    float asg[N + 1];        // sign the normal was added with (equalities may be flipped)
    int iq = 0;
    // membership of the active set: one bit per general/equality row, two words for the box rows
    uint64_t in_hi = 0ull, in_lo = 0ull, in_gen = 0ull;
    const int max_iter = 4 * (n + P.p + P.meq) + 32;
    int iter = 0;
    for (;; ++iter) {
      if (iter >= max_iter) { status |= PK_STATUS_ITER_LIMIT; break; }
      // step 1: most violated constraint (normalised by the row norm, as quadprog)
      int ip = -1;
      float worst = 0.f, sgn = 1.f;
      for (int id = 0; id < m; ++id) {
        float scale, rhs;
        if (id < P.meq + P.p) {
          if ((in_gen >> id) & 1ull) continue;
          const float* row = (id < P.meq) ? P.E[id] : P.G[id - P.meq];
          float nn = 0.f;
          for (int k = 0; k < n; ++k) nn = fmaf(row[k], row[k], nn);
          if (!(nn > 0.f)) {
            // empty row: 0 <= h (or 0 = f) either holds or never will
            const float s0 = (id < P.meq) ? -fabsf(P.f[id]) : P.h[id - P.meq];
            if (s0 < 0.f) status |= PK_STATUS_NO_SOLUTION;
            continue;
          }
          scale = rsqrtf(nn);
          rhs = (id < P.meq) ? P.f[id] : P.h[id - P.meq];
        } else {
          const int bid = id - P.meq - P.p;
          const int i = bid >> 1;
          if ((bid & 1) ? ((in_lo >> i) & 1ull) : ((in_hi >> i) & 1ull)) continue;
          rhs = (bid & 1) ? P.lo[i] : P.hi[i];
          if (!(fabsf(rhs) < 3.0e38f)) continue;  // infinite bound: no row
          scale = 1.f;
        }
        float s = value(P, id, x);
        float sg = 1.f;
        if (id < P.meq) { sg = (s > 0.f) ? -1.f : 1.f; s = -fabsf(s); }
        s *= scale;
        const float tol = 1e-6f * fabsf(rhs) * scale + 1e-9f;
        if (s < -tol && s < worst) { worst = s; ip = id; sgn = sg; }
      }
      if (status & PK_STATUS_NO_SOLUTION) break;
      if (ip < 0) break;
      u[iq] = 0.f;
      bool added = false;
      for (int inner = 0; inner <= n + P.p + P.meq + 2 && !added; ++inner) {
        // step 2a: d = J^T n+, z = J2 d2, r = Ra^-1 d1
        normal_times_J(P, J, ip, sgn, dv);
        float dd = 0.f, d2 = 0.f;
        for (int i = 0; i < n; ++i) {
          dd = fmaf(dv[i], dv[i], dd);
          if (i >= iq) d2 = fmaf(dv[i], dv[i], d2);
        }
        const bool dependent = !(d2 > 1e-10f * dd);
        for (int i = 0; i < n; ++i) {
          float s = 0.f;
          for (int k = iq; k < n; ++k) s = fmaf(J[i][k], dv[k], s);
          z[i] = s;
        }
        for (int i = iq - 1; i >= 0; --i) {
          float s = dv[i];
          for (int k = i + 1; k < iq; ++k) s = fmaf(-Ra[ut(i, k)], r[k], s);
          r[i] = s / Ra[ut(i, i)];
        }
        // step 2b: step lengths
        float t1 = 3.0e38f;
        int l = -1;
        for (int k = 0; k < iq; ++k)
          if (act[k] >= P.meq && r[k] > 0.f) {
            const float t = fmaxf(u[k], 0.f) / r[k];
            if (t < t1) { t1 = t; l = k; }
          }
        float t2 = 3.0e38f;
        if (!dependent) {
          const float zn = normal_dot(P, ip, sgn, z);
          float sp = value(P, ip, x);
          if (ip < P.meq) sp *= sgn;
          if (zn > 0.f) t2 = fmaxf(-sp, 0.f) / zn;
        }
        const float t = fminf(t1, t2);
        if (!(t < 3.0e38f)) { status |= PK_STATUS_NO_SOLUTION; break; }
        if (t2 < 3.0e38f)
          for (int k = 0; k < n; ++k) x[k] = fmaf(t, z[k], x[k]);
        for (int k = 0; k < iq; ++k) u[k] = fmaf(-t, r[k], u[k]);
        u[iq] += t;
        if (t2 <= t1) {
          // full step: the constraint enters.  Givens rotations on columns iq.. of J zero d[iq+1..]
          for (int j = n - 1; j > iq; --j) {
            const float a = dv[j - 1], bb = dv[j];
            if (bb == 0.f) continue;
            const float hh = sqrtf(fmaf(a, a, bb * bb));
            const float c = a / hh, s = bb / hh;
            dv[j - 1] = hh;
            dv[j] = 0.f;
            for (int k = 0; k < n; ++k) {
              const float ja = J[k][j - 1], jb = J[k][j];
              J[k][j - 1] = fmaf(c, ja, s * jb);
              J[k][j] = fmaf(-s, ja, c * jb);
            }
          }
          for (int k = 0; k <= iq; ++k) Ra[ut(k, iq)] = dv[k];
          act[iq] = (short)ip;
          asg[iq] = sgn;
          if (ip < P.meq + P.p) in_gen |= (1ull << ip);
          else {
            const int bid = ip - P.meq - P.p;
            if (bid & 1) in_lo |= (1ull << (bid >> 1)); else in_hi |= (1ull << (bid >> 1));
          }
          ++iq;
          added = true;
          break;
        }
        // partial step: constraint l leaves the active set
        {
          const int idl = act[l];
          if (idl < P.meq + P.p) in_gen &= ~(1ull << idl);
          else {
            const int bid = idl - P.meq - P.p;
            if (bid & 1) in_lo &= ~(1ull << (bid >> 1)); else in_hi &= ~(1ull << (bid >> 1));
          }
        }
        // remove column l of Ra: rows above l shift left, the Hessenberg part below is
        // rotated back to triangular (same rotations on the columns of J)
        for (int i = 0; i < l; ++i)
          for (int k = l; k < iq - 1; ++k) Ra[ut(i, k)] = Ra[ut(i, k + 1)];
        for (int j = l; j < iq - 1; ++j) {
          const float a = Ra[ut(j, j + 1)], bb = Ra[ut(j + 1, j + 1)];
          const float hh = sqrtf(fmaf(a, a, bb * bb));
          const float c = (hh > 0.f) ? a / hh : 1.f, s = (hh > 0.f) ? bb / hh : 0.f;
          for (int k = j; k < iq - 1; ++k) {
            const float ra = Ra[ut(j, k + 1)], rb = Ra[ut(j + 1, k + 1)];
            Ra[ut(j, k)] = fmaf(c, ra, s * rb);
            Ra[ut(j + 1, k + 1)] = fmaf(-s, ra, c * rb);
          }
          for (int k = 0; k < n; ++k) {
            const float ja = J[k][j], jb = J[k][j + 1];
            J[k][j] = fmaf(c, ja, s * jb);
            J[k][j + 1] = fmaf(-s, ja, c * jb);
          }
        }
        for (int c = l; c < iq - 1; ++c) { act[c] = act[c + 1]; asg[c] = asg[c + 1]; u[c] = u[c + 1]; }
        u[iq - 1] = u[iq];
        --iq;
      }
      if (status & PK_STATUS_NO_SOLUTION) break;
      if (!added) { status |= PK_STATUS_ITER_LIMIT; break; }
    }
    if (status & PK_STATUS_NO_SOLUTION) return status;

    // ---- polish on the final active set ----
    for (int pass = 0; pass < 2 && iq < n + 1; ++pass) {
      // back onto the active constraints: N_a^T delta = -rho, delta = J1 w, Ra^T w = -rho
      for (int k = 0; k < iq; ++k) {
        float s = value(P, act[k], x);
        if (act[k] < P.meq) s *= asg[k];
        float w = -s;
        for (int i = 0; i < k; ++i) w = fmaf(-Ra[ut(i, k)], r[i], w);
        r[k] = w / Ra[ut(k, k)];
      }
      for (int i = 0; i < n; ++i) {
        float s = x[i];
        for (int k = 0; k < iq; ++k) s = fmaf(J[i][k], r[k], s);
        x[i] = s;
      }
      if (iq >= n) break;
      // projected Newton step with the factored gradient
      float rho[KA];
      for (int rr = 0; rr < K; ++rr) {
        float s = P.b[rr];
        for (int j = 0; j < n; ++j) s = fmaf(P.A[rr][j], x[j], s);
        rho[rr] = s;
      }
      for (int i = 0; i < n; ++i) {
        float g = P.d[i] * fmaf(P.d[i], x[i], P.beta[i]);
        for (int rr = 0; rr < K; ++rr) g = fmaf(P.A[rr][i], rho[rr], g);
        z[i] = g;
      }
      for (int k = iq; k < n; ++k) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s = fmaf(J[i][k], z[i], s);
        dv[k] = s;
      }
      for (int i = 0; i < n; ++i) {
        float s = x[i];
        for (int k = iq; k < n; ++k) s = fmaf(-J[i][k], dv[k], s);
        x[i] = s;
      }
    }
    // coordinates on a bound sit exactly on it
    for (int i = 0; i < n; ++i) {
      if ((in_hi >> i) & 1ull) x[i] = P.hi[i];
      else if ((in_lo >> i) & 1ull) x[i] = P.lo[i];
    }
    return status;
  }
};

}  // namespace pk
