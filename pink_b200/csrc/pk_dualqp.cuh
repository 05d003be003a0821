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

#ifndef PK_DUALQP_POLISH
#define PK_DUALQP_POLISH 3
#endif

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
    short act[N + 1] = {};   // active constraint ids
    float asg[N + 1] = {};   // sign the normal was added with (equalities may be flipped)
    int iq = 0;
    // membership of the active set: one bit per general/equality row, two words for the box rows
    uint64_t in_hi = 0ull, in_lo = 0ull, in_gen = 0ull;
    // Constraints that cannot enter (dependent on the active set, nothing to release) while
    // violated by less than the final check accepts: opposing rows that pin a direction
    // (lo == hi, a barrier with p_min == p_max) look like that once rounding leaves x 1e-8 on
    // the wrong side of one of them.  Up to four ids, 8 bits each, 0xff = free.
    unsigned skip = 0xffffffffu;
    auto skipped = [&](int id) {
      const unsigned v = (unsigned)id;
      return ((skip & 255u) == v) | (((skip >> 8) & 255u) == v) | (((skip >> 16) & 255u) == v) | ((skip >> 24) == v);
    };
    // tolerance base of constraint id: |rhs| in units of its normalised row, + 1e-3
    auto tol_base = [&](int id, float& scale) {
      float rhs;
      scale = 1.f;
      if (id < P.meq + P.p) {
        const float* row = (id < P.meq) ? P.E[id] : P.G[id - P.meq];
        float nn = 0.f;
        for (int k = 0; k < n; ++k) nn = fmaf(row[k], row[k], nn);
        scale = (nn > 0.f) ? rsqrtf(nn) : 0.f;
        rhs = (id < P.meq) ? P.f[id] : P.h[id - P.meq];
      } else {
        const int bid = id - P.meq - P.p;
        rhs = (bid & 1) ? P.lo[bid >> 1] : P.hi[bid >> 1];
      }
      return fabsf(rhs) * scale + 1e-3f;
    };
    const int max_iter = 4 * (n + P.p + P.meq) + 32;
    int iter = 0;
    // x is carried in fp64: slacks are then exact functions of the fp32 data, so that
    // after a polish (below) violations far below the fp32 resolution of x are seen
    double xd[N];
    for (int i = 0; i < n; ++i) xd[i] = (double)x[i];

    auto slack = [&](int id) -> double {  // s(x) >= 0 form; equalities: E x - f
      if (id < P.meq) {
        double sv = -(double)P.f[id];
        for (int c = 0; c < n; ++c) sv += (double)P.E[id][c] * xd[c];
        return sv;
      }
      if (id < P.meq + P.p) {
        double sv = (double)P.h[id - P.meq];
        for (int c = 0; c < n; ++c) sv -= (double)P.G[id - P.meq][c] * xd[c];
        return sv;
      }
      const int bid = id - P.meq - P.p;
      const int c = bid >> 1;
      return (bid & 1) ? xd[c] - (double)P.lo[c] : (double)P.hi[c] - xd[c];
    };
    auto set_member = [&](int id, bool on) {
      uint64_t* word;
      int bit;
      if (id < P.meq + P.p) { word = &in_gen; bit = id; }
      else {
        const int bid = id - P.meq - P.p;
        word = (bid & 1) ? &in_lo : &in_hi;
        bit = bid >> 1;
      }
      if (on) *word |= (1ull << bit); else *word &= ~(1ull << bit);
    };
    // remove active constraint l: column l of Ra goes, the Hessenberg part below it is
    // rotated back to triangular (same rotations on the columns of J)
    auto drop = [&](int l) {
      set_member(act[l], false);
      for (int i = 0; i < l; ++i)
        for (int k = l; k < iq - 1; ++k) Ra[ut(i, k)] = Ra[ut(i, k + 1)];
      for (int j = l; j < iq - 1; ++j) {
        const float a = Ra[ut(j, j + 1)], bb = Ra[ut(j + 1, j + 1)];
        const float hh = hypotf(a, bb);
        const float c = (hh > 0.f) ? a / hh : 1.f, sn = (hh > 0.f) ? bb / hh : 0.f;
        for (int k = j; k < iq - 1; ++k) {
          const float ra = Ra[ut(j, k + 1)], rb = Ra[ut(j + 1, k + 1)];
          Ra[ut(j, k)] = fmaf(c, ra, sn * rb);
          Ra[ut(j + 1, k + 1)] = fmaf(-sn, ra, c * rb);
        }
        for (int k = 0; k < n; ++k) {
          const float ja = J[k][j], jb = J[k][j + 1];
          J[k][j] = fmaf(c, ja, sn * jb);
          J[k][j + 1] = fmaf(-sn, ja, c * jb);
        }
      }
      for (int c = l; c < iq - 1; ++c) { act[c] = act[c + 1]; asg[c] = asg[c + 1]; u[c] = u[c + 1]; }
      u[iq - 1] = u[iq];
      --iq;
    };
    // Projected Newton steps on the current active manifold with J as the (fp32-accurate)
    // inverse-Hessian factor and everything accumulated in fp64; refreshes the multipliers
    // from the gradient, u = Ra^-1 J1^T grad f.  Along weakly curved directions of the
    // reduced Hessian (posture cost 0.1 next to CoM cost 200) fp32 steps leave errors
    // far above the parity tolerance; this removes them.
    auto polish = [&]() {
      double w[N], gd[N], rho[KA], ud[N + 1];
      for (int k = 0; k < iq; ++k) ud[k] = (double)u[k];
      for (int pass = 0; pass < PK_DUALQP_POLISH; ++pass) {
        // back onto the active constraints: N_a^T delta = -s, delta = J1 w, Ra^T w = -s
        for (int k = 0; k < iq; ++k) {
          double sv = slack(act[k]);
          if (act[k] < P.meq) sv *= (double)asg[k];
          double ww = -sv;
          for (int i = 0; i < k; ++i) ww -= (double)Ra[ut(i, k)] * w[i];
          w[k] = ww / (double)Ra[ut(k, k)];
        }
        for (int i = 0; i < n; ++i) {
          double sacc = xd[i];
          for (int k = 0; k < iq; ++k) sacc += (double)J[i][k] * w[k];
          xd[i] = sacc;
        }
        // stationarity residual  grad f(x) - N_a u  (normals = gradients of s_k >= 0)
        for (int rr = 0; rr < K; ++rr) {
          double sacc = (double)P.b[rr];
          for (int j = 0; j < n; ++j) sacc += (double)P.A[rr][j] * xd[j];
          rho[rr] = sacc;
        }
        for (int i = 0; i < n; ++i) {
          double g = (double)P.d[i] * ((double)P.d[i] * xd[i] + (double)P.beta[i]);
          for (int rr = 0; rr < K; ++rr) g += (double)P.A[rr][i] * rho[rr];
          gd[i] = g;
        }
        for (int k = 0; k < iq; ++k) {
          const int id = act[k];
          if (id < P.meq) {
            for (int c = 0; c < n; ++c) gd[c] -= ud[k] * (double)asg[k] * (double)P.E[id][c];
          } else if (id < P.meq + P.p) {
            for (int c = 0; c < n; ++c) gd[c] += ud[k] * (double)P.G[id - P.meq][c];
          } else {
            const int bid = id - P.meq - P.p;
            gd[bid >> 1] -= (bid & 1) ? ud[k] : -ud[k];
          }
        }
        // w <- J^T residual: the first iq entries correct the multipliers (Ra du = w1),
        // the rest is the null-space Newton step.  Working on the residual (not on the
        // gradient) keeps the fixed point exact when J2 has drifted off N_a by fp32
        // rounding: multipliers of 1e3..1e4 would otherwise leak into the step.
        for (int k = 0; k < n; ++k) {
          double sacc = 0.0;
          for (int i = 0; i < n; ++i) sacc += (double)J[i][k] * gd[i];
          w[k] = sacc;
        }
        for (int i = 0; i < n; ++i) {
          double sacc = xd[i];
          for (int k = iq; k < n; ++k) sacc -= (double)J[i][k] * w[k];
          xd[i] = sacc;
        }
        for (int k = iq - 1; k >= 0; --k) {
          double sacc = w[k];
          for (int c = k + 1; c < iq; ++c) sacc -= (double)Ra[ut(k, c)] * w[c];
          w[k] = sacc / (double)Ra[ut(k, k)];
          ud[k] += w[k];
        }
      }
      for (int k = 0; k < iq; ++k) u[k] = (float)ud[k];
      for (int i = 0; i < n; ++i) x[i] = (float)xd[i];
    };

    // Rounds: dual iteration in fp32 -> polish -> drop wrong-signed multipliers -> look
    // again for violations, now resolved far below fp32; done when a polish leaves nothing.
    float vtol = 1e-6f;
    for (int round = 0; round < 6 && !(status & (PK_STATUS_NO_SOLUTION | PK_STATUS_ITER_LIMIT)); ++round) {
      int changed = 0;
      for (;; ++iter) {
        if (iter >= max_iter) { status |= PK_STATUS_ITER_LIMIT; break; }
        // step 1: most violated constraint (normalised by the row norm, as quadprog)
        int ip = -1;
        float worst = 0.f, sgn = 1.f;
        for (int id = 0; id < m; ++id) {
          float scale, rhs;
          if (skipped(id)) continue;
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
          float s = (float)slack(id);
          float sg = 1.f;
          if (id < P.meq) { sg = (s > 0.f) ? -1.f : 1.f; s = -fabsf(s); }
          s *= scale;
          const float tol = vtol * (fabsf(rhs) * scale + 1e-3f);
          if (s < -tol && s < worst) { worst = s; ip = id; sgn = sg; }
        }
        if (status & PK_STATUS_NO_SOLUTION) break;
        if (ip < 0) break;
        ++changed;
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
            float sacc = 0.f;
            for (int k = iq; k < n; ++k) sacc = fmaf(J[i][k], dv[k], sacc);
            z[i] = sacc;
          }
          for (int i = iq - 1; i >= 0; --i) {
            float sacc = dv[i];
            for (int k = i + 1; k < iq; ++k) sacc = fmaf(-Ra[ut(i, k)], r[k], sacc);
            r[i] = sacc / Ra[ut(i, i)];
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
            float sp = (float)slack(ip);
            if (ip < P.meq) sp *= sgn;
            if (zn > 0.f) t2 = fmaxf(-sp, 0.f) / zn;
          }
          const float t = fminf(t1, t2);
#ifdef PK_DUALQP_TRACE
          printf("round %d iter %d ip %d sgn %g iq %d dd %g d2 %g dep %d t1 %g (l %d) t2 %g viol %g\n", round, iter, ip, sgn,
                 iq, dd, d2, (int)dependent, t1, l, t2, worst);
#endif
          if (!(t < 3.0e38f)) {
            float scale;
            const float base = tol_base(ip, scale);
            float viol = (float)slack(ip);
            viol = ((ip < P.meq) ? fabsf(viol) : -viol) * scale;
            if (viol <= 1e-4f * base && (skip >> 24) == 255u) {
              skip = (skip << 8) | (unsigned)ip;  // tolerated: state untouched, look elsewhere
              added = true;
              break;
            }
            status |= PK_STATUS_NO_SOLUTION;
            break;
          }
          if (t2 < 3.0e38f)
            for (int k = 0; k < n; ++k) xd[k] += (double)t * (double)z[k];
          for (int k = 0; k < iq; ++k) u[k] = fmaf(-t, r[k], u[k]);
          u[iq] += t;
          if (t2 <= t1) {
            // full step: the constraint enters.  Givens rotations on columns iq.. of J zero d[iq+1..]
            for (int j = n - 1; j > iq; --j) {
              const float a = dv[j - 1], bb = dv[j];
              if (bb == 0.f) continue;
              const float hh = hypotf(a, bb);  // no underflow of the squares: entries of d can be ~1e-25
              if (!(hh > 0.f)) continue;
              const float c = a / hh, sn = bb / hh;
              dv[j - 1] = hh;
              dv[j] = 0.f;
              for (int k = 0; k < n; ++k) {
                const float ja = J[k][j - 1], jb = J[k][j];
                J[k][j - 1] = fmaf(c, ja, sn * jb);
                J[k][j] = fmaf(-sn, ja, c * jb);
              }
            }
            for (int k = 0; k <= iq; ++k) Ra[ut(k, iq)] = dv[k];
            act[iq] = (short)ip;
            asg[iq] = sgn;
            set_member(ip, true);
            ++iq;
            added = true;
            break;
          }
          drop(l);  // partial step: constraint l leaves the active set
        }
        if (status & PK_STATUS_NO_SOLUTION) break;
        if (!added) { status |= PK_STATUS_ITER_LIMIT; break; }
      }
      if (status & (PK_STATUS_NO_SOLUTION | PK_STATUS_ITER_LIMIT)) break;
      if (round > 0 && !changed) break;  // a polished point without violations: done
      // polish; release inequality multipliers that the accurate gradient shows negative
      for (int rel = 0; rel <= n; ++rel) {
        polish();
        float umax = 0.f, umin = 0.f;
        int l = -1;
        for (int k = 0; k < iq; ++k) {
          umax = fmaxf(umax, fabsf(u[k]));
          if (act[k] >= P.meq && u[k] < umin) { umin = u[k]; l = k; }
        }
        if (l < 0 || umin >= -1e-6f * umax) break;
        u[iq] = 0.f;
        drop(l);
        ++changed;
      }
      vtol = 1e-9f;
    }
    if (status & PK_STATUS_NO_SOLUTION) return status;
    if (!(status & PK_STATUS_ITER_LIMIT)) {
      // the rounds may run out on inconsistent rows that fp32 keeps "almost consistent":
      // a point that still violates a constraint (active or not) is not a solution
      for (int id = 0; id < m; ++id) {
        float scale = 1.f, rhs;
        if (id < P.meq + P.p) {
          const float* row = (id < P.meq) ? P.E[id] : P.G[id - P.meq];
          float nn = 0.f;
          for (int k = 0; k < n; ++k) nn = fmaf(row[k], row[k], nn);
          if (!(nn > 0.f)) continue;
          scale = rsqrtf(nn);
          rhs = (id < P.meq) ? P.f[id] : P.h[id - P.meq];
        } else {
          const int bid = id - P.meq - P.p;
          rhs = (bid & 1) ? P.lo[bid >> 1] : P.hi[bid >> 1];
          if (!(fabsf(rhs) < 3.0e38f)) continue;
        }
        float sv = (float)slack(id);
        if (id < P.meq) sv = -fabsf(sv);
        if (sv * scale < -1e-4f * (fabsf(rhs) * scale + 1e-3f)) return status | PK_STATUS_NO_SOLUTION;
      }
    }
    // Inconsistent rows that fp32 data rounding has turned "consistent" meet at infinity:
    // a displacement beyond 1e3 (rad or m, per step) on an unbounded coordinate (floating
    // base) is reported as no solution, like the exactly inconsistent case; also traps NaN.
    {
      bool sane = true;  // (fmaxf would drop a NaN)
      for (int i = 0; i < n; ++i) sane = sane && (fabsf(x[i]) < 1e3f);
      if (!sane) return status | PK_STATUS_NO_SOLUTION;
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
