// Warp-cooperative dual active-set QP (Goldfarb-Idnani in square-root form) for the
// tree kernel: the method of pk_dualqp.cuh with one instance per warp and the state
// (J = R^-1, the triangular factor Ra of the active normals, multipliers) in the warp's
// slice of shared memory instead of thread-local memory.  Box rows, dense inequality rows
// (barriers: pink/barriers/barrier.py:206-254; floating-base limit) and the equality rows of
// solve_ik(..., constraints=...) (pink/solve_ik.py:125-149).
//
//   * R from TreeStep::eqp with nothing fixed (Householder QR of [diag(d); A]); J = R^-1
//     by columns (lane c owns column c);
//   * a constraint enters with ONE Householder reflection of the columns iq.. of J (row
//     parallel, no dependent chain) instead of the Givens sequence of the scalar code;
//     it leaves with Givens rotations on Ra / J (the Hessenberg part only);
//   * x is carried in fp64 and the active set is polished by KKT-residual refinement in
//     fp64 exactly as in pk_dualqp.cuh (see the comments there), lane-parallel.
//
// Written in the lane-block style of pk_warp.cuh (host-simulable).
#pragma once

#include "pk_tree.cuh"

#ifndef PK_DUALQP_POLISH
#define PK_DUALQP_POLISH 3
#endif

namespace pk {

struct TreeDual {
  // constraint ids: [0, meq) equalities E x = f (solve_ik(..., constraints=...), entered with
  // the sign that makes them violated and never dropped); [meq, pp) dense rows G x <= h;
  // pp + 2 i: x_i <= hi_i; pp + 2 i + 1: x_i >= lo_i
  static PK_HD int solve(float* W, const TreePlan& L) {
    const int n = L.nv, K = L.K, p = L.p, meq = L.meq, pp = L.meq + L.p, ld = L.ldj;
    const float* E = W + L.o_E;
    const float* fe = W + L.o_fe;
    float* en = W + L.o_en;    // 1 / |E_r|
    float* asg = W + L.o_asg;  // sign an active equality was entered with
    const float* A = W + L.o_A;
    const float* bv = W + L.o_b;
    const float* dg = W + L.o_d;
    const float* beta = W + L.o_beta;
    const float* lo = W + L.o_lo;
    const float* hi = W + L.o_hi;
    const float* G = W + L.o_G;
    const float* hg = W + L.o_hg;
    float* gn = W + L.o_gn;  // 1 / |G_r|
    float* x = W + L.o_x;
    float* J = W + L.o_J;
    float* RA = W + L.o_RA;
    float* dv = W + L.o_dv;
    float* z = W + L.o_z;
    float* r = W + L.o_r;
    float* u = W + L.o_u;
    int* act = reinterpret_cast<int*>(W + L.o_act);
    double* xd = reinterpret_cast<double*>(W + L.o_xd);
    double* wd = reinterpret_cast<double*>(W + L.o_wd);
    double* gd = reinterpret_cast<double*>(W + L.o_gd);
    double* rhod = reinterpret_cast<double*>(W + L.o_rhod);
    double* ud = reinterpret_cast<double*>(W + L.o_ud);
    int status = 0;

    // ---- unconstrained minimiser and R ----
    PK_LANES(l) {
      #pragma unroll 1
      for (int i = l; i < n; i += 32) x[i] = 0.f;
    }
    PK_WSYNC();
    if (!TreeStep::eqp(W, L, 0ull)) return PK_STATUS_NOT_POSDEF;
    const float* Rd = W + L.o_rd;
    const float* Ru = W + L.o_ru;
    const float* y = W + L.o_y;
    // J = R^-1, column by column by back-substitution (upper triangular).  The work of
    // column c grows like c^2 / 2, so lane l takes the pair (l, n-1-l): balanced, and no
    // second pass in which the longest columns run alone (n <= 64).
    PK_LANES(l) {
      #pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int c = h == 0 ? l : n - 1 - l;
        if (c < 0 || c >= n || 2 * l > n - 1 || (h == 1 && c == l)) continue;
        #pragma unroll 1
        for (int i = n - 1; i >= 0; --i) {
          float s = 0.f;
          if (i <= c) {
            s = (i == c) ? 1.f : 0.f;
            #pragma unroll 1
            for (int k = i + 1; k <= c; ++k) s = fmaf(-Ru[TreeStep::ru(L, i, k)], J[k * ld + c], s);
            s /= Rd[i];
          }
          J[i * ld + c] = s;
        }
      }
      #pragma unroll 1
      for (int i = l; i < n; i += 32) xd[i] = (double)y[i];
      if (l < p) {
        float nn = 0.f;
        #pragma unroll 1
        for (int k = 0; k < n; ++k) nn = fmaf(G[l * L.lda + k], G[l * L.lda + k], nn);
        gn[l] = (nn > 0.f) ? rsqrtf(nn) : 0.f;
      }
      if (l < meq) {
        float nn = 0.f;
        #pragma unroll 1
        for (int k = 0; k < n; ++k) nn = fmaf(E[l * L.lda + k], E[l * L.lda + k], nn);
        en[l] = (nn > 0.f) ? rsqrtf(nn) : 0.f;
      }
    }
    PK_WSYNC();

    int iq = 0;
    uint64_t in_hi = 0ull, in_lo = 0ull, in_gen = 0ull;
    const int max_iter = 4 * (n + pp) + 32;
    int iter = 0;
    float vtol = 1e-6f;

    // slack of constraint id (s >= 0 form; equalities: E x - f, sign applied by the caller),
    // computed by the whole warp
    auto slack = [&](int id) -> double {
      if (id >= pp) {
        const int c = (id - pp) >> 1;
        return ((id - pp) & 1) ? xd[c] - (double)lo[c] : (double)hi[c] - xd[c];
      }
      const float* row = (id < meq) ? E + id * L.lda : G + (id - meq) * L.lda;
      LaneVar<double> part;
      PK_LANES(l) {
        double s = 0.0;
        #pragma unroll 1
        for (int k = l; k < n; k += 32) s += (double)row[k] * xd[k];
        part[l] = s;
      }
      const double dotv = lane_sum_d(part);
      return (id < meq) ? dotv - (double)fe[id] : (double)hg[id - meq] - dotv;
    };
    auto set_member = [&](int id, bool on) {
      if (id < pp) { if (on) in_gen |= (1ull << id); else in_gen &= ~(1ull << id); return; }
      const int c = (id - pp) >> 1;
      uint64_t& w = ((id - pp) & 1) ? in_lo : in_hi;
      if (on) w |= (1ull << c); else w &= ~(1ull << c);
    };
    // remove active constraint l (Givens on the Hessenberg part of RA and the same columns of J)
    auto drop = [&](int l0) {
      set_member(act[l0], false);
      PK_WSYNC();
      PK_LANES(l) {
        // shift columns l0+1.. of RA one to the left (rows 0..iq-1), lane per row
        #pragma unroll 1
        for (int i = l; i < iq; i += 32) {
          #pragma unroll 1
          for (int k = l0; k < iq - 1; ++k) RA[i * ld + k] = RA[i * ld + k + 1];
          RA[i * ld + iq - 1] = 0.f;
        }
      }
      PK_WSYNC();
      #pragma unroll 1
      for (int j = l0; j < iq - 1; ++j) {
        const float a = RA[j * ld + j], bb = RA[(j + 1) * ld + j];
        const float hh = hypotf(a, bb);
        const float c = (hh > 0.f) ? a / hh : 1.f, sn = (hh > 0.f) ? bb / hh : 0.f;
        PK_WSYNC();
        PK_LANES(l) {
          #pragma unroll 1
          for (int k = j + l; k < iq - 1; k += 32) {
            const float ra = RA[j * ld + k], rb = RA[(j + 1) * ld + k];
            RA[j * ld + k] = fmaf(c, ra, sn * rb);
            RA[(j + 1) * ld + k] = fmaf(-sn, ra, c * rb);
          }
          #pragma unroll 1
          for (int i = l; i < n; i += 32) {
            const float ja = J[i * ld + j], jb = J[i * ld + j + 1];
            J[i * ld + j] = fmaf(c, ja, sn * jb);
            J[i * ld + j + 1] = fmaf(-sn, ja, c * jb);
          }
        }
        PK_WSYNC();
      }
      PK_LANES(l) {
        // act / u shift; done by one lane (short lists)
        if (l == 0) {
          for (int c = l0; c < iq - 1; ++c) { act[c] = act[c + 1]; u[c] = u[c + 1]; asg[c] = asg[c + 1]; }
          u[iq - 1] = u[iq];
        }
      }
      PK_WSYNC();
      --iq;
    };
    // KKT-residual refinement in fp64 on the current active set (see pk_dualqp.cuh)
    auto polish = [&]() {
      PK_LANES(l) {
        #pragma unroll 1
        for (int k = l; k < iq; k += 32) ud[k] = (double)u[k];
      }
      PK_WSYNC();
      #pragma unroll 1
      for (int pass = 0; pass < PK_DUALQP_POLISH; ++pass) {
        // slacks of the active constraints -> wd[k] (right-hand side -s)
        #pragma unroll 1
        for (int k = 0; k < iq; ++k) {
          double sv = slack(act[k]);
          if (act[k] < meq) sv *= (double)asg[k];
          PK_LANES(l) { if (l == 0) wd[k] = -sv; }
        }
        PK_WSYNC();
        // Ra^T w = -s (forward substitution, column oriented)
        #pragma unroll 1
        for (int k = 0; k < iq; ++k) {
          const double wk = wd[k] / (double)RA[k * ld + k];
          PK_WSYNC();
          PK_LANES(l) {
            if (l == 0) wd[k] = wk;
            #pragma unroll 1
            for (int j = k + 1 + l; j < iq; j += 32) wd[j] -= (double)RA[k * ld + j] * wk;
          }
          PK_WSYNC();
        }
        PK_LANES(l) {
          #pragma unroll 1
          for (int i = l; i < n; i += 32) {
            double s = xd[i];
            #pragma unroll 1
            for (int k = 0; k < iq; ++k) s += (double)J[i * ld + k] * wd[k];
            xd[i] = s;
          }
        }
        PK_WSYNC();
        // stationarity residual grad f(x) - N_a u
        PK_LANES(l) {
          #pragma unroll 1
          for (int rr = l; rr < K; rr += 32) {
            double s = (double)bv[rr];
            #pragma unroll 1
            for (int j = 0; j < n; ++j) s += (double)A[rr * L.lda + j] * xd[j];
            rhod[rr] = s;
          }
        }
        PK_WSYNC();
        PK_LANES(l) {
          #pragma unroll 1
          for (int i = l; i < n; i += 32) {
            double g = (double)dg[i] * ((double)dg[i] * xd[i] + (double)beta[i]);
            #pragma unroll 1
            for (int rr = 0; rr < K; ++rr) g += (double)A[rr * L.lda + i] * rhod[rr];
            #pragma unroll 1
            for (int k = 0; k < iq; ++k) {
              const int id = act[k];
              if (id < meq) g -= ud[k] * (double)asg[k] * (double)E[id * L.lda + i];
              else if (id < pp) g += ud[k] * (double)G[(id - meq) * L.lda + i];
              else if (((id - pp) >> 1) == i) g -= ((id - pp) & 1) ? ud[k] : -ud[k];
            }
            gd[i] = g;
          }
        }
        PK_WSYNC();
        PK_LANES(l) {
          #pragma unroll 1
          for (int k = l; k < n; k += 32) {
            double s = 0.0;
            #pragma unroll 1
            for (int i = 0; i < n; ++i) s += (double)J[i * ld + k] * gd[i];
            wd[k] = s;
          }
        }
        PK_WSYNC();
        PK_LANES(l) {
          #pragma unroll 1
          for (int i = l; i < n; i += 32) {
            double s = xd[i];
            #pragma unroll 1
            for (int k = iq; k < n; ++k) s -= (double)J[i * ld + k] * wd[k];
            xd[i] = s;
          }
        }
        PK_WSYNC();
        // Ra du = w1 (back substitution), u += du
        #pragma unroll 1
        for (int k = iq - 1; k >= 0; --k) {
          const double wk = wd[k] / (double)RA[k * ld + k];
          PK_WSYNC();
          PK_LANES(l) {
            if (l == 0) { wd[k] = wk; ud[k] += wk; }
            #pragma unroll 1
            for (int i = l; i < k; i += 32) wd[i] -= (double)RA[i * ld + k] * wk;
          }
          PK_WSYNC();
        }
      }
      PK_LANES(l) {
        #pragma unroll 1
        for (int k = l; k < iq; k += 32) u[k] = (float)ud[k];
      }
      PK_WSYNC();
    };

    // Constraints that cannot enter (dependent on the active set, nothing to release) while
    // violated by less than the final check accepts (see pk_dualqp.cuh): up to four ids,
    // 8 bits each, 0xff = free.
    unsigned skip = 0xffffffffu;
    auto skipped = [&](int id) {
      const unsigned v = (unsigned)id;
      return ((skip & 255u) == v) | (((skip >> 8) & 255u) == v) | (((skip >> 16) & 255u) == v) | ((skip >> 24) == v);
    };
    // most violated constraint at xd (dense rows normalised by their norm); ip = 0x7fffffff:
    // none beyond the tolerance, ip = -1: an empty row that can never hold
    auto most_violated = [&](float tol_rel, int& ip, float& worst, bool all = false) {
        {
          LaneVar<float> bv_, dummy;
          LaneVar<int> bi;
          (void)dummy;
          PK_LANES(l) {
            float best = 0.f;
            int bid = 0x7fffffff;
            #pragma unroll 1
            for (int i = l; i < n; i += 32) {
              if ((all || !(((in_hi >> i) & 1ull) | skipped(pp + 2 * i))) && hi[i] < 3.0e38f) {
                const float s = (float)((double)hi[i] - xd[i]);
                if (s < -tol_rel * (fabsf(hi[i]) + 1e-3f) && s < best) { best = s; bid = pp + 2 * i; }
              }
              if ((all || !(((in_lo >> i) & 1ull) | skipped(pp + 2 * i + 1))) && lo[i] > -3.0e38f) {
                const float s = (float)(xd[i] - (double)lo[i]);
                if (s < -tol_rel * (fabsf(lo[i]) + 1e-3f) && s < best) { best = s; bid = pp + 2 * i + 1; }
              }
            }
            if (l < p && (all || !(((in_gen >> (meq + l)) & 1ull) | skipped(meq + l)))) {
              double sacc = (double)hg[l];
              #pragma unroll 1
              for (int k = 0; k < n; ++k) sacc -= (double)G[l * L.lda + k] * xd[k];
              const float s = (float)sacc * gn[l];
              if (gn[l] == 0.f) { if (hg[l] < 0.f) { best = -3.0e38f; bid = -1; } }
              else if (s < -tol_rel * (fabsf(hg[l]) * gn[l] + 1e-3f) && s < best) { best = s; bid = meq + l; }
            }
            if (l < meq && (all || !(((in_gen >> l) & 1ull) | skipped(l)))) {
              double sacc = -(double)fe[l];
              #pragma unroll 1
              for (int k = 0; k < n; ++k) sacc += (double)E[l * L.lda + k] * xd[k];
              const float s = -fabsf((float)sacc) * en[l];
              if (en[l] == 0.f) { if (fe[l] != 0.f) { best = -3.0e38f; bid = -1; } }
              else if (s < -tol_rel * (fabsf(fe[l]) * en[l] + 1e-3f) && s < best) { best = s; bid = l; }
            }
            bv_[l] = best;
            bi[l] = bid;
          }
          lane_argmin(bv_, bi, worst, ip);
        }
    };
    #pragma unroll 1
    for (int round = 0; round < 6 && !(status & (PK_STATUS_NO_SOLUTION | PK_STATUS_ITER_LIMIT)); ++round) {
      int changed = 0;
      #pragma unroll 1
      for (;; ++iter) {
        if (iter >= max_iter) { status |= PK_STATUS_ITER_LIMIT; break; }
        // step 1: most violated constraint (dense rows normalised by their norm)
        int ip;
        float worst;
        most_violated(vtol, ip, worst);
        if (ip == -1) { status |= PK_STATUS_NO_SOLUTION; break; }  // empty row with h < 0
        if (ip == 0x7fffffff) break;
        ++changed;
        // an equality enters with the sign that makes it a violated inequality
        float sgn = 1.f;
        if (ip < meq) sgn = (slack(ip) > 0.0) ? -1.f : 1.f;
        // normal of the entering constraint as a dense row (nullptr: box row) and its sign
        const float* nrow = (ip < meq) ? E + ip * L.lda : (ip < pp ? G + (ip - meq) * L.lda : nullptr);
        const float nsg = (ip < meq) ? sgn : -1.f;
        PK_LANES(l) { if (l == 0) u[iq] = 0.f; }
        bool added = false;
        #pragma unroll 1
        for (int inner = 0; inner <= n + pp + 2 && !added; ++inner) {
          // step 2a: d = J^T n+, z = J2 d2, r = Ra^-1 d1
          PK_LANES(l) {
            #pragma unroll 1
            for (int i = l; i < n; i += 32) {
              float s;
              if (ip >= pp) {
                const int c = (ip - pp) >> 1;
                s = ((ip - pp) & 1) ? J[c * ld + i] : -J[c * ld + i];
              } else {
                s = 0.f;
                #pragma unroll 1
                for (int k = 0; k < n; ++k) s = fmaf(J[k * ld + i], nrow[k], s);
                s *= nsg;
              }
              dv[i] = s;
            }
          }
          PK_WSYNC();
          float dd, d2;
          {
            LaneVar<float> pa, pb;
            PK_LANES(l) {
              float a = 0.f, b2 = 0.f;
              #pragma unroll 1
              for (int i = l; i < n; i += 32) {
                a = fmaf(dv[i], dv[i], a);
                if (i >= iq) b2 = fmaf(dv[i], dv[i], b2);
              }
              pa[l] = a;
              pb[l] = b2;
            }
            dd = lane_sum(pa);
            d2 = lane_sum(pb);
          }
          const bool dependent = !(d2 > 1e-10f * dd);
          PK_LANES(l) {
            #pragma unroll 1
            for (int i = l; i < n; i += 32) {
              float s = 0.f;
              #pragma unroll 1
              for (int k = iq; k < n; ++k) s = fmaf(J[i * ld + k], dv[k], s);
              z[i] = s;
            }
            #pragma unroll 1
            for (int i = l; i < iq; i += 32) r[i] = dv[i];
          }
          PK_WSYNC();
          #pragma unroll 1
          for (int k = iq - 1; k >= 0; --k) {
            const float rk = r[k] / RA[k * ld + k];
            PK_WSYNC();
            PK_LANES(l) {
              if (l == 0) r[k] = rk;
              #pragma unroll 1
              for (int i = l; i < k; i += 32) r[i] = fmaf(-RA[i * ld + k], rk, r[i]);
            }
            PK_WSYNC();
          }
          // step 2b: step lengths
          float t1;
          int l1;
          {
            LaneVar<float> tv;
            LaneVar<int> ti;
            PK_LANES(l) {
              float best = 3.0e38f;
              int bid = 0x7fffffff;
              #pragma unroll 1
              for (int k = l; k < iq; k += 32)
                if (act[k] >= meq && r[k] > 0.f) {
                  const float t = fmaxf(u[k], 0.f) / r[k];
                  if (t < best) { best = t; bid = k; }
                }
              tv[l] = best;
              ti[l] = bid;
            }
            lane_argmin(tv, ti, t1, l1);
          }
          float t2 = 3.0e38f;
          if (!dependent) {
            float zn;
            if (ip >= pp) {
              const int c = (ip - pp) >> 1;
              zn = ((ip - pp) & 1) ? z[c] : -z[c];
            } else {
              LaneVar<float> part;
              PK_LANES(l) {
                float s = 0.f;
                #pragma unroll 1
                for (int k = l; k < n; k += 32) s = fmaf(nrow[k], z[k], s);
                part[l] = s;
              }
              zn = nsg * lane_sum(part);
            }
            const float sp = (ip < meq) ? sgn * (float)slack(ip) : (float)slack(ip);
            if (zn > 0.f) t2 = fmaxf(-sp, 0.f) / zn;
          }
          const float t = fminf(t1, t2);
          if (!(t < 3.0e38f)) {
            float viol, base;
            if (ip >= pp) {
              const int c = (ip - pp) >> 1;
              base = fabsf(((ip - pp) & 1) ? lo[c] : hi[c]) + 1e-3f;
              viol = -(float)slack(ip);
            } else if (ip >= meq) {
              base = fabsf(hg[ip - meq]) * gn[ip - meq] + 1e-3f;
              viol = -(float)slack(ip) * gn[ip - meq];
            } else {
              base = fabsf(fe[ip]) * en[ip] + 1e-3f;
              viol = fabsf((float)slack(ip)) * en[ip];
            }
            if (viol <= 1e-4f * base && (skip >> 24) == 255u) {
              skip = (skip << 8) | (unsigned)ip;  // tolerated: state untouched, look elsewhere
              added = true;
              break;
            }
            status |= PK_STATUS_NO_SOLUTION;
            break;
          }
          PK_WSYNC();
          PK_LANES(l) {
            if (t2 < 3.0e38f) {
              #pragma unroll 1
              for (int i = l; i < n; i += 32) xd[i] += (double)t * (double)z[i];
            }
            #pragma unroll 1
            for (int k = l; k < iq; k += 32) u[k] = fmaf(-t, r[k], u[k]);
            if (l == 0) u[iq] += t;
          }
          PK_WSYNC();
          if (t2 <= t1) {
            // full step: the constraint enters; one Householder reflection of J[:, iq..]
            // maps d2 onto alpha e_1
            const float nrm = sqrtf(d2);
            const float alpha = (dv[iq] > 0.f) ? -nrm : nrm;
            const float v0 = dv[iq] - alpha;
            const float vnorm2 = fmaf(v0, v0, d2 - dv[iq] * dv[iq]);
            const float tau = (vnorm2 > 0.f) ? 2.f / vnorm2 : 0.f;
            PK_LANES(l) {
              #pragma unroll 1
              for (int i = l; i < n; i += 32) {
                float w = J[i * ld + iq] * v0;
                #pragma unroll 1
                for (int k = iq + 1; k < n; ++k) w = fmaf(J[i * ld + k], dv[k], w);
                w *= tau;
                J[i * ld + iq] = fmaf(-w, v0, J[i * ld + iq]);
                #pragma unroll 1
                for (int k = iq + 1; k < n; ++k) J[i * ld + k] = fmaf(-w, dv[k], J[i * ld + k]);
              }
              #pragma unroll 1
              for (int k = l; k < iq; k += 32) RA[k * ld + iq] = dv[k];
              if (l == 0) {
                RA[iq * ld + iq] = alpha;
                act[iq] = ip;
                asg[iq] = sgn;
              }
            }
            PK_WSYNC();
            set_member(ip, true);
            ++iq;
            added = true;
            break;
          }
          drop(l1);  // partial step: constraint l1 leaves the active set
        }
        if (status & PK_STATUS_NO_SOLUTION) break;
        if (!added) { status |= PK_STATUS_ITER_LIMIT; break; }
      }
      if (status & (PK_STATUS_NO_SOLUTION | PK_STATUS_ITER_LIMIT)) break;
      if (round > 0 && !changed) break;
      #pragma unroll 1
      for (int rel = 0; rel <= n; ++rel) {
        polish();
        float umin, umax;
        int lmin;
        {
          LaneVar<float> mn, mx;
          LaneVar<int> mi;
          PK_LANES(l) {
            float a = 0.f, b2 = 0.f;
            int bi = 0x7fffffff;
            #pragma unroll 1
            for (int k = l; k < iq; k += 32) {
              b2 = fmaxf(b2, fabsf(u[k]));
              if (act[k] >= meq && u[k] < a) { a = u[k]; bi = k; }
            }
            mn[l] = a;
            mi[l] = bi;
            mx[l] = -b2;
          }
          int dummy;
          lane_argmin(mn, mi, umin, lmin);
          lane_argmin(mx, mi, umax, dummy);
          umax = -umax;
        }
        if (lmin == 0x7fffffff || umin >= -1e-6f * umax) break;
        PK_LANES(l) { if (l == 0) u[iq] = 0.f; }
        PK_WSYNC();
        drop(lmin);
        ++changed;
      }
      vtol = 1e-9f;
    }
    if (!(status & (PK_STATUS_NO_SOLUTION | PK_STATUS_ITER_LIMIT))) {
      // the rounds may run out on inconsistent rows that fp32 keeps "almost consistent":
      // a point that still violates a constraint is not a solution
      int ipf;
      float wf;
      most_violated(1e-4f, ipf, wf, true);
      if (ipf != 0x7fffffff) status |= PK_STATUS_NO_SOLUTION;
    }
    if (status & PK_STATUS_NO_SOLUTION) {
      PK_LANES(l) {
        #pragma unroll 1
        for (int i = l; i < n; i += 32) x[i] = 0.f;
      }
      PK_WSYNC();
      return status;
    }
    // result: coordinates on a bound sit exactly on it; absurd magnitudes = rows that meet
    // at infinity (see pk_dualqp.cuh)
    {
      LaneVar<int> bad;
      PK_LANES(l) {
        int f = 0;
        #pragma unroll 1
        for (int i = l; i < n; i += 32) {
          float xi = (float)xd[i];
          if ((in_hi >> i) & 1ull) xi = hi[i];
          else if ((in_lo >> i) & 1ull) xi = lo[i];
          if (!(fabsf(xi) < 1e3f)) f = 1;
          x[i] = xi;
        }
        bad[l] = f;
      }
      PK_WSYNC();
      if (lane_or(bad)) {
        status |= PK_STATUS_NO_SOLUTION;
        PK_LANES(l) {
          #pragma unroll 1
          for (int i = l; i < n; i += 32) x[i] = 0.f;
        }
        PK_WSYNC();
      }
    }
    return status;
  }
};

PK_HD int tree_dual_solve(float* W, const TreePlan& L) { return TreeDual::solve(W, L); }

}  // namespace pk
