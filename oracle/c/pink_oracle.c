/*
 * C port of the oracle (fp64) for fixed-base joint trees with FrameTask +
 * PostureTask and the two default limits: the path of BASELINE configs 1, 2, 5
 * (examples/arm_ur5.py).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py): used by the
 * tests as a second, independent fp64 checker that is fast enough for full-size
 * batches, and by bench.py as the CPU baseline ("kind": "port") - Pink,
 * Pinocchio and quadprog themselves cannot be installed offline.
 *
 * It follows the reference step by step and, unlike the CUDA path, keeps the
 * reference's own formulation: dense H = sum (W J)^T (W J) + mu I
 * (pink/tasks/task.py:145-166), dense G = [P; -P; P; -P], h stacked in the
 * order of pink/solve_ik.py:100-113, and a Goldfarb-Idnani dual active-set QP
 * (what solver="quadprog" runs at pink/solve_ik.py:270) with Cholesky factor,
 * J = L^-T Q and R kept up to date by Givens rotations.
 *
 * Build: make -C oracle/c   ->  oracle/c/libpink_oracle.so
 */
#include <math.h>
#include <pthread.h>
#include <limits.h>
#include <linux/futex.h>
#include <sched.h>
#include <stdio.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXJ 16
#define MAXM (4 * MAXJ)
#define MAXT 4

typedef struct {
  int njoints;
  const int *parent, *jtype;
  const double *joint_placement; /* [nj][12] */
  const double *axis;            /* [nj][3]  */
  const double *q_min, *q_max, *v_max;
} OcModel;

typedef struct {
  int n_frame_tasks;
  int frame_body[MAXT];        /* joint the frame is fixed to (-1 world) */
  double frame_placement[MAXT][12];
  double frame_cost[MAXT][6];
  double frame_gain[MAXT], frame_lm[MAXT];
  int has_posture;
  double posture_cost, posture_gain, posture_lm;
  double posture_target[MAXJ];
  double dt, damping, cfg_gain;
  int use_cfg_limit, use_vel_limit, safety_break;
} OcProblem;

/* ---- small dense helpers ------------------------------------------------------ */
typedef struct { double R[9], p[3]; } SE3;

static void se3_from12(const double *t, SE3 *T) {
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T->R[3 * i + j] = t[4 * i + j]; T->p[i] = t[4 * i + 3]; }
}
static void mat3_mul(const double *A, const double *B, double *C) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3_vec(const double *A, const double *v, double *o) {
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static void mat3T_vec(const double *A, const double *v, double *o) {
  for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
static void se3_mul(const SE3 *a, const SE3 *b, SE3 *o) {
  SE3 r; mat3_mul(a->R, b->R, r.R); mat3_vec(a->R, b->p, r.p);
  for (int i = 0; i < 3; ++i) r.p[i] += a->p[i];
  *o = r;
}
static void se3_act_inv(const SE3 *a, const SE3 *b, SE3 *o) { /* a^-1 b */
  SE3 r; double d[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    r.R[3 * i + j] = a->R[i] * b->R[j] + a->R[3 + i] * b->R[3 + j] + a->R[6 + i] * b->R[6 + j];
  for (int i = 0; i < 3; ++i) d[i] = b->p[i] - a->p[i];
  mat3T_vec(a->R, d, r.p);
  *o = r;
}
static void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* log3 with Pinocchio's branches: w, theta */
static void log3(const double *R, double *w, double *theta_out) {
  double tr = R[0] + R[4] + R[8];
  double ct = 0.5 * (tr - 1.0); if (ct > 1.0) ct = 1.0; if (ct < -1.0) ct = -1.0;
  double theta = acos(ct);
  double vee[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  if (theta >= M_PI - 1e-2) {
    double omc = 1.0 - ct;
    double d[3] = {R[0], R[4], R[8]};
    for (int k = 0; k < 3; ++k) {
      double t = (d[k] - ct) / omc; if (t < 0.0) t = 0.0;
      w[k] = (vee[k] >= 0.0 ? 1.0 : -1.0) * theta * sqrt(t);
    }
  } else {
    double fac = theta < 1e-4 ? 0.5 * (1.0 + theta * theta / 6.0 + 7.0 * pow(theta, 4) / 360.0) : 0.5 * theta / sin(theta);
    for (int k = 0; k < 3; ++k) w[k] = fac * vee[k];
  }
  *theta_out = theta;
}
static void alpha_beta(double theta, double *alpha, double *beta) {
  double t2 = theta * theta;
  if (theta < 1e-4) { *alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0; *beta = 1.0 / 12.0 + t2 / 720.0; }
  else {
    double st = sin(theta), ct = cos(theta);
    *alpha = theta * st / (2.0 * (1.0 - ct));
    *beta = 1.0 / t2 - st / (2.0 * theta * (1.0 - ct));
  }
}
static void log6(const SE3 *T, double *e) {
  double w[3], theta, alpha, beta, wxp[3];
  log3(T->R, w, &theta); alpha_beta(theta, &alpha, &beta);
  double wp = w[0] * T->p[0] + w[1] * T->p[1] + w[2] * T->p[2];
  cross3(w, T->p, wxp);
  for (int k = 0; k < 3; ++k) { e[k] = alpha * T->p[k] - 0.5 * wxp[k] + beta * wp * w[k]; e[3 + k] = w[k]; }
}
static void hat(const double *v, double *M) {
  M[0] = 0; M[1] = -v[2]; M[2] = v[1]; M[3] = v[2]; M[4] = 0; M[5] = -v[0]; M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}
/* Jlog6 = [[A, B], [0, A]] */
static void jlog6(const SE3 *T, double *J /*6x6*/) {
  double w[3], theta, alpha, a, A[9], C[9], B[9], W[9], Pm[9];
  log3(T->R, w, &theta); alpha_beta(theta, &alpha, &a);
  double t2 = theta * theta, d = 1.0 - t2 * a;
  hat(w, W); hat(T->p, Pm);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    A[3 * i + j] = a * w[i] * w[j] + (i == j ? d : 0.0) + 0.5 * W[3 * i + j];
  double bd;
  if (theta < 1e-4) bd = 1.0 / 360.0 + t2 / 7560.0;
  else { double st = sin(theta), ct = cos(theta); bd = -2.0 / (t2 * t2) + (1.0 + st / theta) / (t2 * 2.0 * (1.0 - ct)); }
  const double *p = T->p;
  double wp = w[0] * p[0] + w[1] * p[1] + w[2] * p[2];
  double v3[3];
  for (int k = 0; k < 3; ++k) v3[k] = bd * wp * w[k] - (t2 * bd + 2.0 * a) * p[k];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    C[3 * i + j] = v3[i] * w[j] + a * w[i] * p[j] + (i == j ? a * wp : 0.0) + 0.5 * Pm[3 * i + j];
  mat3_mul(C, A, B);
  memset(J, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    J[6 * i + j] = A[3 * i + j]; J[6 * i + 3 + j] = B[3 * i + j]; J[6 * (3 + i) + 3 + j] = A[3 * i + j];
  }
}

/* ---- Goldfarb-Idnani dual active set (quadprog semantics) ------------------------- */
/* minimise 1/2 x^T G x + g0^T x  s.t.  CI^T x + ci0 >= 0   (CI is n x m, column i = normal i) */
static int gi_solve(int n, int m, const double *G, const double *g0, const double *CI, const double *ci0, double *x) {
  double L[MAXJ * MAXJ], Jm[MAXJ * MAXJ], R[MAXJ * MAXJ], d[MAXJ], z[MAXJ], r[MAXJ], u[MAXM], np_[MAXJ], nrm[MAXM];
  int A[MAXM], iai[MAXM], iq = 0;
  /* Cholesky G = L L^T */
  for (int j = 0; j < n; ++j) {
    double s = G[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    if (s <= 0.0) return 1;
    L[j * n + j] = sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = G[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
    for (int i = 0; i < j; ++i) L[i * n + j] = 0.0;
  }
  /* J = L^-T : solve L^T J = I column by column */
  for (int c = 0; c < n; ++c) {
    for (int i = n - 1; i >= 0; --i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * Jm[k * n + c];
      Jm[i * n + c] = s / L[i * n + i];
    }
  }
  /* x = -G^-1 g0 = -J J^T g0 */
  for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += Jm[k * n + i] * g0[k]; d[i] = s; }
  for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += Jm[i * n + k] * d[k]; x[i] = -s; }
  for (int i = 0; i < m; ++i) {
    double s = 0; for (int k = 0; k < n; ++k) s += CI[k * m + i] * CI[k * m + i];
    nrm[i] = sqrt(s) > 0 ? sqrt(s) : 1.0; iai[i] = i;
  }
  memset(R, 0, sizeof(R));
  for (int iter = 0; iter < 40 * (n + m) + 100; ++iter) {
    /* step 1: most violated constraint (normalised, as quadprog) */
    int ip = -1; double worst = -1e-10;
    for (int i = 0; i < m; ++i) {
      if (iai[i] < 0) continue;
      double s = ci0[i]; for (int k = 0; k < n; ++k) s += CI[k * m + i] * x[k];
      s /= nrm[i];
      if (s < worst) { worst = s; ip = i; }
    }
    if (ip < 0) return 0;
    for (int k = 0; k < n; ++k) np_[k] = CI[k * m + ip];
    u[iq] = 0.0;
    for (;;) {
      /* step 2a: d = J^T n+, z = J2 d2, r = R^-1 d1 */
      for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += Jm[k * n + i] * np_[k]; d[i] = s; }
      for (int i = 0; i < n; ++i) { double s = 0; for (int k = iq; k < n; ++k) s += Jm[i * n + k] * d[k]; z[i] = s; }
      for (int i = iq - 1; i >= 0; --i) {
        double s = d[i]; for (int k = i + 1; k < iq; ++k) s -= R[i * n + k] * r[k];
        r[i] = s / R[i * n + i];
      }
      /* step 2b: step lengths */
      double t1 = INFINITY; int l = -1;
      for (int k = 0; k < iq; ++k) if (r[k] > 0.0 && u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = k; }
      double zz = 0; for (int k = 0; k < n; ++k) zz += z[k] * z[k];
      double t2 = INFINITY;
      double sp = ci0[ip]; for (int k = 0; k < n; ++k) sp += np_[k] * x[k];
      if (zz > 1e-28) { double zn = 0; for (int k = 0; k < n; ++k) zn += z[k] * np_[k]; t2 = -sp / zn; }
      double t = t1 < t2 ? t1 : t2;
      if (!isfinite(t)) return 2; /* infeasible */
      if (!isfinite(t2)) {
        for (int k = 0; k < iq; ++k) u[k] -= t * r[k];
        u[iq] += t;
      } else {
        for (int k = 0; k < n; ++k) x[k] += t * z[k];
        for (int k = 0; k < iq; ++k) u[k] -= t * r[k];
        u[iq] += t;
        if (t2 <= t1) {
          /* full step: add constraint ip. Givens rotations zero d[iq+1 .. n-1] */
          for (int j = n - 1; j > iq; --j) {
            double a = d[j - 1], b = d[j];
            if (b == 0.0) continue;
            double h = hypot(a, b), c = a / h, s = b / h;
            d[j - 1] = h; d[j] = 0.0;
            for (int k = 0; k < n; ++k) {
              double ja = Jm[k * n + j - 1], jb = Jm[k * n + j];
              Jm[k * n + j - 1] = c * ja + s * jb;
              Jm[k * n + j] = -s * ja + c * jb;
            }
          }
          for (int k = 0; k <= iq; ++k) R[k * n + iq] = d[k];
          A[iq] = ip; iai[ip] = -1; ++iq;
          break;
        }
      }
      /* drop constraint l: remove column l of R, restore triangularity */
      iai[A[l]] = A[l];
      for (int c = l; c < iq - 1; ++c) {
        A[c] = A[c + 1]; u[c] = u[c + 1];
        for (int k = 0; k < n; ++k) R[k * n + c] = R[k * n + c + 1];
      }
      u[iq - 1] = u[iq];
      for (int k = 0; k < n; ++k) R[k * n + iq - 1] = 0.0;
      --iq;
      for (int j = l; j < iq; ++j) {
        double a = R[j * n + j], b = R[(j + 1) * n + j];
        if (b == 0.0) continue;
        double h = hypot(a, b), c = a / h, s = b / h;
        for (int k = j; k < iq; ++k) {
          double ra = R[j * n + k], rb = R[(j + 1) * n + k];
          R[j * n + k] = c * ra + s * rb;
          R[(j + 1) * n + k] = -s * ra + c * rb;
        }
        for (int k = 0; k < n; ++k) {
          double ja = Jm[k * n + j], jb = Jm[k * n + j + 1];
          Jm[k * n + j] = c * ja + s * jb;
          Jm[k * n + j + 1] = -s * ja + c * jb;
        }
      }
    }
  }
  return 3;
}

/* ---- one IK step --------------------------------------------------------------- */
static int solve_one(const OcModel *M, const OcProblem *P, const double *q, const double *targets /*[nft][12]*/,
                     double *v) {
  const int n = M->njoints;
  for (int j = 0; j < n; ++j) v[j] = 0.0;
  /* Configuration.check_limits (pink/configuration.py:181-201) */
  int outside = 0;
  for (int j = 0; j < n; ++j) {
    if (M->q_max[j] <= M->q_min[j] + 1e-6) continue;
    if (q[j] < M->q_min[j] - 1e-6 || q[j] > M->q_max[j] + 1e-6) outside = 1;
  }
  if (outside && P->safety_break) return 2;
  /* forward kinematics */
  SE3 Tw[MAXJ];
  for (int j = 0; j < n; ++j) {
    SE3 X, Mo, Tl; se3_from12(M->joint_placement + 12 * j, &X);
    const double *a = M->axis + 3 * j;
    memset(&Mo, 0, sizeof(Mo));
    if (M->jtype[j] == 0) {
      double s = sin(q[j]), c = cos(q[j]), t = 1.0 - c;
      double Rr[9] = {t * a[0] * a[0] + c, t * a[0] * a[1] - s * a[2], t * a[0] * a[2] + s * a[1],
                      t * a[0] * a[1] + s * a[2], t * a[1] * a[1] + c, t * a[1] * a[2] - s * a[0],
                      t * a[0] * a[2] - s * a[1], t * a[1] * a[2] + s * a[0], t * a[2] * a[2] + c};
      memcpy(Mo.R, Rr, sizeof(Rr));
    } else {
      Mo.R[0] = Mo.R[4] = Mo.R[8] = 1.0;
      for (int k = 0; k < 3; ++k) Mo.p[k] = q[j] * a[k];
    }
    se3_mul(&X, &Mo, &Tl);
    if (M->parent[j] < 0) Tw[j] = Tl; else se3_mul(&Tw[M->parent[j]], &Tl, &Tw[j]);
  }
  /* objective (pink/solve_ik.py:55-60, pink/tasks/task.py:145-166) */
  double H[MAXJ * MAXJ], c[MAXJ];
  for (int i = 0; i < n * n; ++i) H[i] = 0.0;
  for (int i = 0; i < n; ++i) { H[i * n + i] = P->damping; c[i] = 0.0; }
  for (int t = 0; t < P->n_frame_tasks; ++t) {
    SE3 Xf, Tf, Tt, Tbt, Ttb; se3_from12(P->frame_placement[t], &Xf); se3_from12(targets + 12 * t, &Tt);
    if (P->frame_body[t] < 0) Tf = Xf; else se3_mul(&Tw[P->frame_body[t]], &Xf, &Tf);
    se3_act_inv(&Tf, &Tt, &Tbt); se3_act_inv(&Tt, &Tf, &Ttb);
    double e[6], JL[36], Jf[6 * MAXJ], Jt[6 * MAXJ];
    log6(&Tbt, e); jlog6(&Ttb, JL);
    for (int j = 0; j < n; ++j) {
      double lin[3] = {0, 0, 0}, ang[3] = {0, 0, 0}, aw[3], dp[3], l2[3], a2[3];
      int sup = 0; for (int b = P->frame_body[t]; b >= 0; b = M->parent[b]) if (b == j) sup = 1;
      if (sup) {
        mat3_vec(Tw[j].R, M->axis + 3 * j, aw);
        if (M->jtype[j] == 0) { for (int k = 0; k < 3; ++k) dp[k] = Tf.p[k] - Tw[j].p[k]; cross3(aw, dp, lin); memcpy(ang, aw, sizeof(ang)); }
        else memcpy(lin, aw, sizeof(lin));
      }
      mat3T_vec(Tf.R, lin, l2); mat3T_vec(Tf.R, ang, a2);
      for (int k = 0; k < 3; ++k) { Jf[k * n + j] = l2[k]; Jf[(3 + k) * n + j] = a2[k]; }
    }
    for (int r = 0; r < 6; ++r) for (int j = 0; j < n; ++j) {
      double s = 0; for (int k = 0; k < 6; ++k) s += JL[6 * r + k] * Jf[k * n + j];
      Jt[r * n + j] = -s;
    }
    double mu = 0, we[6];
    for (int r = 0; r < 6; ++r) { we[r] = P->frame_cost[t][r] * (-P->frame_gain[t] * e[r]); mu += we[r] * we[r]; }
    mu *= P->frame_lm[t];
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        double s = 0; for (int r = 0; r < 6; ++r) s += P->frame_cost[t][r] * Jt[r * n + i] * P->frame_cost[t][r] * Jt[r * n + j];
        H[i * n + j] += s;
      }
      H[i * n + i] += mu;
      double s = 0; for (int r = 0; r < 6; ++r) s += we[r] * P->frame_cost[t][r] * Jt[r * n + i];
      c[i] -= s;
    }
  }
  if (P->has_posture) {
    double w = P->posture_cost, mu = 0;
    for (int i = 0; i < n; ++i) { double we = w * (-P->posture_gain * (q[i] - P->posture_target[i])); mu += we * we; }
    mu *= P->posture_lm;
    for (int i = 0; i < n; ++i) {
      H[i * n + i] += w * w + mu;
      c[i] -= w * (-P->posture_gain * (q[i] - P->posture_target[i])) * w;
    }
  }
  /* inequalities G x <= h (pink/solve_ik.py:94-122) in quadprog form C^T x >= b: C = -G^T, b = -h */
  double CI[MAXJ * MAXM], ci0[MAXM]; int m = 0;
  int idx_c[MAXJ], nc = 0, idx_v[MAXJ], nvl = 0;
  for (int j = 0; j < n; ++j) {
    if (P->use_cfg_limit && M->q_max[j] < 1e20 && M->q_max[j] > M->q_min[j] + 1e-10) idx_c[nc++] = j;
    if (P->use_vel_limit && M->v_max[j] < 1e20 && M->v_max[j] > 1e-10) idx_v[nvl++] = j;
  }
  m = 2 * nc + 2 * nvl;
  for (int i = 0; i < n * m; ++i) CI[i] = 0.0;
  int row = 0;
  for (int k = 0; k < nc; ++k, ++row) { CI[idx_c[k] * m + row] = -1.0; ci0[row] = P->cfg_gain * (M->q_max[idx_c[k]] - q[idx_c[k]]); }
  for (int k = 0; k < nc; ++k, ++row) { CI[idx_c[k] * m + row] = 1.0; ci0[row] = -P->cfg_gain * (M->q_min[idx_c[k]] - q[idx_c[k]]); }
  for (int k = 0; k < nvl; ++k, ++row) { CI[idx_v[k] * m + row] = -1.0; ci0[row] = P->dt * M->v_max[idx_v[k]]; }
  for (int k = 0; k < nvl; ++k, ++row) { CI[idx_v[k] * m + row] = 1.0; ci0[row] = P->dt * M->v_max[idx_v[k]]; }
  double x[MAXJ];
  int rc = gi_solve(n, m, H, c, CI, ci0, x);
  if (rc) return 1;
  for (int j = 0; j < n; ++j) v[j] = x[j] / P->dt; /* pink/solve_ik.py:274 */
  return 0;
}

/* ---- batch entry points ----------------------------------------------------------- */
typedef struct {
  const OcModel *M; const OcProblem *P; const double *q, *targets; double *v; int32_t *status; int64_t B; int tstride;
  int threads;
} Job;

/* Persistent worker pool.
 *  - Workers are created once (the pool grows when a call asks for more), each pinned to one
 *    CPU of the process's affinity set (distinct physical cores first, then their
 *    hyper-thread siblings).
 *  - A job is published by bumping a futex word; workers claim chunks of instances from one
 *    64-bit counter that carries the job's epoch in its high bits (guided self-scheduling:
 *    chunk = remaining / (4 threads), at least 8), so a worker that wakes up late - or still
 *    holds an older job - can never touch the current job's index space.
 *  - The caller works too and the call returns when all INSTANCES are done, not when all
 *    workers have checked in: a worker whose CPU is busy with somebody else's process (the
 *    GPU boxes are shared) costs at most its one chunk in flight.
 * The first version of this pool woke the workers through a condition variable and waited
 * for every one of them to report back: on a 128-thread box the mutex convoy and the
 * stragglers made a 2 ms job take 15-20 ms (parallel efficiency 0.11).                   */
#define POOL_MAX 512
#define EPOCH_SHIFT 40
#define INDEX_MASK ((1ull << EPOCH_SHIFT) - 1ull)
static pthread_t g_th[POOL_MAX];
static int g_tid[POOL_MAX];
static int g_nworkers = 0;
static uint32_t g_epoch = 0;   /* futex word: current job number */
static uint64_t g_ctr = 0;     /* (epoch << 40) | next unclaimed instance */
static int64_t g_done = 0;     /* instances finished in the current job */
static Job g_job;
static pthread_mutex_t g_call_mu = PTHREAD_MUTEX_INITIALIZER; /* one batch call at a time */

static void futex_wait(uint32_t *addr, uint32_t val) { syscall(SYS_futex, addr, FUTEX_WAIT_PRIVATE, val, 0, 0, 0); }
static void futex_wake_all(uint32_t *addr) { syscall(SYS_futex, addr, FUTEX_WAKE_PRIVATE, INT_MAX, 0, 0, 0); }

static void run_chunks(const Job *j, uint64_t epoch) {
  const int n = j->M->njoints;
  for (;;) {
    uint64_t c = __atomic_load_n(&g_ctr, __ATOMIC_ACQUIRE);
    if ((c >> EPOCH_SHIFT) != epoch) return;          /* a newer job owns the counter */
    const int64_t lo = (int64_t)(c & INDEX_MASK);
    if (lo >= j->B) return;
    int64_t chunk = (j->B - lo) / (4 * (int64_t)j->threads);
    if (chunk < 8) chunk = 8;
    if (chunk > j->B - lo) chunk = j->B - lo;
    if (!__atomic_compare_exchange_n(&g_ctr, &c, c + (uint64_t)chunk, 1, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) continue;
    for (int64_t i = lo; i < lo + chunk; ++i)
      j->status[i] = solve_one(j->M, j->P, j->q + i * n, j->targets + i * j->tstride, j->v + i * n);
    __atomic_add_fetch(&g_done, chunk, __ATOMIC_ACQ_REL);
  }
}

static void *pool_worker(void *arg) {
  const int id = *(int *)arg;
  uint32_t seen = 0;
  for (;;) {
    uint32_t e;
    int spin = 0;
    while ((e = __atomic_load_n(&g_epoch, __ATOMIC_ACQUIRE)) == seen) {
      if (++spin < 200) __builtin_ia32_pause();      /* consecutive bench steps arrive back to back */
      else futex_wait(&g_epoch, seen);
    }
    seen = e;
    const Job job = g_job;                             /* complete: written before the epoch was published */
    if (__atomic_load_n(&g_epoch, __ATOMIC_ACQUIRE) != e) continue; /* already superseded */
    if (id < job.threads) run_chunks(&job, (uint64_t)e);
  }
  return 0;
}

/* CPUs of the affinity set, one hyper-thread per physical core first. */
static int ordered_cpus(int *out) {
  cpu_set_t allowed;
  int n = 0, cpus[CPU_SETSIZE], ncpu = 0;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return 0;
  for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  char taken[CPU_SETSIZE]; memset(taken, 0, sizeof(taken));
  for (int pass = 0; pass < 2; ++pass)
    for (int k = 0; k < ncpu; ++k) {
      const int c = cpus[k];
      if (taken[c]) continue;
      int first = c;  /* lowest sibling id */
      char path[128]; snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list", c);
      FILE *f = fopen(path, "r");
      if (f) { if (fscanf(f, "%d", &first) != 1) first = c; fclose(f); }
      if (pass == 0 && first != c) continue;   /* siblings go to the second pass */
      taken[c] = 1; out[n++] = c;
    }
  return n;
}

static void pool_grow(int threads) {
  int cpus[CPU_SETSIZE];
  const int ncpu = ordered_cpus(cpus);
  const char *pin = getenv("OC_POOL_PIN");
  const int do_pin = !(pin && pin[0] == '0');
  while (g_nworkers < threads && g_nworkers < POOL_MAX) {
    const int t = g_nworkers;
    g_tid[t] = t;
    if (pthread_create(&g_th[t], 0, pool_worker, &g_tid[t])) break;
    if (do_pin && ncpu > 0) {
      /* worker t shares the machine with the calling thread, which works too: shift by one */
      cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[(t + 1) % ncpu], &one);
      pthread_setaffinity_np(g_th[t], sizeof(one), &one);
    }
    g_nworkers++;
  }
}

/* Number of pool workers alive (diagnostics for the bench line). */
int oc_pool_size(void) { return g_nworkers; }

/* Solve B instances with `threads` threads (the caller + threads - 1 workers).
 * targets: [B][12 * n_frame_tasks]. Returns 0. */
int oc_solve_ik_batch(const OcModel *M, const OcProblem *P, const double *q, const double *targets, double *v,
                      int32_t *status, int64_t B, int threads) {
  if (M->njoints > MAXJ || P->n_frame_tasks > MAXT) return 1;
  if (B >= (int64_t)INDEX_MASK) return 1;
  if (threads < 1) threads = 1;
  if (threads > POOL_MAX) threads = POOL_MAX;
  pthread_mutex_lock(&g_call_mu);
  if (threads > 1) pool_grow(threads - 1);
  Job job = {M, P, q, targets, v, status, B, 12 * P->n_frame_tasks, threads - 1};
  if (job.threads > g_nworkers) job.threads = g_nworkers;
  const uint32_t e = g_epoch + 1;
  g_job = job;                                     /* workers with id < job.threads take part */
  Job mine = job; mine.threads = job.threads + 1;  /* the caller's chunk sizing counts itself */
  __atomic_store_n(&g_done, 0, __ATOMIC_RELAXED);
  __atomic_store_n(&g_ctr, (uint64_t)e << EPOCH_SHIFT, __ATOMIC_RELEASE);
  __atomic_store_n(&g_epoch, e, __ATOMIC_RELEASE);
  if (job.threads > 0) futex_wake_all(&g_epoch);
  run_chunks(&mine, (uint64_t)e);
  while (__atomic_load_n(&g_done, __ATOMIC_ACQUIRE) < B) __builtin_ia32_pause();
  pthread_mutex_unlock(&g_call_mu);
  return 0;
}
