// CPU build of the kernel bodies (pk_chain.cuh / pk_generic.cuh) for the test
// suite ONLY: lets `pytest -m "not gpu"` exercise the exact fp32 arithmetic and
// control flow of the CUDA kernels against the fp64 oracle on machines without
// a GPU.  It is never loaded by the product package (pink_b200/_cabi.py loads
// libpink_b200.so and nothing else) and is not a fallback.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pink_b200.h"
#include "../../pink_b200/csrc/pk_marshal.hpp"
#include "../../pink_b200/csrc/pk_dualqp.cuh"

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

template <int NJ>
void run_chain(const pk::HostModel& hm, const pk::DevProblem& P, const float* q, const float* targets, float* v,
               int32_t* status, int64_t B) {
  pk::ChainParams<NJ> C;
  pk::make_chain_params<NJ>(hm, P, &C, P.ext);
  for (int64_t i = 0; i < B; ++i) {
    float qi[NJ], vi[NJ];
    for (int k = 0; k < NJ; ++k) qi[k] = q[i * NJ + k];
    int st = 0;
    const float* trow = targets + i * (int64_t)P.target_stride;
    switch (C.n_frame_tasks) {
      case 0: pk::ik_step_chain<NJ, 0>(C, qi, trow, vi, st); break;
      case 1: pk::ik_step_chain<NJ, 1>(C, qi, trow, vi, st); break;
      default: pk::ik_step_chain<NJ, 2>(C, qi, trow, vi, st); break;
    }
    for (int k = 0; k < NJ; ++k) v[i * NJ + k] = vi[k];
    if (status) status[i] = st;
  }
}

// Sub-warp chain kernel body (pk_coop.cuh) with L lanes per instance, lanes emulated by loops.
template <int NJ, int NFT, int L>
void run_coop_nft(const pk::HostModel& hm, const pk::DevProblem& P, const float* q, const float* targets, float* v,
                  int32_t* status, int64_t B) {
  using Step = pk::CoopStep<NJ, NFT, L>;
  typename Step::Params C;
  pk::make_coop_params<NJ, L>(hm, P, &C, P.ext);
  const pk::Group<L> G{};
  auto jc = [&](int j) -> const pk::CoopJoint& { return C.joint[j]; };
  for (int64_t i = 0; i < B; ++i) {
    pk::GVar<typename Step::Lane, L> S;
    for (int h = 0; h < L; ++h)
      for (int k = 0; k < Step::NC; ++k) {
        const int j = h * Step::NC + k;
        S[h].q[k] = j < NJ ? q[i * NJ + j] : 0.f;
      }
    int st = 0;
    const float* trow = targets + i * (int64_t)P.target_stride;
    pk::ik_step_coop<NJ, NFT, L>(G, C, jc, trow, S, st);
    for (int h = 0; h < L; ++h)
      for (int k = 0; k < Step::NC; ++k) {
        const int j = h * Step::NC + k;
        if (j < NJ) v[i * NJ + j] = S[h].x[k];
      }
    if (status) status[i] = st;
  }
}

template <int NJ, int L>
void run_coop(const pk::HostModel& hm, const pk::DevProblem& P, const float* q, const float* targets, float* v,
              int32_t* status, int64_t B) {
  int nft = 0;
  for (int t = 0; t < P.ntasks; ++t) nft += P.tasks[t].type == PK_TASK_FRAME;
  switch (nft) {
    case 0: run_coop_nft<NJ, 0, L>(hm, P, q, targets, v, status, B); break;
    case 1: run_coop_nft<NJ, 1, L>(hm, P, q, targets, v, status, B); break;
    default: run_coop_nft<NJ, 2, L>(hm, P, q, targets, v, status, B); break;
  }
}

template <int L>
bool run_coop_nj(const pk::HostModel& hm, const pk::DevProblem& P, const float* q, const float* targets, float* v,
                 int32_t* status, int64_t B) {
  switch (hm.njoints) {
    case 2: run_coop<2, L>(hm, P, q, targets, v, status, B); return true;
    case 3: run_coop<3, L>(hm, P, q, targets, v, status, B); return true;
    case 4: run_coop<4, L>(hm, P, q, targets, v, status, B); return true;
    case 5: run_coop<5, L>(hm, P, q, targets, v, status, B); return true;
    case 6: run_coop<6, L>(hm, P, q, targets, v, status, B); return true;
    case 7: run_coop<7, L>(hm, P, q, targets, v, status, B); return true;
  }
  return false;
}

struct Args {
  const float* q; const float* targets; float* v; int32_t* status; float* H; float* c; float* h;
  float* e; float* J; int task_index; int task_k; float* oMf; float* com; float* Jf; int jac_frame;
  float* G; float* hG; float* E; float* f; float* lo; float* hi;
};

// make_dev_problem + host image of the optional parts, wired through P.ext
std::string make_problem(const pk::HostModel& hm, const PkProblemDesc* prob, pk::DevProblem* P, pk::HostExtras* hx) {
  const std::string e = pk::make_dev_problem(hm, prob, P, hx);
  if (e.empty() && hx->present) P->ext = &hx->X;
  return e;
}

void run_generic(const pk::HostModel& hm, const pk::DevProblem& P, const Args& A, int64_t B) {
  const pk::DevModel M = hm.host_view();
  const int nv = M.nv;
  static thread_local pk::Generic<PK_MAX_JOINTS, PK_MAX_NV> G;
  for (int64_t i = 0; i < B; ++i) {
    pk::GenericOut out;
    out.v = A.v ? A.v + i * nv : nullptr;
    out.status = A.status ? A.status + i : nullptr;
    out.H = A.H ? A.H + i * nv * nv : nullptr;
    out.c = A.c ? A.c + i * nv : nullptr;
    out.h = A.h ? A.h + i * 4 * nv : nullptr;
    out.e = A.e ? A.e + i * A.task_k : nullptr;
    out.J = A.J ? A.J + i * A.task_k * nv : nullptr;
    out.task_index = A.task_index;
    out.oMf = A.oMf ? A.oMf + i * M.nframes * 12 : nullptr;
    out.com = A.com ? A.com + i * 3 : nullptr;
    out.Jf = A.Jf ? A.Jf + i * 6 * nv : nullptr;
    out.jac_frame = A.jac_frame;
    out.G = A.G ? A.G + i * PK_MAX_INEQ_ROWS * nv : nullptr;
    out.hG = A.hG ? A.hG + i * PK_MAX_INEQ_ROWS : nullptr;
    out.E = A.E ? A.E + i * PK_MAX_EQ_ROWS * nv : nullptr;
    out.f = A.f ? A.f + i * PK_MAX_EQ_ROWS : nullptr;
    out.lo = A.lo ? A.lo + i * nv : nullptr;
    out.hi = A.hi ? A.hi + i * nv : nullptr;
    G.step(M, P, A.q + i * M.nq, A.targets ? A.targets + i * (int64_t)P.target_stride : nullptr, out);
  }
}
}  // namespace

extern "C" {

const char* hs_last_error(void) { return g_err.c_str(); }

int hs_model_create(const PkModelDesc* d, void** out) {
  pk::HostModel* m = new pk::HostModel();
  const std::string e = pk::build_host_model(d, m);
  if (!e.empty()) { delete m; return fail(e); }
  *out = m;
  return 0;
}
void hs_model_destroy(void* m) { delete (pk::HostModel*)m; }

// path: 0 auto (same selection as the CUDA library), 1 force the general path, 2 force the tree kernel body
int hs_solve_ik(void* model, const PkProblemDesc* prob, const float* q, const float* targets, float* v,
                int32_t* status, int64_t B, int path, int* used_chain) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  pk::DevProblem P;
  pk::HostExtras hx;
  const std::string e = make_problem(hm, prob, &P, &hx);
  if (!e.empty()) return fail(e);
  const bool chain = path == 0 && pk::chain_eligible(hm, P, hx.present && !hx.box_only());
  if (used_chain) *used_chain = chain ? 1 : 0;
  // path 11 / 12 / 14 / 18: the sub-warp chain kernel body with 1 / 2 / 4 / 8 lanes per instance
  if (path > 10) {
    if (!pk::chain_eligible(hm, P, hx.present && !hx.box_only())) return fail("problem does not fit the chain kernel");
    if (used_chain) *used_chain = 1;
    bool ok = false;
    switch (path - 10) {
      case 1: ok = run_coop_nj<1>(hm, P, q, targets, v, status, B); break;
      case 2: ok = run_coop_nj<2>(hm, P, q, targets, v, status, B); break;
      case 4: ok = run_coop_nj<4>(hm, P, q, targets, v, status, B); break;
      case 8: ok = run_coop_nj<8>(hm, P, q, targets, v, status, B); break;
    }
    return ok ? 0 : fail("unsupported lanes per instance");
  }
  if (chain) {
    switch (hm.njoints) {
      case 2: run_chain<2>(hm, P, q, targets, v, status, B); return 0;
      case 3: run_chain<3>(hm, P, q, targets, v, status, B); return 0;
      case 4: run_chain<4>(hm, P, q, targets, v, status, B); return 0;
      case 5: run_chain<5>(hm, P, q, targets, v, status, B); return 0;
      case 6: run_chain<6>(hm, P, q, targets, v, status, B); return 0;
      case 7: run_chain<7>(hm, P, q, targets, v, status, B); return 0;
    }
  }
  if (path == 2 || (path == 0 && !chain)) {
    bool ok = false;
    const pk::TreePlan L = pk::make_tree_plan(hm, P, &ok, hx.present ? &hx.X : nullptr);
    if (ok) {
      if (used_chain) *used_chain = 2;
      const pk::DevModel M = hm.host_view();
      std::vector<float> W(L.words + 8);
      float* Wp = W.data();
      if (((uintptr_t)Wp & 7) != 0) ++Wp;  // the dual method keeps doubles in the workspace
      for (int64_t i = 0; i < B; ++i)
        pk::TreeStep::run(M, P, L, q + i * L.nq, targets ? targets + i * (int64_t)L.stride : nullptr, Wp,
                          v + i * L.nv, status ? status + i : nullptr);
      return 0;
    }
    if (path == 2) return fail("problem does not fit the tree kernel");
  }
  Args A{};
  A.q = q; A.targets = targets; A.v = v; A.status = status; A.task_index = -1;
  run_generic(hm, P, A, B);
  return 0;
}

int hs_build_ik(void* model, const PkProblemDesc* prob, const float* q, const float* targets, float* H, float* c,
                float* h, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  pk::DevProblem P;
  pk::HostExtras hx;
  const std::string e = make_problem(hm, prob, &P, &hx);
  if (!e.empty()) return fail(e);
  Args A{};
  A.q = q; A.targets = targets; A.H = H; A.c = c; A.h = h; A.task_index = -1;
  run_generic(hm, P, A, B);
  return 0;
}

int hs_constraint_rows(void* model, const PkProblemDesc* prob, const float* q, const float* targets, float* G,
                       float* hG, float* E, float* f, float* lo, float* hi, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  pk::DevProblem P;
  pk::HostExtras hx;
  const std::string e = make_problem(hm, prob, &P, &hx);
  if (!e.empty()) return fail(e);
  Args A{};
  A.q = q; A.targets = targets; A.G = G; A.hG = hG; A.E = E; A.f = f; A.lo = lo; A.hi = hi; A.task_index = -1;
  run_generic(hm, P, A, B);
  return 0;
}

int hs_task_terms(void* model, const PkProblemDesc* prob, int task_index, const float* q, const float* targets,
                  float* eo, float* J, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  pk::DevProblem P;
  pk::HostExtras hx;
  const std::string e = make_problem(hm, prob, &P, &hx);
  if (!e.empty()) return fail(e);
  if (task_index < 0 || task_index >= P.ntasks) return fail("task_index out of range");
  Args A{};
  A.q = q; A.targets = targets; A.e = eo; A.J = J; A.task_index = task_index;
  const int type = P.tasks[task_index].type;
  A.task_k = type == PK_TASK_COM ? 3
             : type == PK_TASK_LINEAR ? P.tasks[task_index].rows
                                      : (pk::is_diag_task(type) ? hm.nv - (hm.free_flyer ? 6 : 0) : 6);
  run_generic(hm, P, A, B);
  return 0;
}

int hs_forward_kinematics(void* model, const float* q, float* oMf, float* com, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  pk::DevProblem P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < PK_MAX_NV; ++i) { P.chk_lo[i] = -INFINITY; P.chk_hi[i] = INFINITY; }
  Args A{};
  A.q = q; A.oMf = oMf; A.com = com; A.task_index = -1;
  run_generic(hm, P, A, B);
  return 0;
}

int hs_frame_jacobian(void* model, int frame, const float* q, float* J, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  if (frame < 0 || frame >= hm.nframes) return fail("frame index out of range");
  pk::DevProblem P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < PK_MAX_NV; ++i) { P.chk_lo[i] = -INFINITY; P.chk_hi[i] = INFINITY; }
  Args A{};
  A.q = q; A.Jf = J; A.jac_frame = frame; A.task_index = -1;
  run_generic(hm, P, A, B);
  return 0;
}

// q (+) v dt with the body of integrate_kernel (pk_math.cuh)
int hs_integrate(void* model, const float* q, const float* v, float dt, float* q_out, int64_t B) {
  const pk::HostModel& hm = *(pk::HostModel*)model;
  for (int64_t i = 0; i < B; ++i)
    pk::integrate_configuration(hm.nq, hm.free_flyer, q + i * hm.nq, v + i * hm.nv, dt, q_out + i * hm.nq);
  return 0;
}

// Direct access to the dual active-set QP (pk_dualqp.cuh) for unit tests:
//   min 1/2 |A x + b|^2 + 1/2 sum (d_i x_i + beta_i)^2,  lo <= x <= hi,  G x <= h,  E x = f
// A[K][n], G[p][n], E[meq][n] row-major; returns the solver's status bits.
int hs_dual_qp(int K, int n, int p, int meq, const float* A, const float* b, const float* d, const float* beta,
               const float* lo, const float* hi, const float* G, const float* h, const float* E, const float* f,
               float* x) {
  using QP = pk::DualQP<pk::kGenericMaxRows, PK_MAX_NV, PK_MAX_INEQ_ROWS, PK_MAX_EQ_ROWS>;
  if (K > pk::kGenericMaxRows || n > PK_MAX_NV || p > PK_MAX_INEQ_ROWS || meq > PK_MAX_EQ_ROWS) return -1;
  static thread_local float Aa[pk::kGenericMaxRows][PK_MAX_NV], Ga[PK_MAX_INEQ_ROWS][PK_MAX_NV], Ea[PK_MAX_EQ_ROWS][PK_MAX_NV];
  for (int r = 0; r < K; ++r) for (int c = 0; c < n; ++c) Aa[r][c] = A[r * n + c];
  for (int r = 0; r < p; ++r) for (int c = 0; c < n; ++c) Ga[r][c] = G[r * n + c];
  for (int r = 0; r < meq; ++r) for (int c = 0; c < n; ++c) Ea[r][c] = E[r * n + c];
  QP::Problem P{Aa, b, d, beta, lo, hi, Ga, h, Ea, f, K, n, p, meq};
  float xs[PK_MAX_NV];
  const int st = QP::run(P, xs);
  for (int i = 0; i < n; ++i) x[i] = xs[i];
  return st;
}

// Direct access to the warp-cooperative dual QP (pk_treedual.cuh) for unit tests; same
// problem as hs_dual_qp without equalities.
int hs_tree_dual_qp(int K, int n, int p, int meq, const float* A, const float* b, const float* d, const float* beta,
                    const float* lo, const float* hi, const float* G, const float* h, const float* E, const float* f,
                    float* x) {
  pk::TreePlan L;
  memset(&L, 0, sizeof(L));
  L.nj = 1; L.nq = n; L.nv = n; L.K = K; L.p = p; L.meq = meq; L.npairs = 0;
  pk::tree_layout(L, 1, 0);
  std::vector<float> W(L.words + 8, 0.f);
  float* Wp = W.data();
  if (((uintptr_t)Wp & 7) != 0) ++Wp;  // doubles in the workspace
  for (int r = 0; r < K; ++r) {
    for (int c = 0; c < n; ++c) Wp[L.o_A + r * L.lda + c] = A[r * n + c];
    Wp[L.o_b + r] = b[r];
  }
  for (int r = 0; r < p; ++r) {
    for (int c = 0; c < n; ++c) Wp[L.o_G + r * L.lda + c] = G[r * n + c];
    Wp[L.o_hg + r] = h[r];
  }
  for (int r = 0; r < meq; ++r) {
    for (int c = 0; c < n; ++c) Wp[L.o_E + r * L.lda + c] = E[r * n + c];
    Wp[L.o_fe + r] = f[r];
  }
  for (int i = 0; i < n; ++i) {
    Wp[L.o_d + i] = d[i]; Wp[L.o_beta + i] = beta[i]; Wp[L.o_lo + i] = lo[i]; Wp[L.o_hi + i] = hi[i];
  }
  const int st = pk::TreeDual::solve(Wp, L);
  for (int i = 0; i < n; ++i) x[i] = Wp[L.o_x + i];
  return st;
}
}
