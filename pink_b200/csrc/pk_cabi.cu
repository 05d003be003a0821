// C ABI of pink_b200 (include/pink_b200.h): model tables, problem marshalling,
// kernel selection and launches.  No torch types; built with plain nvcc into
// pink_b200/libpink_b200.so.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pink_b200.h"
#include "pk_chain.cuh"
#include "pk_generic.cuh"
#include "pk_marshal.hpp"

namespace {

thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

int fail(const std::string& msg) {
  g_error = msg;
  return 1;
}

#define PK_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t err__ = (expr);                                                         \
    if (err__ != cudaSuccess)                                                           \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(err__));              \
  } while (0)

}  // namespace

// --------------------------------------------------------------------------------------
// model
// --------------------------------------------------------------------------------------

struct PkModel {
  int device = 0;
  pk::HostModel hm;
  int njoints = 0, free_flyer = 0, nq = 0, nv = 0, nframes = 0;
  void* dev_buf = nullptr;
  pk::DevModel dev{};
  // staging of the host entry point
  std::mutex mu;
  float* st_q = nullptr;
  float* st_t = nullptr;
  float* st_v = nullptr;
  int32_t* st_s = nullptr;
  int64_t st_cap = 0;
  int st_tstride = 0;
  cudaStream_t st_streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t st_fork = nullptr;
  cudaEvent_t st_join[3] = {nullptr, nullptr, nullptr};
};

namespace {

template <typename T>
size_t align_up(size_t off) {
  return (off + alignof(T) - 1) / alignof(T) * alignof(T);
}

}  // namespace

extern "C" int pk_abi_version(void) { return PK_ABI_VERSION; }
extern "C" const char* pk_last_error(void) { return g_error.c_str(); }
extern "C" int64_t pk_launch_count(void) { return g_launches.load(); }

extern "C" int pk_model_create(const PkModelDesc* d, int device, PkModel** out) {
  if (!d || !out) return fail("pk_model_create: null argument");
  PkModel* m = new PkModel();
  const std::string err_msg = pk::build_host_model(d, &m->hm);
  if (!err_msg.empty()) {
    delete m;
    return fail("pk_model_create: " + err_msg);
  }
  m->device = device;
  const int nj = m->hm.njoints;
  const int ff = m->hm.free_flyer;
  m->njoints = nj;
  m->free_flyer = ff;
  m->nq = m->hm.nq;
  m->nv = m->hm.nv;
  m->nframes = m->hm.nframes;

  // one device buffer holding every table
  size_t off = 0;
  auto reserve = [&](size_t bytes, size_t align) {
    off = (off + align - 1) / align * align;
    size_t at = off;
    off += bytes;
    return at;
  };
  const size_t o_anc = reserve(sizeof(uint64_t) * (nj + 2), 8);
  const size_t o_parent = reserve(sizeof(int) * std::max(nj, 1), 4);
  const size_t o_jtype = reserve(sizeof(int) * std::max(nj, 1), 4);
  const size_t o_jX = reserve(sizeof(float) * 12 * std::max(nj, 1), 16);
  const size_t o_axis = reserve(sizeof(float) * 3 * std::max(nj, 1), 4);
  const size_t o_fb = reserve(sizeof(int) * std::max(d->nframes, 1), 4);
  const size_t o_fX = reserve(sizeof(float) * 12 * std::max(d->nframes, 1), 16);
  const size_t o_mass = reserve(sizeof(float) * (nj + 1), 4);
  const size_t o_com = reserve(sizeof(float) * 3 * (nj + 1), 4);
  std::vector<char> host(off, 0);
  memcpy(host.data() + o_anc, m->hm.anc.data(), sizeof(uint64_t) * (nj + 2));
  if (nj) {
    memcpy(host.data() + o_parent, m->hm.parent.data(), sizeof(int) * nj);
    memcpy(host.data() + o_jtype, m->hm.jtype.data(), sizeof(int) * nj);
    memcpy(host.data() + o_jX, m->hm.jX.data(), sizeof(float) * 12 * nj);
    memcpy(host.data() + o_axis, m->hm.axis.data(), sizeof(float) * 3 * nj);
  }
  if (d->nframes) {
    memcpy(host.data() + o_fb, m->hm.frame_body.data(), sizeof(int) * d->nframes);
    memcpy(host.data() + o_fX, m->hm.fX.data(), sizeof(float) * 12 * d->nframes);
  }
  memcpy(host.data() + o_mass, m->hm.mass.data(), sizeof(float) * (nj + 1));
  memcpy(host.data() + o_com, m->hm.com.data(), sizeof(float) * 3 * (nj + 1));

  cudaError_t err = cudaSetDevice(device);
  if (err == cudaSuccess) err = cudaMalloc(&m->dev_buf, off);
  if (err == cudaSuccess) err = cudaMemcpy(m->dev_buf, host.data(), off, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    const std::string msg = std::string("pk_model_create: ") + cudaGetErrorString(err);
    if (m->dev_buf) cudaFree(m->dev_buf);
    delete m;
    return fail(msg);
  }
  char* base = (char*)m->dev_buf;
  m->dev.njoints = nj;
  m->dev.free_flyer = ff;
  m->dev.nq = m->nq;
  m->dev.nv = m->nv;
  m->dev.nframes = d->nframes;
  m->dev.anc = (const uint64_t*)(base + o_anc);
  m->dev.parent = (const int*)(base + o_parent);
  m->dev.jtype = (const int*)(base + o_jtype);
  m->dev.jX = (const float*)(base + o_jX);
  m->dev.axis = (const float*)(base + o_axis);
  m->dev.frame_body = (const int*)(base + o_fb);
  m->dev.fX = (const float*)(base + o_fX);
  m->dev.mass = (const float*)(base + o_mass);
  m->dev.com = (const float*)(base + o_com);
  m->dev.total_mass = m->hm.total_mass;
  *out = m;
  return 0;
}

extern "C" void pk_model_destroy(PkModel* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->dev_buf) cudaFree(m->dev_buf);
  if (m->st_q) cudaFree(m->st_q);
  if (m->st_t) cudaFree(m->st_t);
  if (m->st_v) cudaFree(m->st_v);
  if (m->st_s) cudaFree(m->st_s);
  for (int i = 0; i < 3; ++i) {
    if (m->st_streams[i]) cudaStreamDestroy(m->st_streams[i]);
    if (m->st_join[i]) cudaEventDestroy(m->st_join[i]);
  }
  if (m->st_fork) cudaEventDestroy(m->st_fork);
  delete m;
}

// --------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------

namespace pk {

template <int NJ, int NFT>
__global__ void __launch_bounds__(128) ik_chain_kernel(const __grid_constant__ ChainParams<NJ> P,
                                                       const float* __restrict__ q,
                                                       const float* __restrict__ targets,
                                                       float* __restrict__ v, int32_t* __restrict__ status,
                                                       int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float qi[NJ], vi[NJ];
  const float* qrow = q + i * NJ;
  if constexpr (NJ % 2 == 0) {
#pragma unroll
    for (int k = 0; k < NJ / 2; ++k) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(qrow) + k);
      qi[2 * k] = t.x;
      qi[2 * k + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NJ; ++k) qi[k] = __ldg(qrow + k);
  }
  int st;
  ik_step_chain<NJ, NFT>(P, qi, targets + i * (int64_t)P.target_stride, vi, st);
  float* vrow = v + i * NJ;
  if constexpr (NJ % 2 == 0) {
#pragma unroll
    for (int k = 0; k < NJ / 2; ++k) reinterpret_cast<float2*>(vrow)[k] = make_float2(vi[2 * k], vi[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < NJ; ++k) vrow[k] = vi[k];
  }
  if (status) status[i] = st;
}

// General path: one instance per thread, per-thread arrays in local memory.
struct GenericArgs {
  const float* q;
  const float* targets;
  float* v;
  int32_t* status;
  float* H;
  float* c;
  float* h;
  float* e;
  float* J;
  int task_index;
  int task_k;
  float* oMf;
  float* com;
  float* Jf;
  int jac_frame;
};

template <int NJMAX, int NVMAX>
__global__ void __launch_bounds__(64) ik_generic_kernel(const DevModel M, const __grid_constant__ DevProblem P,
                                                        const GenericArgs A, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int nv = M.nv;
  GenericOut out;
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
  Generic<NJMAX, NVMAX> G;
  G.step(M, P, A.q + i * M.nq, A.targets ? A.targets + i * (int64_t)P.target_stride : nullptr, out);
}

// q (+) v dt
__global__ void integrate_kernel(int nq, int nv, int free_flyer, const float* __restrict__ q,
                                 const float* __restrict__ v, float dt, float* __restrict__ qo, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* qi = q + i * nq;
  const float* vi = v + i * nv;
  float* o = qo + i * nq;
  int rq = 0, rv = 0;
  if (free_flyer) {
    rq = 7; rv = 6;
    // M <- M exp6(v dt): translation += R V(w) vlin, quaternion <- quaternion * exp(w/2)
    const M3 R = quat_to_matrix(qi[3], qi[4], qi[5], qi[6]);
    const V3 vl = dt * v3(vi[0], vi[1], vi[2]);
    const V3 w = dt * v3(vi[3], vi[4], vi[5]);
    const float x = dot(w, w);
    const float th = sqrtf(x);
    float b, c;  // (1 - cos)/th^2, (th - sin)/th^3
    if (th < 1e-2f) {
      b = 0.5f - x / 24.f + x * x / 720.f;
      c = 1.f / 6.f - x / 120.f + x * x / 5040.f;
    } else {
      float s, co;
      sincos_f(th, &s, &co);
      b = (1.f - co) / x;
      c = (th - s) / (x * th);
    }
    const V3 wv = cross(w, vl);
    const V3 t = vl + b * wv + c * cross(w, wv);
    const V3 p = mul(R, t);
    o[0] = qi[0] + p.x; o[1] = qi[1] + p.y; o[2] = qi[2] + p.z;
    float sh, ch;
    sincos_f(0.5f * th, &sh, &ch);
    const float k = th < 1e-4f ? 0.5f : sh / th;
    const float dx = k * w.x, dy = k * w.y, dz = k * w.z, dw = ch;
    const float n0 = rsqrtf(qi[3] * qi[3] + qi[4] * qi[4] + qi[5] * qi[5] + qi[6] * qi[6]);
    const float ax = qi[3] * n0, ay = qi[4] * n0, az = qi[5] * n0, aw = qi[6] * n0;
    float rx = aw * dx + ax * dw + ay * dz - az * dy;
    float ry = aw * dy - ax * dz + ay * dw + az * dx;
    float rz = aw * dz + ax * dy - ay * dx + az * dw;
    float rw = aw * dw - ax * dx - ay * dy - az * dz;
    const float n1 = rsqrtf(rx * rx + ry * ry + rz * rz + rw * rw);
    o[3] = rx * n1; o[4] = ry * n1; o[5] = rz * n1; o[6] = rw * n1;
  }
  for (int j = 0; j < nq - rq; ++j) o[rq + j] = fmaf(vi[rv + j], dt, qi[rq + j]);
}

}  // namespace pk

namespace {

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int NJ>
int launch_chain(const PkModel* m, const pk::DevProblem& P, const float* q, const float* targets, float* v,
                 int32_t* status, int64_t B, cudaStream_t stream) {
  pk::ChainParams<NJ> C;
  pk::make_chain_params<NJ>(m->hm, P, &C);
  static const int block = std::min(128, std::max(32, env_int("PK_CHAIN_BLOCK", 64)));
  const int64_t grid = (B + block - 1) / block;
  switch (C.n_frame_tasks) {
    case 0: pk::ik_chain_kernel<NJ, 0><<<(unsigned)grid, block, 0, stream>>>(C, q, targets, v, status, B); break;
    case 1: pk::ik_chain_kernel<NJ, 1><<<(unsigned)grid, block, 0, stream>>>(C, q, targets, v, status, B); break;
    default: pk::ik_chain_kernel<NJ, 2><<<(unsigned)grid, block, 0, stream>>>(C, q, targets, v, status, B); break;
  }
  g_launches.fetch_add(1);
  PK_CUDA(cudaGetLastError());
  return 0;
}

int launch_generic(const PkModel* m, const pk::DevProblem& P, const pk::GenericArgs& A, int64_t B,
                   cudaStream_t stream) {
  const int block = 64;
  const int64_t grid = (B + block - 1) / block;
  if (m->nv <= 8 && m->njoints <= 8)
    pk::ik_generic_kernel<8, 8><<<(unsigned)grid, block, 0, stream>>>(m->dev, P, A, B);
  else if (m->nv <= 36 && m->njoints <= 30)
    pk::ik_generic_kernel<30, 36><<<(unsigned)grid, block, 0, stream>>>(m->dev, P, A, B);
  else
    pk::ik_generic_kernel<PK_MAX_JOINTS, PK_MAX_NV><<<(unsigned)grid, block, 0, stream>>>(m->dev, P, A, B);
  g_launches.fetch_add(1);
  PK_CUDA(cudaGetLastError());
  return 0;
}

int solve_device(const PkModel* m, const pk::DevProblem& P, const float* q, const float* targets, float* v,
                 int32_t* status, int64_t B, cudaStream_t stream) {
  if (B == 0) return 0;
  static const int force_generic = env_int("PK_FORCE_GENERIC", 0);
  if (!force_generic && pk::chain_eligible(m->hm, P)) {
    switch (m->njoints) {
      case 2: return launch_chain<2>(m, P, q, targets, v, status, B, stream);
      case 3: return launch_chain<3>(m, P, q, targets, v, status, B, stream);
      case 4: return launch_chain<4>(m, P, q, targets, v, status, B, stream);
      case 5: return launch_chain<5>(m, P, q, targets, v, status, B, stream);
      case 6: return launch_chain<6>(m, P, q, targets, v, status, B, stream);
      case 7: return launch_chain<7>(m, P, q, targets, v, status, B, stream);
      default: break;
    }
  }
  pk::GenericArgs A{};
  A.q = q;
  A.targets = targets;
  A.v = v;
  A.status = status;
  A.task_index = -1;
  return launch_generic(m, P, A, B, stream);
}

int check_common(const PkModel* m, const void* q, int64_t B) {
  if (!m) return fail("null model");
  if (B < 0) return fail("negative batch size");
  if (B > 0 && !q) return fail("null q");
  if (B > (int64_t)2147483647 * 32) return fail("batch too large");
  return 0;
}

}  // namespace

// --------------------------------------------------------------------------------------
// entry points
// --------------------------------------------------------------------------------------

extern "C" int pk_solve_ik_batched(const PkModel* m, const PkProblemDesc* prob, const float* q,
                                   const float* targets, float* v, int32_t* status, int64_t B, void* stream) {
  if (check_common(m, q, B)) return 1;
  if (B > 0 && !v) return fail("null v");
  pk::DevProblem P;
  {
    const std::string perr = pk::make_dev_problem(m->hm, prob, &P);
    if (!perr.empty()) return fail(perr);
  }
  if (B > 0 && P.target_stride > 0 && !targets) return fail("null targets");
  return solve_device(m, P, q, targets, v, status, B, (cudaStream_t)stream);
}

extern "C" int pk_solve_ik_batched_host(PkModel* m, const PkProblemDesc* prob, const float* q_host,
                                        const float* targets_host, float* v_host, int32_t* status_host,
                                        int64_t B, void* stream_) {
  if (check_common(m, q_host, B)) return 1;
  if (B > 0 && !v_host) return fail("null v");
  pk::DevProblem P;
  {
    const std::string perr = pk::make_dev_problem(m->hm, prob, &P);
    if (!perr.empty()) return fail(perr);
  }
  if (B > 0 && P.target_stride > 0 && !targets_host) return fail("null targets");
  if (B == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  std::lock_guard<std::mutex> lock(m->mu);
  PK_CUDA(cudaSetDevice(m->device));
  const int ts = P.target_stride;
  if (B > m->st_cap || ts > m->st_tstride) {
    // (re)size staging; happens on the first call or when the batch grows
    PK_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 3; ++i)
      if (m->st_streams[i]) PK_CUDA(cudaStreamSynchronize(m->st_streams[i]));
    if (m->st_q) cudaFree(m->st_q);
    if (m->st_t) cudaFree(m->st_t);
    if (m->st_v) cudaFree(m->st_v);
    if (m->st_s) cudaFree(m->st_s);
    m->st_q = m->st_t = m->st_v = nullptr;
    m->st_s = nullptr;
    const int64_t cap = std::max<int64_t>(B, m->st_cap);
    const int tcap = std::max(ts, m->st_tstride);
    PK_CUDA(cudaMalloc(&m->st_q, sizeof(float) * cap * m->nq));
    PK_CUDA(cudaMalloc(&m->st_t, sizeof(float) * cap * std::max(tcap, 1)));
    PK_CUDA(cudaMalloc(&m->st_v, sizeof(float) * cap * m->nv));
    PK_CUDA(cudaMalloc(&m->st_s, sizeof(int32_t) * cap));
    m->st_cap = cap;
    m->st_tstride = tcap;
  }
  if (!m->st_fork) {
    PK_CUDA(cudaEventCreateWithFlags(&m->st_fork, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) {
      PK_CUDA(cudaStreamCreateWithFlags(&m->st_streams[i], cudaStreamNonBlocking));
      PK_CUDA(cudaEventCreateWithFlags(&m->st_join[i], cudaEventDisableTiming));
    }
  }
  // chunks round-robin over three internal streams: the H2D of chunk k+1, the
  // kernel of chunk k and the D2H of chunk k-1 overlap (PCIe is full duplex)
  static const int64_t chunk_env = env_int("PK_HOST_CHUNK", 16384);
  const int64_t chunk = std::max<int64_t>(1024, chunk_env);
  const int64_t nchunks = (B + chunk - 1) / chunk;
  const int nstreams = (int)std::min<int64_t>(3, nchunks);
  PK_CUDA(cudaEventRecord(m->st_fork, stream));
  for (int i = 0; i < nstreams; ++i) PK_CUDA(cudaStreamWaitEvent(m->st_streams[i], m->st_fork, 0));
  for (int64_t k = 0; k < nchunks; ++k) {
    cudaStream_t s = m->st_streams[k % nstreams];
    const int64_t b0 = k * chunk;
    const int64_t nb = std::min(chunk, B - b0);
    PK_CUDA(cudaMemcpyAsync(m->st_q + b0 * m->nq, q_host + b0 * m->nq, sizeof(float) * nb * m->nq,
                            cudaMemcpyHostToDevice, s));
    if (ts > 0)
      PK_CUDA(cudaMemcpyAsync(m->st_t + b0 * ts, targets_host + b0 * ts, sizeof(float) * nb * ts,
                              cudaMemcpyHostToDevice, s));
    if (solve_device(m, P, m->st_q + b0 * m->nq, m->st_t + b0 * ts, m->st_v + b0 * m->nv, m->st_s + b0, nb, s))
      return 1;
    PK_CUDA(cudaMemcpyAsync(v_host + b0 * m->nv, m->st_v + b0 * m->nv, sizeof(float) * nb * m->nv,
                            cudaMemcpyDeviceToHost, s));
    if (status_host)
      PK_CUDA(cudaMemcpyAsync(status_host + b0, m->st_s + b0, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < nstreams; ++i) {
    PK_CUDA(cudaEventRecord(m->st_join[i], m->st_streams[i]));
    PK_CUDA(cudaStreamWaitEvent(stream, m->st_join[i], 0));
  }
  return 0;
}

extern "C" int pk_build_ik_batched(const PkModel* m, const PkProblemDesc* prob, const float* q,
                                   const float* targets, float* H, float* c, float* h, int64_t B, void* stream) {
  if (check_common(m, q, B)) return 1;
  pk::DevProblem P;
  {
    const std::string perr = pk::make_dev_problem(m->hm, prob, &P);
    if (!perr.empty()) return fail(perr);
  }
  if (B == 0) return 0;
  if (!H) return fail("null H");
  pk::GenericArgs A{};
  A.q = q;
  A.targets = targets;
  A.H = H;
  A.c = c;
  A.h = h;
  A.task_index = -1;
  return launch_generic(m, P, A, B, (cudaStream_t)stream);
}

extern "C" int pk_task_terms_batched(const PkModel* m, const PkProblemDesc* prob, int32_t task_index,
                                     const float* q, const float* targets, float* e, float* J, int64_t B,
                                     void* stream) {
  if (check_common(m, q, B)) return 1;
  pk::DevProblem P;
  {
    const std::string perr = pk::make_dev_problem(m->hm, prob, &P);
    if (!perr.empty()) return fail(perr);
  }
  if (task_index < 0 || task_index >= P.ntasks) return fail("task_index out of range");
  if (B == 0) return 0;
  pk::GenericArgs A{};
  A.q = q;
  A.targets = targets;
  A.e = e;
  A.J = J;
  A.task_index = task_index;
  const int type = P.tasks[task_index].type;
  A.task_k = type == PK_TASK_COM ? 3 : (type == PK_TASK_POSTURE ? m->nv - (m->free_flyer ? 6 : 0) : 6);
  // H must be accumulated for the task loop to run; give the kernel no v/H outputs
  // but keep ntasks > 0 so the loop executes
  return launch_generic(m, P, A, B, (cudaStream_t)stream);
}

extern "C" int pk_forward_kinematics_batched(const PkModel* m, const float* q, float* oMf, float* com,
                                             int64_t B, void* stream) {
  if (check_common(m, q, B)) return 1;
  if (B == 0) return 0;
  pk::DevProblem P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < PK_MAX_NV; ++i) { P.chk_lo[i] = -INFINITY; P.chk_hi[i] = INFINITY; }
  pk::GenericArgs A{};
  A.q = q;
  A.oMf = oMf;
  A.com = com;
  A.task_index = -1;
  return launch_generic(m, P, A, B, (cudaStream_t)stream);
}

extern "C" int pk_frame_jacobian_batched(const PkModel* m, int32_t frame, const float* q, float* J, int64_t B,
                                         void* stream) {
  if (check_common(m, q, B)) return 1;
  if (frame < 0 || frame >= m->nframes) return fail("frame index out of range");
  if (B == 0) return 0;
  pk::DevProblem P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < PK_MAX_NV; ++i) { P.chk_lo[i] = -INFINITY; P.chk_hi[i] = INFINITY; }
  pk::GenericArgs A{};
  A.q = q;
  A.Jf = J;
  A.jac_frame = frame;
  A.task_index = -1;
  return launch_generic(m, P, A, B, (cudaStream_t)stream);
}

extern "C" int pk_integrate_batched(const PkModel* m, const float* q, const float* v, float dt, float* q_out,
                                    int64_t B, void* stream) {
  if (check_common(m, q, B)) return 1;
  if (B == 0) return 0;
  if (!v || !q_out) return fail("null v or q_out");
  const int block = 128;
  const int64_t grid = (B + block - 1) / block;
  pk::integrate_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(m->nq, m->nv, m->free_flyer, q, v, dt,
                                                                             q_out, B);
  g_launches.fetch_add(1);
  PK_CUDA(cudaGetLastError());
  return 0;
}
