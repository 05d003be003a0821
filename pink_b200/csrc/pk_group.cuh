// Sub-warp cooperative layer of the chain kernel (pk_coop.cuh): one robot instance is
// handled by a GROUP of L consecutive lanes of a warp (L = 1, 2, 4, 8; 32 / L instances per
// warp).  Same idea as pk_warp.cuh, one level down: the algorithm is a sequence of
// lane-parallel blocks (PK_GLANES) that exchange data only through the primitives below;
// the same source compiles for the host, where a group is a plain loop over L lanes
// (tests/hostsim runs it on the CPU: test harness only).
//
// Rules that keep both builds equivalent:
//  * per-lane state lives in a GVar<T, L> (a register set on the device, an array of L on
//    the host); inside PK_GLANES(G, h) a lane touches only its own entry [h];
//  * a lane reads another lane's state only through g_get / g_sum / g_any / g_argmin, and
//    only state written in an EARLIER PK_GLANES block;
//  * values that are identical on every lane of the group (results of the primitives,
//    anything computed from them or from per-instance inputs) are ordinary variables
//    outside the blocks; branches on them are group-uniform, so the shuffles below - which
//    name only the group's lanes in their mask - are safe when different groups of a warp
//    take different paths.
#pragma once

#include "pk_math.cuh"

namespace pk {

#if defined(__CUDA_ARCH__)

template <int L>
struct Group {
  int h;          // this lane's index in its group
  unsigned mask;  // the group's lanes within the warp
  __device__ __forceinline__ Group() {
    const unsigned lane = threadIdx.x & 31u;
    h = (L == 1) ? 0 : (int)(lane & (unsigned)(L - 1));
    mask = (L >= 32) ? 0xffffffffu : (((1u << L) - 1u) << (lane & ~(unsigned)(L - 1)));
  }
};

#define PK_GLANES(G, h) for (int h = (G).h, pk_gonce_##h = 1; pk_gonce_##h; pk_gonce_##h = 0)

template <class T, int L>
struct GVar {
  T v;
  __device__ __forceinline__ T& operator[](int) { return v; }
  __device__ __forceinline__ const T& operator[](int) const { return v; }
};

// f(state of lane `src`) on every lane (src may differ per lane: shifts)
template <int L, class T, class F>
__device__ __forceinline__ float g_get(const Group<L>& G, const GVar<T, L>& S, int src, F f) {
  const float own = f(S.v);
  if (L == 1) return own;
  return __shfl_sync(G.mask, own, src, L);
}
template <int L, class T, class F>
__device__ __forceinline__ int g_get_int(const Group<L>& G, const GVar<T, L>& S, int src, F f) {
  const int own = f(S.v);
  if (L == 1) return own;
  return __shfl_sync(G.mask, own, src, L);
}
// sum over the group of f(lane state), on every lane (xor butterfly)
template <int L, class T, class F>
__device__ __forceinline__ float g_sum(const Group<L>& G, const GVar<T, L>& S, F f) {
  float s = f(S.v);
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) s += __shfl_xor_sync(G.mask, s, o, L);
  return s;
}
template <int L, class T, class F>
__device__ __forceinline__ bool g_any(const Group<L>& G, const GVar<T, L>& S, F f) {
  const bool p = f(S.v);
  if (L == 1) return p;
  return __any_sync(G.mask, p);
}
// smallest f.value over the group and the integer tag that comes with it (ties: smaller tag)
template <int L, class T, class FV, class FI>
__device__ __forceinline__ void g_argmin(const Group<L>& G, const GVar<T, L>& S, FV fv, FI fi, float& best, int& tag) {
  float s = fv(S.v);
  int i = fi(S.v);
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) {
    const float so = __shfl_xor_sync(G.mask, s, o, L);
    const int io = __shfl_xor_sync(G.mask, i, o, L);
    if (so < s || (so == s && io < i)) { s = so; i = io; }
  }
  best = s;
  tag = i;
}

#else  // host emulation: a group is a loop over L lanes

template <int L>
struct Group {};

template <int L>
inline constexpr int group_size(const Group<L>&) { return L; }

#define PK_GLANES(G, h) for (int h = 0; h < pk::group_size(G); ++h)

template <class T, int L>
struct GVar {
  T v[L];
  T& operator[](int h) { return v[h]; }
  const T& operator[](int h) const { return v[h]; }
};

template <int L, class T, class F>
inline float g_get(const Group<L>&, const GVar<T, L>& S, int src, F f) {
  return f(S.v[src < 0 ? 0 : (src >= L ? L - 1 : src)]);
}
template <int L, class T, class F>
inline int g_get_int(const Group<L>&, const GVar<T, L>& S, int src, F f) {
  return f(S.v[src < 0 ? 0 : (src >= L ? L - 1 : src)]);
}
template <int L, class T, class F>
inline float g_sum(const Group<L>&, const GVar<T, L>& S, F f) {
  // same pairing as the xor butterfly on the device, so that rounding matches
  float t[L];
  for (int l = 0; l < L; ++l) t[l] = f(S.v[l]);
  for (int o = L / 2; o > 0; o >>= 1) {
    float n[L];
    for (int l = 0; l < L; ++l) n[l] = t[l] + t[l ^ o];
    for (int l = 0; l < L; ++l) t[l] = n[l];
  }
  return t[0];
}
template <int L, class T, class F>
inline bool g_any(const Group<L>&, const GVar<T, L>& S, F f) {
  bool p = false;
  for (int l = 0; l < L; ++l) p = p || f(S.v[l]);
  return p;
}
template <int L, class T, class FV, class FI>
inline void g_argmin(const Group<L>&, const GVar<T, L>& S, FV fv, FI fi, float& best, int& tag) {
  best = fv(S.v[0]);
  tag = fi(S.v[0]);
  for (int l = 1; l < L; ++l) {
    const float s = fv(S.v[l]);
    const int i = fi(S.v[l]);
    if (s < best || (s == best && i < tag)) { best = s; tag = i; }
  }
}

#endif

}  // namespace pk
