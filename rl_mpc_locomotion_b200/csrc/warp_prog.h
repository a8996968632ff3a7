// warp_prog.h — the "lane program" dialect the solver is written in.
//
// The solver (mpc_warp_solver.cuh) is one lock-step program executed by MPC_NL lanes that own
// one robot's QP end to end (MPC_NL = 64: two warps = one CTA per robot; 32: one warp).  It
// is written as a sequence of lane-local blocks (LANES_BEGIN ... LANES_END) separated by
// exchange points (shared-memory hand-offs + WSYNC, or the W* collectives below).
//
//  * Under nvcc (the product) LANES_BEGIN/END are plain braces: every thread executes the
//    block once with its own `lane`, per-lane variables are registers, collectives are
//    __shfl_xor_sync butterflies (+ one shared-memory hop between the two warps), WSYNC is
//    __syncwarp() / __syncthreads().
//  * With -DMPC_EMULATE (g++, tests/warp_emu only — a development and CI aid so that the
//    kernel logic can be exercised without a GPU; it is NOT shipped, not importable from
//    the package and never used by bench.py or the public API) LANES_BEGIN/END become a
//    `for lane in 0..MPC_NL-1` loop, per-lane variables become arrays and collectives become
//    loops over that array.  Because every cross-lane hand-off happens only at the block
//    boundaries, the sequential emulation is exact.
#pragma once

#ifndef MPC_NL
#define MPC_NL 64
#endif
static_assert(MPC_NL == 32 || MPC_NL == 64, "MPC_NL must be 32 or 64");

#if defined(__CUDACC__) && !defined(MPC_EMULATE)
// ------------------------------------------------------------------ device build
#define MPC_FN __device__ __forceinline__
#define MPC_LANEFN __device__ __forceinline__
// a lane program kept out of line (ONE copy in the kernel: instruction-cache footprint); only for
// rarely-called phases that live entirely in shared memory - a call spills the caller's registers
#define MPC_NOINLINE_FN __device__ __noinline__
#define MPC_LANE_ARG , const int lane
#define MPC_LANE_PASS , lane
#define LANES_BEGIN {
#define LANES_END }
#define PER_LANE(T, name) T name
#define PER_LANE_ARR(T, name, dims) T name dims
#define _(name) name
#if MPC_NL == 32
#define WSYNC() __syncwarp()
#else
#define WSYNC() __syncthreads()
#endif

namespace mpcw {
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
// one exchange buffer per CTA (one robot per CTA) for the cross-warp hop: 8 scalar slots + 2x16
// for the vector reduction
__device__ __forceinline__ double* xbuf() {
  __shared__ double buf[8 + 32];
  return buf;
}
// K maxima at once: K butterflies, ONE shared-memory hop (2 barriers instead of 2K)
template <int K>
__device__ __forceinline__ void wmaxv(const double (&v)[K], double (&out)[K]) {
  static_assert(K <= 16, "exchange buffer holds 16 values per warp");
#pragma unroll
  for (int q = 0; q < K; ++q) {
    double t = v[q];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {   // compare-select instead of fmax (7 instructions): the operands are
      const double u = shfl_xor_d(t, m);  // running maxima of |.| that started from 0, never NaN
      t = u > t ? u : t;
    }
    out[q] = t;
  }
#if MPC_NL == 64
  double* b = xbuf() + 8;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int q = 0; q < K; ++q) b[(threadIdx.x >> 5) * 16 + q] = out[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < K; ++q) out[q] = b[16 + q] > b[q] ? b[16 + q] : b[q];
  __syncthreads();
#endif
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const double u = shfl_xor_d(v, m);
    v = u > v ? u : v;
  }
#if MPC_NL == 64
  double* b = xbuf();
  if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
  __syncthreads();
  v = b[1] > b[0] ? b[1] : b[0];
  __syncthreads();
#endif
  return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
#if MPC_NL == 64
  double* b = xbuf();
  if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
  __syncthreads();
  v = b[0] + b[1];
  __syncthreads();
#endif
  return v;
}
__device__ __forceinline__ int wsum_i(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
#if MPC_NL == 64
  int* b = reinterpret_cast<int*>(xbuf() + 4);
  if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
  __syncthreads();
  v = b[0] + b[1];
  __syncthreads();
#endif
  return v;
}
// exclusive prefix sum of an int over the lanes (lane order)
__device__ __forceinline__ int wexscan_i(int v, int lane) {
  int s = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, s, d);
    if ((lane & 31) >= d) s += t;
  }
#if MPC_NL == 64
  int* b = reinterpret_cast<int*>(xbuf() + 6);
  if (lane == 31) b[0] = s;  // total of warp 0
  __syncthreads();
  if (lane >= 32) s += b[0];
  __syncthreads();
#endif
  return s - v;
}
}  // namespace mpcw
#define WMAX(v) mpcw::wmax(v)
#define WSUM(v) mpcw::wsum(v)
#define WSUM_I(v) mpcw::wsum_i(v)
// dst[K] (uniform) = element-wise max over the lanes of the per-lane array src[K]
#define WMAXV(dst, src) mpcw::wmaxv(src, dst)
// dst (per-lane int) = exclusive scan of src (per-lane int)
#define WEXSCAN_I(dst, src) dst = mpcw::wexscan_i(src, lane)

#else
// ------------------------------------------------------------------ emulation build
#include <cmath>
struct double2 {
  double x, y;
};
#define MPC_FN inline
#define MPC_LANEFN inline
#define MPC_NOINLINE_FN inline
#define MPC_LANE_ARG
#define MPC_LANE_PASS
#define LANES_BEGIN for (int lane = 0; lane < MPC_NL; ++lane) {
#define LANES_END }
#define PER_LANE(T, name) T name[MPC_NL]
#define PER_LANE_ARR(T, name, dims) T name[MPC_NL] dims
#define _(name) name[lane]
#define WSYNC() ((void)0)

namespace mpcw {
inline double wmax(const double (&v)[MPC_NL]) {
  double m = v[0];
  for (int i = 1; i < MPC_NL; ++i) m = std::fmax(m, v[i]);
  return m;
}
template <int K>
inline void wmaxv(const double (&v)[MPC_NL][K], double (&out)[K]) {
  for (int q = 0; q < K; ++q) {
    double m = v[0][q];
    for (int i = 1; i < MPC_NL; ++i) m = std::fmax(m, v[i][q]);
    out[q] = m;
  }
}
inline double wsum(const double (&v)[MPC_NL]) {
  // same association order as the device build: a 32-lane butterfly per warp, then warp0+warp1
  double tot[MPC_NL / 32];
  for (int w = 0; w < MPC_NL / 32; ++w) {
    double t[32];
    for (int i = 0; i < 32; ++i) t[i] = v[32 * w + i];
    for (int m = 16; m > 0; m >>= 1) {
      double u[32];
      for (int i = 0; i < 32; ++i) u[i] = t[i] + t[i ^ m];
      for (int i = 0; i < 32; ++i) t[i] = u[i];
    }
    tot[w] = t[0];
  }
  return MPC_NL == 64 ? tot[0] + tot[MPC_NL / 32 - 1] : tot[0];
}
inline int wsum_i(const int (&v)[MPC_NL]) {
  int s = 0;
  for (int i = 0; i < MPC_NL; ++i) s += v[i];
  return s;
}
inline void wexscan_i(int (&dst)[MPC_NL], const int (&src)[MPC_NL]) {
  int s = 0;
  for (int i = 0; i < MPC_NL; ++i) {
    int v = src[i];
    dst[i] = s;
    s += v;
  }
}
}  // namespace mpcw
#define WMAX(v) mpcw::wmax(v)
#define WSUM(v) mpcw::wsum(v)
#define WSUM_I(v) mpcw::wsum_i(v)
#define WMAXV(dst, src) mpcw::wmaxv(src, dst)
#define WEXSCAN_I(dst, src) mpcw::wexscan_i(dst, src)
#endif
