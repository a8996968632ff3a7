// warp_prog.h — the "warp program" dialect the solver is written in.
//
// The solver (mpc_warp_solver.cuh) is one warp-synchronous program: 32 lanes own one
// robot's QP end to end.  It is written as a sequence of lane-local blocks
// (LANES_BEGIN ... LANES_END) separated by warp-level exchange points (shared-memory
// hand-offs + WSYNC, or the W* collectives below).
//
//  * Under nvcc (the product) LANES_BEGIN/END are plain braces: every thread executes the
//    block once with its own `lane`, per-lane variables are registers, collectives are
//    __shfl_xor_sync butterflies, WSYNC is __syncwarp().
//  * With -DMPC_EMULATE (g++, tests/warp_emu only — a development and CI aid so that the
//    kernel logic can be exercised without a GPU; it is NOT shipped, not importable from
//    the package and never used by bench.py or the public API) LANES_BEGIN/END become a
//    `for lane in 0..31` loop, per-lane variables become arrays of 32 and collectives
//    become loops over that array.  Because every cross-lane hand-off happens only at
//    the block boundaries, the sequential emulation is exact.
#pragma once

#if defined(__CUDACC__) && !defined(MPC_EMULATE)
// ------------------------------------------------------------------ device build
#define MPC_FN __device__ __forceinline__
#define MPC_LANEFN __device__ __forceinline__
#define LANES_BEGIN {
#define LANES_END }
#define PER_LANE(T, name) T name
#define PER_LANE_ARR(T, name, dims) T name dims
#define _(name) name
#define WSYNC() __syncwarp()

namespace mpcw {
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor_d(v, m));
  return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}
__device__ __forceinline__ int wsum_i(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
__device__ __forceinline__ int wor_i(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// exclusive prefix sum of an int over the warp (lane order)
__device__ __forceinline__ int wexscan_i(int v, int lane) {
  int s = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, s, d);
    if (lane >= d) s += t;
  }
  return s - v;
}
}  // namespace mpcw
#define WMAX(v) mpcw::wmax(v)
#define WSUM(v) mpcw::wsum(v)
#define WSUM_I(v) mpcw::wsum_i(v)
#define WOR_I(v) mpcw::wor_i(v)
// dst (per-lane int) = exclusive scan of src (per-lane int)
#define WEXSCAN_I(dst, src) dst = mpcw::wexscan_i(src, lane)

#else
// ------------------------------------------------------------------ emulation build
#include <cmath>
#define MPC_FN inline
#define MPC_LANEFN inline
#define LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define LANES_END }
#define PER_LANE(T, name) T name[32]
#define PER_LANE_ARR(T, name, dims) T name[32] dims
#define _(name) name[lane]
#define WSYNC() ((void)0)

namespace mpcw {
inline double wmax(const double (&v)[32]) {
  double m = v[0];
  for (int i = 1; i < 32; ++i) m = std::fmax(m, v[i]);
  return m;
}
inline double wsum(const double (&v)[32]) {
  // same butterfly order as the device build so the rounding matches
  double t[32];
  for (int i = 0; i < 32; ++i) t[i] = v[i];
  for (int m = 16; m > 0; m >>= 1) {
    double u[32];
    for (int i = 0; i < 32; ++i) u[i] = t[i] + t[i ^ m];
    for (int i = 0; i < 32; ++i) t[i] = u[i];
  }
  return t[0];
}
inline int wsum_i(const int (&v)[32]) {
  int s = 0;
  for (int i = 0; i < 32; ++i) s += v[i];
  return s;
}
inline int wor_i(const int (&v)[32]) {
  int s = 0;
  for (int i = 0; i < 32; ++i) s |= v[i];
  return s;
}
inline void wexscan_i(int (&dst)[32], const int (&src)[32]) {
  int s = 0;
  for (int i = 0; i < 32; ++i) {
    int v = src[i];
    dst[i] = s;
    s += v;
  }
}
}  // namespace mpcw
#define WMAX(v) mpcw::wmax(v)
#define WSUM(v) mpcw::wsum(v)
#define WSUM_I(v) mpcw::wsum_i(v)
#define WOR_I(v) mpcw::wor_i(v)
#define WEXSCAN_I(dst, src) mpcw::wexscan_i(dst, src)
#endif
