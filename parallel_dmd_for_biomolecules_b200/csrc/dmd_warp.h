// dmd_warp.h -- the 32-lane warp primitives the engine is written against.
//
// Device build (nvcc, sm_100a): real shuffles / ballots, DMD_W == 32.
// DMD_HOST_TRACE build (g++, tests/host_trace only): a 1-lane "warp" (DMD_W == 1) so the SAME engine source
// can be single-stepped on the CPU against the oracle to debug the event-loop logic without a GPU.  That
// build is test scaffolding; it is never linked into libdmdb200.so and is not a fallback.
#pragma once
#include <stdint.h>

#if defined(DMD_HOST_TRACE)
#include <cmath>
#include <cstring>
#define DMD_DEV inline
#define DMD_W 1
namespace dmd {
struct Warp {
  static inline int lane() { return 0; }
  static inline void sync() {}
  static inline unsigned ballot(bool p) { return p ? 1u : 0u; }
  static inline bool any(bool p) { return p; }
  static inline int shfl(int v, int) { return v; }
  static inline double shfl(double v, int) { return v; }
  static inline int shfl_down(int v, int) { return v; }
  static inline double shfl_down(double v, int) { return v; }
  static inline int shfl_xor(int v, int) { return v; }
  static inline double shfl_xor(double v, int) { return v; }
  static inline long long shfl_xor(long long v, int) { return v; }
};
DMD_DEV int dmd_ffs(unsigned m) { return __builtin_ffs((int)m); }
DMD_DEV int dmd_popc(unsigned m) { return __builtin_popcount(m); }
DMD_DEV double dmd_sqrt(double x) { return std::sqrt(x); }
DMD_DEV double dmd_round(double x) { return std::round(x); }
DMD_DEV double dmd_fma(double a, double b, double c) { return std::fma(a, b, c); }
DMD_DEV double dmd_hi_lo(int hi, unsigned lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | lo;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}
DMD_DEV int dmd_hi(double d) {
  uint64_t b;
  std::memcpy(&b, &d, 8);
  return (int)(b >> 32);
}
DMD_DEV unsigned dmd_lo(double d) {
  uint64_t b;
  std::memcpy(&b, &d, 8);
  return (unsigned)b;
}
}  // namespace dmd
#else
#define DMD_DEV __device__ __forceinline__
#define DMD_W 32
namespace dmd {
struct Warp {
  static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  static __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p); }
  static __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static __device__ __forceinline__ int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static __device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static __device__ __forceinline__ long long shfl_xor(long long v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
  }
};
DMD_DEV int dmd_ffs(unsigned m) { return __ffs((int)m); }
DMD_DEV int dmd_popc(unsigned m) { return __popc(m); }
DMD_DEV double dmd_sqrt(double x) { return sqrt(x); }    // IEEE-exact fp64 sqrt on device
DMD_DEV double dmd_round(double x) { return round(x); }  // round half away from zero == Fortran dnint
DMD_DEV double dmd_fma(double a, double b, double c) { return __fma_rn(a, b, c); }  // explicit: -fmad=false stays on
DMD_DEV double dmd_hi_lo(int hi, unsigned lo) { return __hiloint2double(hi, (int)lo); }
DMD_DEV int dmd_hi(double d) { return __double2hiint(d); }
DMD_DEV unsigned dmd_lo(double d) { return (unsigned)__double2loint(d); }
}  // namespace dmd
#endif

namespace dmd {

// fp64 -> uint64 whose unsigned order equals the numeric order (no NaNs on this path)
DMD_DEV void ord_split(double v, unsigned& hi, unsigned& lo) {
  int h = dmd_hi(v);
  unsigned l = dmd_lo(v);
  if (h < 0) {
    hi = ~(unsigned)h;
    lo = ~l;
  } else {
    hi = (unsigned)h | 0x80000000u;
    lo = l;
  }
}
DMD_DEV double ord_join(unsigned hi, unsigned lo) {
  if (hi & 0x80000000u) return dmd_hi_lo((int)(hi & 0x7fffffffu), lo);
  return dmd_hi_lo((int)~hi, ~lo);
}

// the same image as one 64-bit word (unsigned order == numeric order) and back
DMD_DEV unsigned long long ord_bits64(double v) {
  unsigned hi, lo;
  ord_split(v, hi, lo);
  return ((unsigned long long)hi << 32) | lo;
}
DMD_DEV double ord_value64(unsigned long long b) { return ord_join((unsigned)(b >> 32), (unsigned)b); }

// (ordered image, key) arg-min over the warp: smallest value, ties -> smallest key; all lanes get the result
DMD_DEV void warp_argmin_ord(unsigned long long& v, int& key) {
#if DMD_W > 1
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const bool c1 = hi == mhi;
  const unsigned mlo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == mlo;
  const unsigned mkey = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)key : 0xffffffffu);
  v = ((unsigned long long)mhi << 32) | mlo;
  key = (int)mkey;
#endif
}

// (value, key) lexicographic arg-min over the warp: smallest value, ties -> smallest key (key >= 0).
// All lanes get the result.  Three REDUX.MIN instructions on the device instead of a 5-round shuffle tree.
DMD_DEV void warp_argmin(double& v, int& key) {
#if DMD_W > 1
  unsigned hi, lo;
  ord_split(v, hi, lo);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const bool c1 = hi == mhi;
  const unsigned mlo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == mlo;
  const unsigned mkey = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)key : 0xffffffffu);
  v = ord_join(mhi, mlo);
  key = (int)mkey;
#endif
}

// minimum value only
DMD_DEV double warp_min(double v) {
#if DMD_W > 1
  unsigned hi, lo;
  ord_split(v, hi, lo);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  v = ord_join(mhi, mlo);
#endif
  return v;
}

DMD_DEV unsigned warp_or(unsigned x) {
#if DMD_W > 1
  x = __reduce_or_sync(0xffffffffu, x);
#endif
  return x;
}

DMD_DEV double warp_max(double v) {
#if DMD_W > 1
#pragma unroll
  for (int m = DMD_W / 2; m >= 1; m >>= 1) {
    double ov = Warp::shfl_xor(v, m);
    if (ov > v) v = ov;
  }
#endif
  return v;
}

DMD_DEV int warp_sum(int v) {
#if DMD_W > 1
  v = __reduce_add_sync(0xffffffffu, v);
#endif
  return v;
}

}  // namespace dmd
