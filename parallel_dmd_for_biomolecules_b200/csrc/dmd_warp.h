// dmd_warp.h -- the lane-group primitives the engine is written against.  A "warp" of the engine is a group of
// DMD_W lanes that owns one replica:
//   DMD_W == 32  one hardware warp (CTA-per-replica and whole-GPU engines, bulk / energy / retemp kernels)
//   DMD_W == 16  HALF a hardware warp: the warp-per-replica event loop runs TWO replicas per hardware warp.  The loop
//                is bound by the latency of its dependent gathers, not by issue slots or lanes,
//                and the register file limits an SM to 28 hardware warps -- two replicas per warp double the events
//                in flight at the same register cost.  Every collective names the 16 lanes of its own half, so the
//                two halves may diverge (different event types, trip counts, cold paths) and re-converge freely.
//   DMD_W == 1   DMD_HOST_TRACE build (g++, tests/host_trace only): a 1-lane "warp" so the SAME engine source can be
//                single-stepped on the CPU against the oracle.  Test scaffolding; never linked into libdmdb200.so.
// This header (like dmd_engine.h) has no include guard: dmd_cuda.cu includes it once per lane count, each time
// inside its own namespace (DMD_VARIANT_BEGIN / DMD_VARIANT_END), with DMD_W set by the includer.
#include "dmd_math.h"

#ifndef DMD_W
#error "define DMD_W (lanes per replica) before including dmd_warp.h"
#endif
#ifndef DMD_VARIANT_BEGIN
#define DMD_VARIANT_BEGIN
#define DMD_VARIANT_END
#endif

namespace dmd {
DMD_VARIANT_BEGIN

#if defined(DMD_HOST_TRACE)
struct Warp {
  static inline int lane() { return 0; }
  static inline void sync() {}
  static inline unsigned ballot(bool p) { return p ? 1u : 0u; }
  static inline bool any(bool p) { return p; }
  static inline int shfl(int v, int) { return v; }
  static inline double shfl(double v, int) { return v; }
  static inline int shfl_xor(int v, int) { return v; }
  static inline double shfl_xor(double v, int) { return v; }
  static inline long long shfl_xor(long long v, int) { return v; }
};
#else
struct Warp {
  static __device__ __forceinline__ int lane() { return threadIdx.x & (DMD_W - 1); }
  // first lane of the group inside its hardware warp, and the hardware lane mask of the group
  static __device__ __forceinline__ int base() { return DMD_W == 32 ? 0 : (int)(threadIdx.x & 31u & ~(unsigned)(DMD_W - 1)); }
  static __device__ __forceinline__ unsigned mask() {
    return DMD_W == 32 ? 0xffffffffu : (((1u << (DMD_W & 31)) - 1u) << base());
  }
  static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
  // bit k of the result = lane k of the group
  static __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(mask(), p) >> base(); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(mask(), p); }
  static __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(mask(), v, src, DMD_W); }
  static __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(mask(), v, src, DMD_W); }
  static __device__ __forceinline__ int shfl_xor(int v, int m) { return __shfl_xor_sync(mask(), v, m, DMD_W); }
  static __device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(mask(), v, m, DMD_W); }
  static __device__ __forceinline__ long long shfl_xor(long long v, int m) { return __shfl_xor_sync(mask(), v, m, DMD_W); }
  // REDUX over the lanes `rel` of the group (bit k = lane k); every lane of `rel` calls with the same `rel`
  static __device__ __forceinline__ unsigned reduce_min(unsigned rel, unsigned v) { return __reduce_min_sync(rel << base(), v); }
  static __device__ __forceinline__ unsigned reduce_max(unsigned v) { return __reduce_max_sync(mask(), v); }
};
#endif
constexpr unsigned WARP_ALL = DMD_W == 32 ? 0xffffffffu : ((1u << (DMD_W & 31)) - 1u);  // every lane of the group

// (ordered image, key) arg-min over the lanes `rel` of the group: smallest value, ties -> smallest key (key >= 0);
// all lanes of `rel` get the result.  Three REDUX.MIN instructions on the device instead of a shuffle tree.
DMD_DEV void seg_argmin_ord(unsigned& hi, unsigned& lo, int& key, unsigned rel) {
#if DMD_W > 1
  const unsigned mhi = Warp::reduce_min(rel, hi);
  const bool c1 = hi == mhi;
  const unsigned mlo = Warp::reduce_min(rel, c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == mlo;
  const unsigned mkey = Warp::reduce_min(rel, c2 ? (unsigned)key : 0xffffffffu);
  hi = mhi;
  lo = mlo;
  key = (int)mkey;
#endif
}
DMD_DEV void warp_argmin_ord(unsigned long long& v, int& key) {
#if DMD_W > 1
  unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  seg_argmin_ord(hi, lo, key, WARP_ALL);
  v = ((unsigned long long)hi << 32) | lo;
#endif
}
// (value, key) lexicographic arg-min over the lanes `rel` / over the whole group
DMD_DEV void seg_argmin(double& v, int& key, unsigned rel) {
#if DMD_W > 1
  unsigned hi, lo;
  ord_split(v, hi, lo);
  seg_argmin_ord(hi, lo, key, rel);
  v = ord_join(hi, lo);
#endif
}
DMD_DEV void warp_argmin(double& v, int& key) { seg_argmin(v, key, WARP_ALL); }

// minimum value only
DMD_DEV double warp_min(double v) {
#if DMD_W > 1
  unsigned hi, lo;
  ord_split(v, hi, lo);
  const unsigned mhi = Warp::reduce_min(WARP_ALL, hi);
  const unsigned mlo = Warp::reduce_min(WARP_ALL, hi == mhi ? lo : 0xffffffffu);
  v = ord_join(mhi, mlo);
#endif
  return v;
}

DMD_DEV unsigned warp_or(unsigned x) {
#if DMD_W > 1
  x = __reduce_or_sync(Warp::mask(), x);
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
  v = __reduce_add_sync(Warp::mask(), v);
#endif
  return v;
}

DMD_DEV int warp_max_u(int v) {  // v >= 0
#if DMD_W > 1
  v = (int)Warp::reduce_max((unsigned)v);
#endif
  return v;
}

DMD_VARIANT_END
}  // namespace dmd
