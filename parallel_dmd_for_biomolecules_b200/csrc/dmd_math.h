// dmd_math.h -- lane-count independent scalar helpers of the engine: exact fp64 primitives (sqrt, dnint, explicit
// fma), bit access, and the order-preserving integer image of a double that the calendar's min-reductions use.
// Device build: nvcc, sm_100a.  DMD_HOST_TRACE build (g++, tests/host_trace only): the same functions on the CPU.
#pragma once
#include <stdint.h>

#if defined(DMD_HOST_TRACE)
#include <cmath>
#include <cstring>
#define DMD_DEV inline
namespace dmd {
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
namespace dmd {
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

}  // namespace dmd
