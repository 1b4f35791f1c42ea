// host_math_shim.h — lets g++ compile magics_b200/csrc/gbp_math.cuh (test infrastructure only).
// The device intrinsics the header uses are restated with their documented semantics; compile with
// -ffp-contract=off so that, like nvcc -fmad=false, no product-sum is contracted behind our back.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define GBP_DEV inline
#define GBP_NOINLINE_DEV inline

using std::fma;
using std::isfinite;
using std::isinf;
using std::isnan;
using std::sqrt;

inline double __longlong_as_double(long long v) {
  double d;
  std::memcpy(&d, &v, 8);
  return d;
}
inline long long __double_as_longlong(double d) {
  long long v;
  std::memcpy(&v, &d, 8);
  return v;
}
inline int __double2hiint(double d) { return int(uint64_t(__double_as_longlong(d)) >> 32); }
// cvt.rzi.u32.f64: truncation, saturating; NaN -> 0x80000000 on the device (the caller handles NaN first)
inline unsigned __double2uint_rz(double v) {
  if (v != v) return 0x80000000u;
  if (v <= 0.0) return 0u;
  if (v >= 4294967295.0) return 0xffffffffu;
  return unsigned(v);
}
