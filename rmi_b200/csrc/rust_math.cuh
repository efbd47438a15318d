// rust_math.cuh — device-side arithmetic with the exact semantics the reference's Rust
// release build has: saturating float->int casts, masked shifts, wrapping integer ops,
// IEEE-754 round-to-nearest add/mul/div with NO contraction, fma only where the reference
// writes mul_add.  Everything here is __forceinline__ and has no state.
//
// The translation units that include this header are compiled with -fmad=false, and the
// sequential-recurrence code additionally uses the explicit __d*_rn intrinsics so that a
// stray compiler flag cannot fuse a multiply-add the reference does not fuse.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace rmi {

typedef unsigned long long u64;
typedef unsigned int u32;

// Rust `f64 as u64` / `as usize` (saturating, NaN -> 0); reference models/mod.rs:736.
__device__ __forceinline__ u64 f64_to_u64_sat(double v) {
  if (!(v > 0.0)) return 0ull;
  if (v >= 18446744073709551616.0) return ~0ull;
  return (u64)__double2ull_rz(v);
}
// Rust release-mode `<<` / `>>` on u64 with a u8 amount: the amount is masked to 6 bits.
__device__ __forceinline__ u64 shl64(u64 x, unsigned s) { return x << (s & 63u); }
__device__ __forceinline__ u64 shr64(u64 x, unsigned s) { return x >> (s & 63u); }

// f64::max(a, b) in Rust returns the non-NaN operand when one is NaN (maxNum); CUDA's fmax
// has the same rule.
__device__ __forceinline__ double rust_fmax(double a, double b) { return fmax(a, b); }

// ---- key traits: reference models/mod.rs:65-111 (TrainingKey for u64 / u32 / f64) --------
template <class T> struct Key;
template <> struct Key<u64> {
  static constexpr bool is_float = false;
  __device__ __forceinline__ static double as_float(u64 k) { return __ull2double_rn(k); }
  __device__ __forceinline__ static u64 as_int(u64 k) { return k; }
  __device__ __forceinline__ static u64 minus_epsilon(u64 k) { return k - 1ull; }
  __device__ __forceinline__ static u64 plus_epsilon(u64 k) { return k + 1ull; }
  __device__ __forceinline__ static u64 zero_value() { return 0ull; }
  __device__ __forceinline__ static u64 max_value() { return ~0ull; }
};
template <> struct Key<u32> {
  static constexpr bool is_float = false;
  __device__ __forceinline__ static double as_float(u32 k) { return (double)k; }
  __device__ __forceinline__ static u64 as_int(u32 k) { return (u64)k; }
  __device__ __forceinline__ static u32 minus_epsilon(u32 k) { return k - 1u; }
  __device__ __forceinline__ static u32 plus_epsilon(u32 k) { return k + 1u; }
  __device__ __forceinline__ static u32 zero_value() { return 0u; }
  __device__ __forceinline__ static u32 max_value() { return ~0u; }
};
template <> struct Key<double> {
  static constexpr bool is_float = true;
  __device__ __forceinline__ static double as_float(double k) { return k; }
  __device__ __forceinline__ static u64 as_int(double k) { return f64_to_u64_sat(k); }
  __device__ __forceinline__ static double minus_epsilon(double k) { return __dadd_rn(k, -DBL_EPSILON); }
  __device__ __forceinline__ static double plus_epsilon(double k) { return __dadd_rn(k, DBL_EPSILON); }
  __device__ __forceinline__ static double zero_value() { return 0.0; }
  __device__ __forceinline__ static double max_value() { return DBL_MAX; }
};

// ---- correctly rounded a / c for a count divisor c -----------------------------------------
// The reference's Welford step divides by the running count twice per item (linear.rs:27-28).
// With rc = RN(1/c) the three-operation sequence below returns RN(a/c) exactly:
//   q0 = RN(a*rc); rem = a - c*q0 (exact, fma); q = RN(q0 + rem*rc)
// because q0 + rem*rc = (a/c)(1 + eps*delta) with |eps| <= 2^-53 (error of rc) and
// |delta| <= 2^-51 (error of q0), i.e. within 2^-104 of a/c, while a quotient of a double by
// an integer c < 2^32 is never closer than 2^-86 (relative) to a rounding boundary unless it
// is exactly representable (quotients are never exact midpoints).  The sequence is exact only
// while no intermediate underflows/overflows, so magnitudes outside [2^-900, 2^900] (and
// NaN/inf) take the IEEE division instruction sequence instead; a == 0 returns a.
__device__ __forceinline__ double div_by_count(double a, double c, double rc) {
  double aa = fabs(a);
  if (aa > 1.183e-271 && aa < 8.452e270) {   // 2^-900 .. 2^900
    double q0 = __dmul_rn(a, rc);
    double rem = __fma_rn(-c, q0, a);
    return __fma_rn(rem, rc, q0);
  }
  if (a == 0.0) return a;
  return __ddiv_rn(a, c);
}

// x^3 correctly rounded up to a 2^-100 error before the final rounding (double-double
// product); stands in for libm pow(x, 3.0) in cubic_spline.rs:76-93.
__device__ __forceinline__ double cube_dd(double x) {
  double h = __dmul_rn(x, x);
  double l = __fma_rn(x, x, -h);           // x*x = h + l exactly
  double h2 = __dmul_rn(h, x);
  double l2 = __fma_rn(h, x, -h2);         // h*x = h2 + l2 exactly
  l2 = __fma_rn(l, x, l2);                 // + l*x
  return __dadd_rn(h2, l2);
}

}  // namespace rmi
