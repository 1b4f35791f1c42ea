// gbp_math.cuh — per-node FP64 math of the GBP hot path, register resident.
//
// Every routine states the reference function it implements (paths relative to
// crates/magics/src in the reference).  The evaluation order of each sum and
// product follows the order the reference's ndarray expressions produce
// (k-ascending accumulation from 0 for mat*mat, left-to-right for the short
// dot products), with structurally-zero terms dropped: adding an exact 0 or
// multiplying by an exact +-1 does not change an IEEE result, so the bits match
// a literal evaluation as long as the file is compiled with -fmad=false.
#pragma once
#include <cstdint>
#ifdef GBP_HOST_MATH_TEST
// tests/host_math/: the same header compiled by g++ (-ffp-contract=off) so that the device
// arithmetic can be compared with the oracle bit for bit without a GPU
#include "host_math_shim.h"
#else
#include <cuda_runtime.h>
#define GBP_DEV __device__ __forceinline__
#define GBP_NOINLINE_DEV __device__ __noinline__
#endif

namespace gbp {

// An inbox slot holding Message::empty() (message.rs:134-135) is encoded in the
// first double of its record by a NaN with a private payload; arithmetic NaNs
// never carry this payload, so a NaN-poisoned (non-empty) message stays distinct.
constexpr unsigned long long kEmptyBits = 0x7FF8DEADBEEF0001ULL;
GBP_DEV double empty_marker() { return __longlong_as_double((long long)kEmptyBits); }
GBP_DEV bool is_empty_marker(double x) {
  return (unsigned long long)__double_as_longlong(x) == kEmptyBits;
}

// 3x3 minor of a row-major 4x4 with row SR and column SC removed; the
// determinant expression is ndarray-inverse's explicit expansion (third-party
// `Inverse::det`, call sites marginalise_factor_distance.rs:79, variable.rs:278).
template <int SR, int SC>
GBP_DEV double minor3(const double (&m)[16]) {
  constexpr int r0 = SR == 0 ? 1 : 0, r1 = SR <= 1 ? 2 : 1, r2 = SR <= 2 ? 3 : 2;
  constexpr int c0 = SC == 0 ? 1 : 0, c1 = SC <= 1 ? 2 : 1, c2 = SC <= 2 ? 3 : 2;
  const double a = m[r0 * 4 + c0], b = m[r0 * 4 + c1], c = m[r0 * 4 + c2];
  const double d = m[r1 * 4 + c0], e = m[r1 * 4 + c1], f = m[r1 * 4 + c2];
  const double g = m[r2 * 4 + c0], h = m[r2 * 4 + c1], i = m[r2 * 4 + c2];
  return a * e * i + b * f * g + c * d * h - c * e * g - b * d * i - a * f * h;
}

// `Inverse::inv` for 4x4: adjugate / det, None iff det == 0.
//
// The 16 quotients x / det share their divisor.  An IEEE division costs ~15 instructions;
// divide_all computes the correctly rounded reciprocal y = RN(1/det) once and each quotient as
//   q0 = RN(x*y);  r = RN(x - q*det) (exact, one FMA);  q = RN(q + r*y)      (twice)
// which is the correctly rounded x/det whenever q0 is within 1 ulp of it and nothing
// over/underflows (Markstein's theorem; the second pass makes the 1-ulp premise certain).
// Operands outside [2^-500, 2^500] (denormals, inf, NaN) take the plain division, so the
// result is bit-identical to `x / det` for every input.
GBP_DEV bool exp_in_safe_range(double v) {
  const unsigned e = (unsigned(__double2hiint(v)) >> 20) & 0x7ffu;
  return e - 523u < 1001u;  // 2^-500 <= |v| < 2^501
}
// out-of-line: the rare operands share one copy of the IEEE division sequence
GBP_NOINLINE_DEV double plain_div(double x, double det) { return x / det; }

// The same quotients for callers whose numerators are practically never zero (the cofactors of the decoupled
// inverse, gbp_math_axis.cuh): a zero numerator simply takes the plain division with everything else that is out
// of range, so the per-numerator `!= 0` tests and the select after the second correction step go away, and the
// range tests collapse into one unsigned maximum of the exponent fields (biased so that "below the range" wraps
// to "far above it").  Same bits as divide_all for every input; ~70 instructions fewer per k_iterate_axis launch.
template <int N>
GBP_DEV void divide_all_nz(const double (&c)[N], double det, double (&o)[N]) {
#ifdef GBP_LITERAL_DIV
#pragma unroll
  for (int k = 0; k < N; ++k) o[k] = c[k] / det;
#else
  constexpr unsigned kLo = 523u << 20, kSpan = 1001u << 20;  // 2^-500 <= |v| < 2^501
  unsigned worst = (unsigned(__double2hiint(det)) & 0x7ff00000u) - kLo;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const unsigned u = (unsigned(__double2hiint(c[k])) & 0x7ff00000u) - kLo;
    worst = u > worst ? u : worst;
  }
  if (worst < kSpan) {
    const double y = 1.0 / det;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double q0 = c[k] * y;
      double r = fma(-q0, det, c[k]);
      double q = fma(r, y, q0);
      r = fma(-q, det, c[k]);
      o[k] = fma(r, y, q);
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) o[k] = plain_div(c[k], det);
  }
#endif
}

// o[k] = c[k] / det for N numerators, bit-identical to the N divisions.
template <int N>
GBP_DEV void divide_all(const double (&c)[N], double det, double (&o)[N]) {
#ifdef GBP_LITERAL_DIV
#pragma unroll
  for (int k = 0; k < N; ++k) o[k] = c[k] / det;
#else
  // no short-circuit: N independent tests OR-ed together instead of N branches
  bool bad = !exp_in_safe_range(det);
#pragma unroll
  for (int k = 0; k < N; ++k) bad |= (c[k] != 0.0) & !exp_in_safe_range(c[k]);
  if (!bad) {
    const double y = 1.0 / det;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double q0 = c[k] * y;
      double r = fma(-q0, det, c[k]);
      double q = fma(r, y, q0);
#ifndef GBP_DIV_ONE_STEP
      r = fma(-q, det, c[k]);
      q = fma(r, y, q);
#endif
      // a zero numerator keeps its IEEE sign through the first product alone
      o[k] = (c[k] == 0.0) ? q0 : q;
    }
  } else {
    // fully unrolled (N calls of the out-of-line division): a rolled loop would index c/o
    // dynamically and push both arrays into local memory for EVERY path through this function
#pragma unroll
    for (int k = 0; k < N; ++k) o[k] = plain_div(c[k], det);
  }
#endif
}

// `Inverse::inv`, general 4x4: 16 cofactors by explicit expansion, det by Laplace along row 0.
// -DGBP_INV4_GENERAL_NOINLINE: one out-of-line copy instead of one per call site (k_iterate<1,1> is 165 kB of SASS
// and misses the instruction cache 30 % of the time in the dense regime: profiles/r02y).
#ifdef GBP_INV4_GENERAL_NOINLINE
#define GBP_INV4G_DEV GBP_NOINLINE_DEV
#else
#define GBP_INV4G_DEV GBP_DEV
#endif
GBP_INV4G_DEV bool inv4_general(const double (&m)[16], double (&o)[16]) {
  double c[16];
  c[0] = minor3<0, 0>(m);
  c[4] = -minor3<0, 1>(m);
  c[8] = minor3<0, 2>(m);
  c[12] = -minor3<0, 3>(m);
  // Laplace expansion along row 0 (c[4], c[12] carry the cofactor signs; -(a*b) == a*(-b) exactly)
  const double det = m[0] * c[0] - m[1] * (-c[4]) + m[2] * c[8] - m[3] * (-c[12]);
  if (det == 0.0) return false;
  c[1] = -minor3<1, 0>(m);
  c[5] = minor3<1, 1>(m);
  c[9] = -minor3<1, 2>(m);
  c[13] = minor3<1, 3>(m);
  c[2] = minor3<2, 0>(m);
  c[6] = -minor3<2, 1>(m);
  c[10] = minor3<2, 2>(m);
  c[14] = -minor3<2, 3>(m);
  c[3] = -minor3<3, 0>(m);
  c[7] = minor3<3, 1>(m);
  c[11] = -minor3<3, 2>(m);
  c[15] = minor3<3, 3>(m);
  divide_all(c, det, o);
  return true;
}
#ifndef GBP_INV4_DECOUPLED
#define GBP_INV4_DECOUPLED 1
#endif
GBP_DEV bool inv4(const double (&m)[16], double (&o)[16]) {
#if GBP_INV4_DECOUPLED
  // State order is (x, y, vx, vy).  Unless an Obstacle, Tracking or InterRobot message is present the
  // x and y chains never mix: every entry m[r][c] with r + c odd is an exact zero.  Each cofactor of
  // the general expansion below then keeps two of its six triple products (the other four have an
  // exact-zero factor and add +-0), eight cofactors vanish altogether, and det = m00*C00 + m02*C02.
  // Same products, same association, same order: every non-zero of the result has the bits of the
  // general formula (the sign of an exact zero is not tracked).  A non-finite entry reaches det
  // (all eight entries feed it) and sends the matrix down the general path, where 0 * inf matters.
  if ((m[1] == 0.0) & (m[3] == 0.0) & (m[4] == 0.0) & (m[6] == 0.0) & (m[9] == 0.0) & (m[11] == 0.0) &
      (m[12] == 0.0) & (m[14] == 0.0)) {
    double c[8];
    c[0] = (m[5] * m[10]) * m[15] - (m[7] * m[10]) * m[13];  // minor<0,0>
    c[1] = (m[7] * m[8]) * m[13] - (m[5] * m[8]) * m[15];    // minor<0,2>
    const double det = m[0] * c[0] + m[2] * c[1];
    if (isfinite(det)) {
      if (det == 0.0) return false;
      c[2] = (m[0] * m[10]) * m[15] - (m[2] * m[8]) * m[15];  // minor<1,1>
      c[3] = (m[2] * m[8]) * m[13] - (m[0] * m[10]) * m[13];  // minor<1,3>
      c[4] = (m[2] * m[7]) * m[13] - (m[2] * m[5]) * m[15];   // minor<2,0>
      c[5] = (m[0] * m[5]) * m[15] - (m[0] * m[7]) * m[13];   // minor<2,2>
      c[6] = (m[2] * m[7]) * m[8] - (m[0] * m[7]) * m[10];    // minor<3,1>
      c[7] = (m[0] * m[5]) * m[10] - (m[2] * m[5]) * m[8];    // minor<3,3>
      double q[8];
      divide_all(c, det, q);
      // adjugate placement: cofactor <SR,SC> lands at [SC][SR]
      o[0] = q[0];
      o[1] = 0.0;
      o[2] = q[4];
      o[3] = 0.0;
      o[4] = 0.0;
      o[5] = q[2];
      o[6] = 0.0;
      o[7] = q[6];
      o[8] = q[1];
      o[9] = 0.0;
      o[10] = q[5];
      o[11] = 0.0;
      o[12] = 0.0;
      o[13] = q[3];
      o[14] = 0.0;
      o[15] = q[7];
      return true;
    }
  }
#endif
  // x and y are coupled (an Obstacle, Tracking or InterRobot message is in the sum).  (One shared
  // out-of-line copy of this expansion, reached through local memory, measured 2 % slower.)
  return inv4_general(m, o);
}
// Rows 0 and 1 of the inverse only (all an InterRobot Schur complement reads).
GBP_DEV bool inv4_rows01(const double (&m)[16], double (&r0)[4], double (&r1)[4]) {
  double c[8], o[8];
  c[0] = minor3<0, 0>(m);
  c[4] = -minor3<0, 1>(m);
  const double m02 = minor3<0, 2>(m), m03 = minor3<0, 3>(m);
  const double det = m[0] * c[0] - m[1] * (-c[4]) + m[2] * m02 - m[3] * m03;
  if (det == 0.0) return false;
  c[1] = -minor3<1, 0>(m);
  c[5] = minor3<1, 1>(m);
  c[2] = minor3<2, 0>(m);
  c[6] = -minor3<2, 1>(m);
  c[3] = -minor3<3, 0>(m);
  c[7] = minor3<3, 1>(m);
  divide_all(c, det, o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r0[k] = o[k];
    r1[k] = o[4 + k];
  }
  return true;
}

// VariableNode::update_belief_and_create_factor_responses, the part after the
// inbox sum (variable.rs:273-297): mean / covariance from (eta, lam).
// Returns true when the covariance was (re)computed; mu is only overwritten
// when the new covariance is finite.
GBP_DEV bool belief_moments(const double (&eta)[4], const double (&lam)[16], double (&mu)[4],
                            double (&cov)[16], bool &valid) {
  // `any(|x| *x - 1e-6 > 0.0)` (variable.rs:276): x - 1e-6 > 0 exactly when x > 1e-6 (the
  // difference of two doubles never rounds across zero), for NaN and infinities too
  bool nz = false;
#pragma unroll
  for (int k = 0; k < 16; ++k) nz |= lam[k] > 1e-6;
  if (!nz) return false;
  if (!inv4(lam, cov)) return false;
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 16; ++k) fin &= isfinite(cov[k]);
  valid = fin;
  if (fin) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      mu[r] = 0.0 + cov[r * 4 + 0] * eta[0] + cov[r * 4 + 1] * eta[1] + cov[r * 4 + 2] * eta[2] +
              cov[r * 4 + 3] * eta[3];
  }
  return true;
}

// Scalar 4x4 pattern M of the DynamicFactor potential Lambda_p = J^T Qi^-1 J over
// (pos_i, vel_i, pos_i+1, vel_i+1); the 8x8 is M (x) I2 (dynamic.rs:22-52 and
// factor/mod.rs:391-394).  qs = 1/sigma^2.
struct DynM {
  double m[4][4];
};
// The three distinct entries of Qi^-1 / sigma^2 (dynamic.rs:22-52): they depend on the factor's
// delta_t and sigma only, so the engine evaluates them once per factor (k_init_vars) and keeps them
// next to delta_t in Store::dyn_c — six divisions per thread and launch less, same bits.
GBP_DEV void dyn_q(double dt, double qs, double &q11, double &q12, double &q22) {
  const double p3 = 1.0 / ((dt * dt) * dt);
  const double p2 = 1.0 / (dt * dt);
  q11 = (12.0 * p3) * qs;
  q12 = (-6.0 * p2) * qs;
  q22 = (4.0 / dt) * qs;
}
GBP_DEV DynM dyn_potential_q(double dt, double q11, double q12, double q22) {
  const double a[4] = {q11, dt * q11 + q12, -q11, -q12};
  const double b[4] = {q12, dt * q12 + q22, -q12, -q22};
  DynM r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.m[k][0] = a[k];
    r.m[k][1] = a[k] * dt + b[k];
    r.m[k][2] = -a[k];
    r.m[k][3] = -b[k];
  }
  return r;
}
GBP_DEV DynM dyn_potential(double dt, double qs) {
  double q11, q12, q22;
  dyn_q(dt, qs, q11, q12, q22);
  return dyn_potential_q(dt, q11, q12, q22);
}

// FactorNode::update for a DynamicFactor, one outgoing message
// (factor/mod.rs:412-450 + marginalise_factor_distance.rs:55-127).
// KEEP = 0: message to the first variable (slot 0), marginalising slot 1;
// KEEP = 1: message to the second.  `oe`/`ol` is the OTHER variable's message
// (eta, Lambda) if `other_nonempty`.  Returns false for Message::empty().
// `o` = the OTHER variable's message as one record (eta 0..3, Lambda 4..19).
template <int KEEP>
GBP_DEV bool dyn_message(const DynM &M, bool other_nonempty, const double (&o)[20], double (&eta)[4],
                         double (&lam)[16]) {
  constexpr int A = KEEP * 2, B = (1 - KEEP) * 2;  // scalar-block offsets in M
  const double *oe = o, *ol = o + 4;
  // Lambda_bb = potential block + other message; index (kk, dim) -> kk*2 + dim
  double bb[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool same = (r & 1) == (c & 1);
      if (same) {
        const double p = M.m[B + (r >> 1)][B + (c >> 1)];
        bb[r * 4 + c] = other_nonempty ? p + ol[r * 4 + c] : p;
      } else {
        bb[r * 4 + c] = other_nonempty ? 0.0 + ol[r * 4 + c] : 0.0;
      }
    }
  double bi[16];
  if (!inv4(bb, bi)) return false;
  double eb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) eb[k] = other_nonempty ? 0.0 + oe[k] : 0.0;
  // T = Lambda_ab * Binv ; Lambda_ab[(rr,d)][(kk,d)] = M[A+rr][B+kk]
  double T[16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = r & 1, rr = r >> 1;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      T[r * 4 + c] = M.m[A + rr][B + 0] * bi[(0 + d) * 4 + c] + M.m[A + rr][B + 1] * bi[(2 + d) * 4 + c];
  }
  bool inf = false;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double te =
        0.0 + T[r * 4 + 0] * eb[0] + T[r * 4 + 1] * eb[1] + T[r * 4 + 2] * eb[2] + T[r * 4 + 3] * eb[3];
    eta[r] = 0.0 - te;  // eta_p == 0 exactly for the dynamic factor
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int dc = c & 1, cc = c >> 1;
      // (T * Lambda_ba)[r][c], Lambda_ba[(kk,dc)][(cc,dc)] = M[B+kk][A+cc]
      const double tl = T[r * 4 + dc] * M.m[B + 0][A + cc] + T[r * 4 + 2 + dc] * M.m[B + 1][A + cc];
      const double aa = ((r & 1) == dc) ? M.m[A + (r >> 1)][A + cc] : 0.0;
      const double v = aa - tl;
      lam[r * 4 + c] = v;
      inf |= isinf(v);
    }
  }
  return !inf;
}

// One-row measurement factors (Obstacle / Tracking): the message is the whole
// potential (marginalise_factor_distance.rs:63-72), Lambda[k][l] = (J_k*lm)*J_l,
// eta[k] = (J_k*lm)*v0 with v0 = J.x + (z - h) (factor/mod.rs:391-401).
GBP_DEV void unary_add(const double (&J)[4], double v0, double lm, double (&eta)[4],
                       double (&lam)[16]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double g = J[k] * lm;
    eta[k] = eta[k] + g * v0;
#pragma unroll
    for (int l = 0; l < 4; ++l) lam[k * 4 + l] = lam[k * 4 + l] + g * J[l];
  }
}

// FactorNode::update for an InterRobotFactor, only the message to the variable of
// the robot that is NOT the sender of (etaA, lamA, muA)
// (interrobot.rs:91-226, factor/mod.rs:334-454, marginalise_factor_distance.rs).
//   a_first    robot A's variable is slot 0 (A's id < B's id, id.rs:83-118)
//   muA/muB    position part of the linearisation point of each slot (zeros if empty)
//   a_nonempty whether A's variable message exists (else only the potential)
//   dsafe      safety distance, tiny the factor's tiny_offset, lm = 1/sigma^2
// Output: eta[0..1], lam 2x2 (rows/cols 0..1 of the 4x4; the rest is exactly 0).
// InterRobotFactor::skip (interrobot.rs:213-226) on the linearisation point: the squared distance
// of the two position means (no tiny_offset) against the safety distance.  Exactly the first
// test of interrobot_message, exposed so that the caller can decide BEFORE fetching A's 20-double
// (eta, Lambda) record: in a swarm most robots within comms range are outside safety range.
GBP_DEV bool interrobot_skip(bool a_first, const double (&muA)[2], const double (&muB)[2], double dsafe) {
  const double x0 = a_first ? muA[0] : muB[0], x1 = a_first ? muA[1] : muB[1];
  const double x4 = a_first ? muB[0] : muA[0], x5 = a_first ? muB[1] : muA[1];
  const double e0 = x0 - x4, e1 = x1 - x5;
  return (0.0 + e0 * e0) + e1 * e1 >= dsafe * dsafe;
}

// `recA` = A's published record (eta 0..3, Lambda 4..19); only read when a_nonempty.
GBP_DEV bool interrobot_message(bool a_first, const double (&muA)[2], const double (&muB)[2],
                                bool a_nonempty, const double (&recA)[20], double dsafe, double tiny,
                                double lm, double (&eta)[2], double (&lam)[4]) {
  const double *etaA = recA, *lamA = recA + 4;
  const double x0 = a_first ? muA[0] : muB[0], x1 = a_first ? muA[1] : muB[1];
  const double x4 = a_first ? muB[0] : muA[0], x5 = a_first ? muB[1] : muA[1];
  const double e0 = x0 - x4, e1 = x1 - x5;
  if ((0.0 + e0 * e0) + e1 * e1 >= dsafe * dsafe) return false;  // InterRobotFactor::skip
  const double d0 = e0 + tiny, d1 = e1 + tiny;
  const double rad = sqrt((0.0 + d0 * d0) + d1 * d1);
  double j0[2] = {0.0, 0.0}, j4[2] = {0.0, 0.0}, h = 0.0;
  if (rad <= dsafe) {
    const double ca = -1.0 / dsafe / rad, cb = 1.0 / dsafe / rad;
    j0[0] = ca * d0;
    j0[1] = ca * d1;
    j4[0] = cb * d0;
    j4[1] = cb * d1;
    h = 1.0 * (1.0 - rad / dsafe);
  }
  // J.x via ndarray's unrolled_dot pairing (p0+p4)+(p1+p5), then + (z - h)
  const double v0 = ((0.0 + (j0[0] * x0 + j4[0] * x4)) + (j0[1] * x1 + j4[1] * x5)) + (0.0 - h);
  const double jA[2] = {a_first ? j0[0] : j4[0], a_first ? j0[1] : j4[1]};
  const double jB[2] = {a_first ? j4[0] : j0[0], a_first ? j4[1] : j0[1]};
  const double gA[2] = {jA[0] * lm, jA[1] * lm}, gB[2] = {jB[0] * lm, jB[1] * lm};
  double bb[16], eb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double p = (r < 2 && c < 2) ? gA[r] * jA[c] : 0.0;
      bb[r * 4 + c] = a_nonempty ? p + lamA[r * 4 + c] : p;
    }
    const double pe = (r < 2) ? gA[r] * v0 : 0.0;
    eb[r] = a_nonempty ? pe + etaA[r] : pe;
  }
  double i0[4], i1[4];
  if (!inv4_rows01(bb, i0, i1)) return false;
  bool inf = false;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double ab0 = gB[r] * jA[0], ab1 = gB[r] * jA[1];  // Lambda_ab row r
    double T[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) T[c] = ab0 * i0[c] + ab1 * i1[c];
    const double te = 0.0 + T[0] * eb[0] + T[1] * eb[1] + T[2] * eb[2] + T[3] * eb[3];
    eta[r] = gB[r] * v0 - te;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double tl = T[0] * (gA[0] * jB[c]) + T[1] * (gA[1] * jB[c]);
      const double v = gB[r] * jB[c] - tl;
      lam[r * 2 + c] = v;
      inf = inf || isinf(v);
    }
  }
  return !inf;
}

// Rust `as u32` (saturating; obstacle.rs:153-155).
// cvt.rzi.u32.f64 saturates out-of-range values like Rust but maps NaN to 2^31; Rust maps NaN to 0.
GBP_DEV uint32_t sat_u32(double v) { return isnan(v) ? 0u : __double2uint_rz(v); }

}  // namespace gbp
