// gbp_math_axis.cuh — the same per-node FP64 math as gbp_math.cuh for the DECOUPLED regime,
// evaluated by two lanes per variable, one per axis.
//
// State order is (x, y, vx, vy).  While a variable holds no Obstacle, Tracking or InterRobot
// message every 4x4 of the hot path (precisions, Dynamic-factor messages, covariances) has exact
// zeros wherever row + column is odd: the x chain (rows/cols 0, 2) and the y chain (1, 3) never
// mix.  A lane of axis a in {0, 1} then owns
//     eta  -> e[2] = (eta[a], eta[a + 2])
//     4x4  -> P[4] = (m[a][a], m[a][a + 2], m[a + 2][a], m[a + 2][a + 2])   ("its block")
// and every sum, difference and product of gbp_math.cuh splits into the two blocks term by
// term: the dropped terms are products with an exact zero, i.e. +-0 added to a running sum that
// started from +0.0 (never -0), which leaves the sum's bits unchanged as long as every eta is
// finite — callers check that and fall back to the general kernel otherwise.  The one place the
// two axes meet is the determinant of the 4x4 cofactor inverse (the reference's ndarray-inverse
// `.inv()`, marginalise_factor_distance.rs:79, variable.rs:278): each cofactor of the decoupled
// fast path of gbp::inv4 multiplies entries of BOTH blocks, so a lane needs its partner's block Q
// and evaluates exactly the cofactors of inv4 that land in its own block of the result.
// Same products, same association, same order => the bits of gbp_math.cuh (tests/
// test_device_math_host.py fuzzes this header, compiled for the host, against it).
#pragma once
#include "gbp_math.cuh"

namespace gbp {

// Own block of inv4(m) for the decoupled m made of block P (this lane's axis) and Q (the other
// axis).  False: singular (det == 0) or a non-finite determinant (the general path of inv4) —
// the caller sends the robot to the general kernel.
// X = the x chain's block (m0, m2, m8, m10), Y = the y chain's (m5, m7, m13, m15), whichever lane holds which.
GBP_DEV bool inv_axis_xy(int a, const double (&X)[4], const double (&Y)[4], double (&O)[4]) {
  const bool y = a != 0;
  // entries by their row-major index in the 4x4
  const double m0 = X[0], m2 = X[1], m8 = X[2], m10 = X[3];
  const double m5 = Y[0], m7 = Y[1], m13 = Y[2], m15 = Y[3];
  const double c0 = (m5 * m10) * m15 - (m7 * m10) * m13;  // minor<0,0>
  const double c1 = (m7 * m8) * m13 - (m5 * m8) * m15;    // minor<0,2>
  const double det = m0 * c0 + m2 * c1;
  if (!isfinite(det) || det == 0.0) return false;
  double c[4];
  if (!y) {  // o[0], o[2], o[8], o[10]
    c[0] = c0;
    c[1] = (m2 * m7) * m13 - (m2 * m5) * m15;  // minor<2,0>
    c[2] = c1;
    c[3] = (m0 * m5) * m15 - (m0 * m7) * m13;  // minor<2,2>
  } else {   // o[5], o[7], o[13], o[15]
    c[0] = (m0 * m10) * m15 - (m2 * m8) * m15;  // minor<1,1>
    c[1] = (m2 * m7) * m8 - (m0 * m7) * m10;    // minor<3,1>
    c[2] = (m2 * m8) * m13 - (m0 * m10) * m13;  // minor<1,3>
    c[3] = (m0 * m5) * m10 - (m2 * m5) * m8;    // minor<3,3>
  }
  divide_all_nz(c, det, O);
  return true;
}
GBP_DEV bool inv_axis(int a, const double (&P)[4], const double (&Q)[4], double (&O)[4]) {
  const bool y = a != 0;
  const double X[4] = {y ? Q[0] : P[0], y ? Q[1] : P[1], y ? Q[2] : P[2], y ? Q[3] : P[3]};
  const double Y[4] = {y ? P[0] : Q[0], y ? P[1] : Q[1], y ? P[2] : Q[2], y ? P[3] : Q[3]};
  return inv_axis_xy(a, X, Y, O);
}

// belief_moments (variable.rs:273-297) for one axis: false unless the update is taken AND valid
// (the other outcomes — precision below 1e-6 everywhere, singular, non-finite covariance — keep an
// older mean / covariance and belong to the general kernel).  mu = own two components.
GBP_DEV bool belief_axis(int a, const double (&e)[2], const double (&P)[4], const double (&Q)[4], double (&mu)[2]) {
  // `precision_matrix.iter().any(|x| *x > 1e-6)` (variable.rs:276): the first diagonal entry settles it for every
  // variable that has a prior or a Dynamic message; the other seven are looked at only if it does not (nvcc turns
  // the eight-way OR into emulated 64-bit max operations, ~50 instructions)
  bool nz = P[0] > 1e-6;
  if (!nz) {
#pragma unroll
    for (int k = 1; k < 4; ++k) nz |= P[k] > 1e-6;
#pragma unroll
    for (int k = 0; k < 4; ++k) nz |= Q[k] > 1e-6;
  }
  double O[4];
  if (!nz || !inv_axis(a, P, Q, O)) return false;
  bool fin = isfinite(e[0]) & isfinite(e[1]);
#pragma unroll
  for (int k = 0; k < 4; ++k) fin &= isfinite(O[k]);
  if (!fin) return false;
  mu[0] = (0.0 + O[0] * e[0]) + O[1] * e[1];
  mu[1] = (0.0 + O[2] * e[0]) + O[3] * e[1];
  return true;
}

// dyn_message<KEEP> (factor/mod.rs:412-450 + marginalise_factor_distance.rs:55-127) for one axis.
// oe / oP: this axis' part of the OTHER variable's message, oQ: the other axis' block of it.
// False: Message::empty() or a non-finite input — general kernel.
// oX / oY: the x chain's and the y chain's block of the OTHER variable's message (the kernel reads them from the
// even and the odd lane's shared-memory slot: no selects), oe: this axis' part of its vector.
template <int KEEP>
GBP_DEV bool dyn_message_axis_xy(int a, const DynM &M, bool other_nonempty, const double (&oe)[2],
                                 const double (&oX)[4], const double (&oY)[4], double (&eta)[2], double (&lam)[4]) {
  constexpr int A = KEEP * 2, B = (1 - KEEP) * 2;
  double bX[4], bY[4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double p = M.m[B + r][B + c];
      bX[r * 2 + c] = other_nonempty ? p + oX[r * 2 + c] : p;
      bY[r * 2 + c] = other_nonempty ? p + oY[r * 2 + c] : p;
    }
  double I[4];
  if (!inv_axis_xy(a, bX, bY, I)) return false;
  const double eb[2] = {other_nonempty ? 0.0 + oe[0] : 0.0, other_nonempty ? 0.0 + oe[1] : 0.0};
  if (!(isfinite(eb[0]) & isfinite(eb[1]))) return false;
  bool inf = false;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double t0 = M.m[A + r][B + 0] * I[0] + M.m[A + r][B + 1] * I[2];
    const double t1 = M.m[A + r][B + 0] * I[1] + M.m[A + r][B + 1] * I[3];
    const double te = (0.0 + t0 * eb[0]) + t1 * eb[1];
    eta[r] = 0.0 - te;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double tl = t0 * M.m[B + 0][A + c] + t1 * M.m[B + 1][A + c];
      const double v = M.m[A + r][A + c] - tl;
      lam[r * 2 + c] = v;
      inf |= isinf(v);
    }
  }
  return !inf;
}
// oP: this axis' block of the other variable's message, oQ: the other axis' block of it.
template <int KEEP>
GBP_DEV bool dyn_message_axis(int a, const DynM &M, bool other_nonempty, const double (&oe)[2], const double (&oP)[4],
                              const double (&oQ)[4], double (&eta)[2], double (&lam)[4]) {
  return a ? dyn_message_axis_xy<KEEP>(a, M, other_nonempty, oe, oQ, oP, eta, lam)
           : dyn_message_axis_xy<KEEP>(a, M, other_nonempty, oe, oP, oQ, eta, lam);
}

// ---- both Dynamic messages of a variable at once ------------------------------------------------------------
// The two 4x4 cofactor inverses (and the two reciprocals inside them) are independent dependency chains; written
// as one straight-line block they overlap in the FP64 pipe instead of running one after the other (the kernel
// spends a quarter of its issue interval on fixed-latency dependencies).  Same operations per message as
// inv_axis / divide_all / dyn_message_axis, so the same bits.
GBP_DEV void cof_axis_pair(int a, const double (&P1)[4], const double (&Q1)[4], const double (&P2)[4],
                           const double (&Q2)[4], double (&c1)[4], double &det1, double (&c2)[4], double &det2) {
  const bool y = a != 0;
#define GBP_ENTRIES(P, Q, m)                                                                          \
  const double m##0 = y ? Q[0] : P[0], m##2 = y ? Q[1] : P[1], m##8 = y ? Q[2] : P[2], m##10 = y ? Q[3] : P[3]; \
  const double m##5 = y ? P[0] : Q[0], m##7 = y ? P[1] : Q[1], m##13 = y ? P[2] : Q[2], m##15 = y ? P[3] : Q[3];
  GBP_ENTRIES(P1, Q1, u)
  GBP_ENTRIES(P2, Q2, v)
#undef GBP_ENTRIES
  const double u_c0 = (u5 * u10) * u15 - (u7 * u10) * u13, v_c0 = (v5 * v10) * v15 - (v7 * v10) * v13;
  const double u_c1 = (u7 * u8) * u13 - (u5 * u8) * u15, v_c1 = (v7 * v8) * v13 - (v5 * v8) * v15;
  det1 = u0 * u_c0 + u2 * u_c1;
  det2 = v0 * v_c0 + v2 * v_c1;
  if (!y) {
    c1[0] = u_c0;
    c2[0] = v_c0;
    c1[1] = (u2 * u7) * u13 - (u2 * u5) * u15;
    c2[1] = (v2 * v7) * v13 - (v2 * v5) * v15;
    c1[2] = u_c1;
    c2[2] = v_c1;
    c1[3] = (u0 * u5) * u15 - (u0 * u7) * u13;
    c2[3] = (v0 * v5) * v15 - (v0 * v7) * v13;
  } else {
    c1[0] = (u0 * u10) * u15 - (u2 * u8) * u15;
    c2[0] = (v0 * v10) * v15 - (v2 * v8) * v15;
    c1[1] = (u2 * u7) * u8 - (u0 * u7) * u10;
    c2[1] = (v2 * v7) * v8 - (v0 * v7) * v10;
    c1[2] = (u2 * u8) * u13 - (u0 * u10) * u13;
    c2[2] = (v2 * v8) * v13 - (v0 * v10) * v13;
    c1[3] = (u0 * u5) * u10 - (u2 * u5) * u8;
    c2[3] = (v0 * v5) * v10 - (v2 * v5) * v8;
  }
}
// o1 = c1 / det1, o2 = c2 / det2, bit-identical to the divisions (divide_all's scheme, both reciprocals in flight)
GBP_DEV void divide_pair(const double (&c1)[4], double det1, const double (&c2)[4], double det2, double (&o1)[4],
                         double (&o2)[4]) {
  bool bad = !exp_in_safe_range(det1) | !exp_in_safe_range(det2);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    bad |= ((c1[k] != 0.0) & !exp_in_safe_range(c1[k])) | ((c2[k] != 0.0) & !exp_in_safe_range(c2[k]));
  if (!bad) {
    const double y1 = 1.0 / det1, y2 = 1.0 / det2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double p0 = c1[k] * y1, q0 = c2[k] * y2;
      double r1 = fma(-p0, det1, c1[k]), r2 = fma(-q0, det2, c2[k]);
      double p = fma(r1, y1, p0), q = fma(r2, y2, q0);
      r1 = fma(-p, det1, c1[k]);
      r2 = fma(-q, det2, c2[k]);
      p = fma(r1, y1, p);
      q = fma(r2, y2, q);
      o1[k] = (c1[k] == 0.0) ? p0 : p;
      o2[k] = (c2[k] == 0.0) ? q0 : q;
    }
  } else {
    divide_all(c1, det1, o1);
    divide_all(c2, det2, o2);
  }
}
// dyn_message_axis<1> (message to the second variable of Dynamic factor i-1: "L") and dyn_message_axis<0> (to the first
// variable of factor i: "R") together.  ok1 / ok2 = what the single functions return.
GBP_DEV void dyn_message_axis_pair(int a, const DynM &M1, bool ne1, const double (&oe1)[2], const double (&oP1)[4],
                                   const double (&oQ1)[4], const DynM &M2, bool ne2, const double (&oe2)[2],
                                   const double (&oP2)[4], const double (&oQ2)[4], double (&eta1)[2],
                                   double (&lam1)[4], double (&eta2)[2], double (&lam2)[4], bool &ok1, bool &ok2) {
  // message 1: KEEP = 1 -> A = 2, B = 0; message 2: KEEP = 0 -> A = 0, B = 2
  double bP1[4], bQ1[4], bP2[4], bQ2[4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double p1 = M1.m[0 + r][0 + c], p2 = M2.m[2 + r][2 + c];
      bP1[r * 2 + c] = ne1 ? p1 + oP1[r * 2 + c] : p1;
      bQ1[r * 2 + c] = ne1 ? p1 + oQ1[r * 2 + c] : p1;
      bP2[r * 2 + c] = ne2 ? p2 + oP2[r * 2 + c] : p2;
      bQ2[r * 2 + c] = ne2 ? p2 + oQ2[r * 2 + c] : p2;
    }
  double c1[4], c2[4], det1, det2, I1[4], I2[4];
  cof_axis_pair(a, bP1, bQ1, bP2, bQ2, c1, det1, c2, det2);
  ok1 = isfinite(det1) && det1 != 0.0;
  ok2 = isfinite(det2) && det2 != 0.0;
  divide_pair(c1, det1, c2, det2, I1, I2);
  const double eb1[2] = {ne1 ? 0.0 + oe1[0] : 0.0, ne1 ? 0.0 + oe1[1] : 0.0};
  const double eb2[2] = {ne2 ? 0.0 + oe2[0] : 0.0, ne2 ? 0.0 + oe2[1] : 0.0};
  ok1 = ok1 && (isfinite(eb1[0]) & isfinite(eb1[1]));
  ok2 = ok2 && (isfinite(eb2[0]) & isfinite(eb2[1]));
  bool inf1 = false, inf2 = false;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double s0 = M1.m[2 + r][0] * I1[0] + M1.m[2 + r][1] * I1[2];
    const double s1 = M1.m[2 + r][0] * I1[1] + M1.m[2 + r][1] * I1[3];
    const double t0 = M2.m[0 + r][2] * I2[0] + M2.m[0 + r][3] * I2[2];
    const double t1 = M2.m[0 + r][2] * I2[1] + M2.m[0 + r][3] * I2[3];
    eta1[r] = 0.0 - ((0.0 + s0 * eb1[0]) + s1 * eb1[1]);
    eta2[r] = 0.0 - ((0.0 + t0 * eb2[0]) + t1 * eb2[1]);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double v1 = M1.m[2 + r][2 + c] - (s0 * M1.m[0][2 + c] + s1 * M1.m[1][2 + c]);
      const double v2 = M2.m[0 + r][0 + c] - (t0 * M2.m[2][0 + c] + t1 * M2.m[3][0 + c]);
      lam1[r * 2 + c] = v1;
      lam2[r * 2 + c] = v2;
      inf1 |= isinf(v1);
      inf2 |= isinf(v2);
    }
  }
  ok1 = ok1 && !inf1;
  ok2 = ok2 && !inf2;
}

}  // namespace gbp
