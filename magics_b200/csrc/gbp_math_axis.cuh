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
GBP_DEV bool inv_axis(int a, const double (&P)[4], const double (&Q)[4], double (&O)[4]) {
  const bool y = a != 0;
  // entries by their row-major index in the 4x4
  const double m0 = y ? Q[0] : P[0], m2 = y ? Q[1] : P[1], m8 = y ? Q[2] : P[2], m10 = y ? Q[3] : P[3];
  const double m5 = y ? P[0] : Q[0], m7 = y ? P[1] : Q[1], m13 = y ? P[2] : Q[2], m15 = y ? P[3] : Q[3];
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
  divide_all(c, det, O);
  return true;
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
template <int KEEP>
GBP_DEV bool dyn_message_axis(int a, const DynM &M, bool other_nonempty, const double (&oe)[2], const double (&oP)[4],
                              const double (&oQ)[4], double (&eta)[2], double (&lam)[4]) {
  constexpr int A = KEEP * 2, B = (1 - KEEP) * 2;
  double bP[4], bQ[4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double p = M.m[B + r][B + c];
      bP[r * 2 + c] = other_nonempty ? p + oP[r * 2 + c] : p;
      bQ[r * 2 + c] = other_nonempty ? p + oQ[r * 2 + c] : p;
    }
  double I[4];
  if (!inv_axis(a, bP, bQ, I)) return false;
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

}  // namespace gbp
