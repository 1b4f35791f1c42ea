// host_math.cpp — C entry points over the DEVICE math header compiled for the host, so the
// per-node arithmetic of the CUDA engine can be checked against the oracle without a GPU.
#define GBP_HOST_MATH_TEST 1
#include "../../magics_b200/csrc/gbp_math.cuh"
#include "../../magics_b200/csrc/gbp_math_axis.cuh"
#include "../../magics_b200/csrc/gbp_collide.cuh"

extern "C" {
int hm_inv4(const double *m, double *out) {
  double a[16], o[16] = {0};
  for (int k = 0; k < 16; ++k) a[k] = m[k];
  const bool ok = gbp::inv4(a, o);
  for (int k = 0; k < 16; ++k) out[k] = o[k];
  return ok ? 1 : 0;
}
int hm_divide_all16(const double *c, double det, double *out) {
  double a[16], o[16];
  for (int k = 0; k < 16; ++k) a[k] = c[k];
  gbp::divide_all(a, det, o);
  for (int k = 0; k < 16; ++k) out[k] = o[k];
  return 0;
}
// returns 0: not taken, 1: taken (cov, valid written; mu only if valid)
int hm_belief_moments(const double *eta, const double *lam, double *mu, double *cov, int *valid) {
  double e[4], l[16], m[4], c[16];
  for (int k = 0; k < 4; ++k) e[k] = eta[k], m[k] = mu[k];
  for (int k = 0; k < 16; ++k) l[k] = lam[k], c[k] = cov[k];
  bool v = false;
  const bool t = gbp::belief_moments(e, l, m, c, v);
  for (int k = 0; k < 4; ++k) mu[k] = m[k];
  for (int k = 0; k < 16; ++k) cov[k] = c[k];
  *valid = v ? 1 : 0;
  return t ? 1 : 0;
}
// Dynamic factor message to slot `keep` given the other variable's message (eta4, Lambda16) or none.
int hm_dyn_message(int keep, double dt, double qs, int other_nonempty, const double *other, double *eta, double *lam) {
  const gbp::DynM M = gbp::dyn_potential(dt, qs);
  double o[20], e[4], l[16];
  for (int k = 0; k < 20; ++k) o[k] = other ? other[k] : 0.0;
  const bool ok = keep ? gbp::dyn_message<1>(M, other_nonempty != 0, o, e, l) : gbp::dyn_message<0>(M, other_nonempty != 0, o, e, l);
  for (int k = 0; k < 4; ++k) eta[k] = e[k];
  for (int k = 0; k < 16; ++k) lam[k] = l[k];
  return ok ? 1 : 0;
}
// InterRobot factor of robot A toward robot B, message to B's variable (the one the engine pulls).
// rec_a = A's published record (eta4, Lambda16); returns 0 for Message::empty() (skip / singular / inf).
int hm_interrobot_message(int a_first, const double *mu_a, const double *mu_b, int a_nonempty, const double *rec_a,
                          double dsafe, double tiny, double lm, double *eta2, double *lam4) {
  double ma[2] = {mu_a[0], mu_a[1]}, mb[2] = {mu_b[0], mu_b[1]}, rec[20], e[2], l[4];
  for (int k = 0; k < 20; ++k) rec[k] = rec_a[k];
  const bool ok = gbp::interrobot_message(a_first != 0, ma, mb, a_nonempty != 0, rec, dsafe, tiny, lm, e, l);
  for (int k = 0; k < 2; ++k) eta2[k] = e[k];
  for (int k = 0; k < 4; ++k) lam4[k] = l[k];
  return ok ? 1 : 0;
}
int hm_interrobot_skip(int a_first, const double *mu_a, const double *mu_b, double dsafe) {
  double ma[2] = {mu_a[0], mu_a[1]}, mb[2] = {mu_b[0], mu_b[1]};
  return gbp::interrobot_skip(a_first != 0, ma, mb, dsafe) ? 1 : 0;
}

// ---- gbp_math_axis.cuh: one axis of the decoupled regime (block = (m[a][a], m[a][a+2], m[a+2][a], m[a+2][a+2])) ----
int hm_inv_axis(int a, const double *P, const double *Q, double *O) {
  double p[4], q[4], o[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) p[k] = P[k], q[k] = Q[k];
  const bool ok = gbp::inv_axis(a, p, q, o);
  for (int k = 0; k < 4; ++k) O[k] = o[k];
  return ok ? 1 : 0;
}
int hm_belief_axis(int a, const double *e, const double *P, const double *Q, double *mu) {
  double ee[2] = {e[0], e[1]}, p[4], q[4], m[2] = {mu[0], mu[1]};
  for (int k = 0; k < 4; ++k) p[k] = P[k], q[k] = Q[k];
  const bool ok = gbp::belief_axis(a, ee, p, q, m);
  mu[0] = m[0];
  mu[1] = m[1];
  return ok ? 1 : 0;
}
int hm_dyn_message_axis(int keep, int a, double dt, double qs, int other_nonempty, const double *oe, const double *oP,
                        const double *oQ, double *eta, double *lam) {
  const gbp::DynM M = gbp::dyn_potential(dt, qs);
  double e[2] = {oe[0], oe[1]}, p[4], q[4], re[2] = {0, 0}, rl[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) p[k] = oP[k], q[k] = oQ[k];
  const bool ok = keep ? gbp::dyn_message_axis<1>(a, M, other_nonempty != 0, e, p, q, re, rl)
                       : gbp::dyn_message_axis<0>(a, M, other_nonempty != 0, e, p, q, re, rl);
  for (int k = 0; k < 2; ++k) eta[k] = re[k];
  for (int k = 0; k < 4; ++k) lam[k] = rl[k];
  return ok ? 1 : 0;
}
// gbp_collide.cuh: the device predicate of the environment-collision monitor (collider_hits_ball), one collider
// (kind, tx, ty, cos, sin, radius, hx, hy, first vertex, vertex count) against n robot balls.
int hm_collider_hits(int kind, float tx, float ty, float re, float im, float radius, float hx, float hy, int nv,
                     const float *verts, int n, const float *xz, float robot_radius, unsigned char *out) {
  const gbp::ColliderDev c{kind, tx, ty, re, im, radius, hx, hy, 0, nv};
  for (int k = 0; k < n; ++k) out[k] = gbp::collider_hits_ball(c, verts, xz[2 * k], xz[2 * k + 1], robot_radius) ? 1 : 0;
  return 0;
}
}
