// gbp_iterate_axis.cuh — the GBP iteration kernel of the DECOUPLED regime: two lanes per variable.
//
// Same half-steps as k_iterate (gbp_iterate.cuh: EXT = FactorGraph::external_{factor,variable}_iteration
// + both deliveries, factorgraph.rs:719-760, 794-826, robot.rs:1814-1858; INT =
// internal_{factor,variable}_iteration, factorgraph.rs:688-714, 762-790), for the robots whose factor
// graph currently has no coupling between the x and the y chain: no InterRobot factor inside its
// safety distance (every one takes InterRobotFactor::skip, interrobot.rs:213-226), a zero Obstacle
// Jacobian (flat SDF under every variable), no Tracking message.  That is every robot of a swarm
// away from obstacles and from each other — all of BASELINE configs 4 and 5 — and it is where
// k_iterate spends 168 registers on 4x4 blocks that are half exact zeros.
//
// Mapping: lane (variable i, axis a); a robot's 2V lanes are consecutive threads of one CTA; the
// lane pair (t, t ^ 1) of a variable sits in one warp and exchanges its 2x2 blocks by shuffle (the
// 4x4 cofactor inverse multiplies entries of both, gbp_math_axis.cuh); variable i reaches i -+ 1
// (threads t -+ 2) through shared memory.  A lane reads and writes only the rows of its axis in the
// tiled store (16 of a record's 24 rows, 12 of a Dynamic message's 20), so the cross rows cost no
// DRAM traffic; they are exact zeros for every robot handled here (Store::mode invariant below).
//
// Everything k_iterate handles that this kernel does not — an InterRobot factor that is not skipped,
// a non-zero Obstacle Jacobian, Tracking, a belief update that is not taken (variable.rs:276-297), an
// Empty Dynamic message, non-finite values, disabled Dynamic factors — is detected BEFORE the robot's
// first store into live state: what a launch reads is never written by it (records and Dynamic messages
// are double buffered: read [p], write [1 - p]) and every store into single-buffered state waits until
// the robot's lanes have voted.  A robot that fails the vote is marked Store::mode = 1 and appended to
// Store::gen_list; k_iterate runs right after over that list, from the same inputs, and rewrites
// whatever this kernel had already put into the [1 - p] buffers.  k_iterate hands a robot back
// (mode = 0) once two internal halves in a row left its state inside the invariant:
//   mode[r] == 0  =>  every cross entry (row + column odd) of r's records and Dynamic messages is an
//   exact zero, every mirror InterRobot message and Tracking message of r is Empty, every Obstacle
//   message of r is Empty or has a zero Jacobian.
// The covariance is not stored here: Store::cov_lazy[vi] = 1 says "VariableBelief.covariance_matrix is
// inv4 of the current precision", which read-back (k_gather_beliefs) and k_iterate evaluate on demand.
#pragma once
#include "gbp_iterate.cuh"
#include "gbp_math_axis.cuh"

namespace gbp {

#ifndef GBP_AXIS_MAXREG
#define GBP_AXIS_MAXREG 80
#endif

static_assert(GBP_TILED == 1, "k_iterate_axis addresses the tiled store");

struct AxisGeom {
  int rpc;      // robots per CTA
  int threads;  // 2 * V * rpc rounded up to whole warps
  size_t smem;
};
// Robots per CTA so that the CTA's 2*V*rpc lanes fill whole warps (a warp's 16 variable slots are
// then 128-byte aligned in every row of the tiled store), at least 128 threads.
inline AxisGeom axis_geom(int V) {
  int g = V, h = 16;
  while (h) {
    const int t = g % h;
    g = h;
    h = t;
  }
  int rpc = 16 / g;
  while (2 * V * rpc < 128) rpc *= 2;
  if (2 * V * rpc > 1024) rpc = 512 / V > 0 ? 512 / V : 1;  // V > 32: whole warps given up
  AxisGeom q;
  q.rpc = rpc;
  q.threads = (2 * V * rpc + 31) / 32 * 32;
  q.smem = size_t(12) * q.threads * sizeof(double) + size_t(q.threads + 2 * rpc) * sizeof(int);
  return q;
}

// Index of row a (vector part) and of row 4 + 5a (this axis' block of the matrix part) of a record
// whose rows are (eta0..3, Lambda row-major, [mu0..3]).
struct AxisRows {
  int64_t v, m;
};
template <int P>
GBP_DEV AxisRows axis_rows(const Store &s, int a, int64_t vi) {
  const int64_t b = s.at<P>(0, vi);
  return {b + a * kTile, b + (4 + 5 * a) * kTile};
}
GBP_DEV void ld_axis(const double *__restrict__ arr, AxisRows q, double (&e)[2], double (&L)[4]) {
  e[0] = arr[q.v];
  e[1] = arr[q.v + 2 * kTile];
  L[0] = arr[q.m];
  L[1] = arr[q.m + 2 * kTile];
  L[2] = arr[q.m + 8 * kTile];
  L[3] = arr[q.m + 10 * kTile];
}
GBP_DEV void st_axis(double *__restrict__ arr, AxisRows q, const double (&e)[2], const double (&L)[4]) {
  arr[q.v] = e[0];
  arr[q.v + 2 * kTile] = e[1];
  arr[q.m] = L[0];
  arr[q.m + 2 * kTile] = L[1];
  arr[q.m + 8 * kTile] = L[2];
  arr[q.m + 10 * kTile] = L[3];
}

template <bool EXT, bool INT>
__global__ void __maxnreg__(GBP_AXIS_MAXREG)
    k_iterate_axis(const __grid_constant__ Store s, const int p, const uint32_t epoch, const int rpc, const int par) {
  extern __shared__ double sh[];
  const int T = blockDim.x, V = s.V;
  const int t = threadIdx.x, a = t & 1, slot = t >> 1;
  const int rl = slot / V, i = slot - rl * V;
  const int64_t r = int64_t(blockIdx.x) * rpc + rl;
  const bool live = rl < rpc && r < s.Nloc;
  const int64_t vi = live ? r * V + i : 0;
  double *const xr = sh;          // [6][T] variable -> Dynamic factor i   (its right-hand factor)
  double *const xl = sh + 6 * T;  // [6][T] variable -> Dynamic factor i-1 (its left-hand factor)
  int *const xne = reinterpret_cast<int *>(sh + 12 * T);  // [T] the variable has sent a message at all
  int *const xbail = xne + T;                             // [rpc] some lane of the robot needs the general kernel
  int *const xflip = xbail + rpc;                         // [rpc] some e_frozen bit of the robot has to change
  if (t < 2 * rpc) xbail[t] = 0;
  __syncthreads();

  const double *const pubr = s.pub[p];
  double *const pubw = s.pub[1 - p];
  const AxisRows qp = axis_rows<kRec>(s, a, vi), qm = axis_rows<20>(s, a, vi);

  // ---- the robot's flags, this lane's prior and running mean ------------------------------------
  bool was_general = false, idle = true, ant = false;
  uint32_t itf = 0u;
  int64_t eo0 = 0, eo1 = 0;
  int32_t nlow = 0;
  bool own_ne = false;
  double pe[2] = {0.0, 0.0}, pl = 0.0, mu[2] = {0.0, 0.0}, pos = 0.0, vel = 0.0;
  if (live) {
    was_general = s.mode[r] != 0;
    idle = s.idle[r] != 0;
    ant = s.antenna[r] != 0;
    const bool latest = s.latest[r] != 0;
    itf = s.iter_factor[r];
    eo0 = s.eoff[r];
    eo1 = s.eoff[r + 1];
    nlow = s.nlow[r];
    own_ne = s.pub_epoch[p][vi] > 0u;
    pe[0] = s.prior_eta[s.at<4>(a, vi)];
    pe[1] = s.prior_eta[s.at<4>(a + 2, vi)];
    pl = s.prior_lam[vi];
    pos = pubr[qp.v + 20 * kTile];
    vel = pubr[qp.v + 22 * kTile];
    // VariableBelief.mean survives in bel_ext after an external half, else in the published record
    mu[0] = pos;
    mu[1] = vel;
    if (latest) {
      mu[0] = s.bel_ext[qp.v + 20 * kTile];
      mu[1] = s.bel_ext[qp.v + 22 * kTile];
    }
  }
  const bool work = live && !was_general;
  const bool do_ext = EXT && work && !idle && ant;
  const bool do_int = INT && work && !idle;
  // Dynamic factors disabled: whatever they sent while enabled stays in the inbox — general kernel
  bool bad = work && !s.en_dyn;

  // ---- stored Dynamic messages: external inbox sum (variable.rs:263-271; FactorId order prior, dyn(i-1),
  // dyn(i) — mirror, Obstacle and Tracking messages contribute nothing in this regime) and the variable ->
  // factor messages of the previous variable iteration, (record - the factor's own last message)
  // (variable.rs:301-330), handed to the neighbouring variables through shared memory
  double ae[2] = {pe[0], pe[1]}, al[4] = {pl, 0.0, 0.0, pl}, Q[4];
  {
    double eRec[2] = {0.0, 0.0}, LRec[4] = {0.0, 0.0, 0.0, 0.0};
    if (INT && work) ld_axis(pubr, qp, eRec, LRec);
    if (INT && work && idle) {
      // idle robot: its record and messages are carried over to the other buffers unchanged
      st_axis(pubw, qp, eRec, LRec);
      pubw[qp.v + 20 * kTile] = pos;
      pubw[qp.v + 22 * kTile] = vel;
      if (a == 0) s.pub_epoch[1 - p][vi] = s.pub_epoch[p][vi];
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const double *const src = side ? s.m_dynR[p] : s.m_dynL[p];
      double *const x = side ? xr : xl;
      double e[2] = {0.0, 0.0}, L[4] = {0.0, 0.0, 0.0, 0.0};
      bool has = false;
      if (work) {
        ld_axis(src, qm, e, L);
        has = !is_empty_marker(src[qm.v - a * kTile]);  // the Empty marker sits in row 0
        if (INT && idle) {
          double *const dst = side ? s.m_dynR[1 - p] : s.m_dynL[1 - p];
          st_axis(dst, qm, e, L);
          if (a == 1) dst[qm.v - kTile] = src[qm.v - kTile];
        }
      }
      if (EXT && do_ext && has) {
#pragma unroll
        for (int k = 0; k < 2; ++k) ae[k] = ae[k] + e[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) al[k] = al[k] + L[k];
      }
      if (INT) {
#pragma unroll
        for (int k = 0; k < 2; ++k) x[k * T + t] = has ? eRec[k] - e[k] : eRec[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[(2 + k) * T + t] = has ? LRec[k] - L[k] : LRec[k];
      }
    }
    if (INT) xne[t] = own_ne ? 1 : 0;
  }

  double mu_sent[2] = {0.0, 0.0}, mu_ext_new = 0.0;
  uint32_t frz = 0u;
  bool flip = false;

  // =================== external half ====================================
  if (EXT) {
    if (do_ext && i >= 1 && eo1 > eo0) {
      // The pair shares the robot's edges: every InterRobot factor must take `skip`; its head is
      // the neighbour's position mean, the record's epoch, radio/idle bits and the edge scalars.
      mu_sent[0] = s.mu_ext[s.at<2>(0, vi)];
      mu_sent[1] = s.mu_ext[s.at<2>(1, vi)];
      const int64_t elow = eo0 + nlow;
      for (int64_t e = eo0 + a; e < eo1; e += 2) {
        const int A = s.enbr[e];
        const int64_t va = int64_t(A) * V + i;
        const double m0 = pubr[s.at<kRec>(20, va)], m1 = pubr[s.at<kRec>(21, va)];
        const uint32_t epochA = s.pub_epoch[p][va];
        const bool act = s.en_ir && s.antenna[A] != 0 && s.idle[A] == 0;
        const uint32_t birth = s.e_birth[e];
        const bool frozen = (s.e_frozen[e] & 1) != 0;
        const double dsafe = s.e_dsafe[e];
        flip |= frozen == act;  // bit 0 has to end up as !act
        if (act) {
          const bool a_ne = epochA > birth;
          const double muA[2] = {a_ne ? m0 : 0.0, a_ne ? m1 : 0.0};
          double mb[2] = {mu_sent[0], mu_sent[1]};
          if (frozen) {
            const int64_t m = e * (V - 1) + (i - 1);
            mb[0] = s.mu_frozen[m];
            mb[1] = s.mu_frozen[s.EV + m];
          }
          if (!interrobot_skip(e < elow, muA, mb, dsafe)) bad = true;
        } else if (!frozen) {
          // undelivered: A's factor keeps the mean it holds while this belief moves on (robot.rs:1851)
          const int64_t k = (e - eo0) >> 1;
          if (k < 32) frz |= 1u << k;
          else bad = true;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) Q[k] = __shfl_xor_sync(0xffffffffu, al[k], 1);
    if (do_ext) {
      if (!belief_axis(a, ae, al, Q, mu)) bad = true;
      mu_ext_new = mu[0];
      itf += 1;
    }
  }

  // =================== internal half ====================================
  if (INT) {
    // linearisation point of the Obstacle factor: the record's position mean, both axes
    const double other_pos = __shfl_xor_sync(0xffffffffu, pos, 1);
    __syncthreads();
    ae[0] = pe[0];
    ae[1] = pe[1];
    al[0] = pl;
    al[1] = 0.0;
    al[2] = 0.0;
    al[3] = pl;
    if (do_int) {
      // inbox sum of the variable iteration in FactorId order: prior, dyn(i-1), dyn(i); each new message is
      // stored at once (the buffers written here are not read in this launch)
      if (i >= 1) {  // Dynamic factor i-1 -> variable i (slot 1); the other message comes from variable i-1
        const int tn = t - 2, tq = (t ^ 1) - 2;
        const double oe[2] = {xr[tn], xr[T + tn]};
        const double oP[4] = {xr[2 * T + tn], xr[3 * T + tn], xr[4 * T + tn], xr[5 * T + tn]};
        const double oQ[4] = {xr[2 * T + tq], xr[3 * T + tq], xr[4 * T + tq], xr[5 * T + tq]};
        double dc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = s.dyn_c[s.at<4>(k, vi - 1)];
        const DynM M = dyn_potential_q(dc[0], dc[1], dc[2], dc[3]);
        double ne[2], nl[4];
        if (dyn_message_axis<1>(a, M, xne[tn] != 0, oe, oP, oQ, ne, nl)) {
          st_axis(s.m_dynL[1 - p], qm, ne, nl);
#pragma unroll
          for (int k = 0; k < 2; ++k) ae[k] = ae[k] + ne[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) al[k] = al[k] + nl[k];
        } else {
          bad = true;
        }
      }
      if (i <= V - 2) {  // Dynamic factor i -> variable i (slot 0); the other message comes from variable i+1
        const int tn = t + 2, tq = (t ^ 1) + 2;
        const double oe[2] = {xl[tn], xl[T + tn]};
        const double oP[4] = {xl[2 * T + tn], xl[3 * T + tn], xl[4 * T + tn], xl[5 * T + tn]};
        const double oQ[4] = {xl[2 * T + tq], xl[3 * T + tq], xl[4 * T + tq], xl[5 * T + tq]};
        double dc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = s.dyn_c[s.at<4>(k, vi)];
        const DynM M = dyn_potential_q(dc[0], dc[1], dc[2], dc[3]);
        double ne[2], nl[4];
        if (dyn_message_axis<0>(a, M, xne[tn] != 0, oe, oP, oQ, ne, nl)) {
          st_axis(s.m_dynR[1 - p], qm, ne, nl);
#pragma unroll
          for (int k = 0; k < 2; ++k) ae[k] = ae[k] + ne[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) al[k] = al[k] + nl[k];
        } else {
          bad = true;
        }
      }
      if (i >= 1 && i <= V - 2) {
        // Tracking factors run from iteration_count.factor >= 10 on (factorgraph.rs:701): general kernel
        if (s.en_trk && itf >= 10u) bad = true;
        if (s.en_obs) {
          // ObstacleFactor (obstacle.rs:129-188, factor/mod.rs:102-128): the Jacobian must come out zero.
          // Same perturb-and-restore sequence as obstacle_update; the pair splits the lookups.
          if (!(isfinite(pos) & isfinite(vel))) bad = true;  // v0 = J.x - h has to be finite
          const double x = own_ne ? (a ? other_pos : pos) : 0.0, y = own_ne ? (a ? pos : other_pos) : 0.0;
          const double delta = s.jac_delta;
          const double h0 = sdf_measure(s, x, y);
          double px = x, py = y;
          px += delta;
          if (a == 0) {
            if (!(sdf_measure(s, px, py) - h0 == 0.0)) bad = true;
          } else {
            px -= delta;
            py += delta;
            const double h2 = sdf_measure(s, px, py);
            py -= delta;
            const double h3 = sdf_measure(s, px, py);
            if (!((h2 - h0 == 0.0) & (h3 - h0 == 0.0))) bad = true;
          }
        }
      }
      itf += 1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) Q[k] = __shfl_xor_sync(0xffffffffu, al[k], 1);
    if (do_int) {
      if (!belief_axis(a, ae, al, Q, mu)) bad = true;
      // the new record (the buffer is not read in this launch; if the robot fails the vote k_iterate rewrites it)
      st_axis(pubw, qp, ae, al);
      pubw[qp.v + 20 * kTile] = mu[0];
      pubw[qp.v + 22 * kTile] = mu[1];
      if (a == 0) s.pub_epoch[1 - p][vi] = epoch;
    }
  }

  // =================== vote, then the stores into live state ================
  if (bad) xbail[rl] = 1;
  if (flip) xflip[rl] = 1;
  __syncthreads();
  const bool bail = live && xbail[rl] != 0;
  if (live && i == 0 && a == 0 && (was_general || bail)) {
    s.gen_list[atomicAdd(&s.gen_count[par], 1)] = int32_t(r);
    if (!was_general) s.mode[r] = 1;
  }
  if (!work || bail) return;

  if (do_ext) {
    s.mu_ext[s.at<2>(a, vi)] = mu_ext_new;
    while (frz) {
      const int k = __ffs(int(frz)) - 1;
      frz &= frz - 1u;
      const int64_t m = (eo0 + a + 2 * int64_t(k)) * (V - 1) + (i - 1);
      s.mu_frozen[m] = mu_sent[0];
      s.mu_frozen[s.EV + m] = mu_sent[1];
    }
    if (xflip[rl]) {
      // delivered edges hold mu_ext again; undelivered ones are (stay) frozen; the robot's lanes share the edges
      for (int64_t e = eo0 + 2 * i + a; e < eo1; e += 2 * V) {
        const int A = s.enbr[e];
        const uint8_t fr = (s.en_ir && s.antenna[A] != 0 && s.idle[A] == 0) ? 0 : 1;
        const uint8_t cur = s.e_frozen[e];
        if ((cur & 1) != fr) s.e_frozen[e] = uint8_t((cur & 2) | fr);  // bit 1 belongs to the collision monitor
      }
    }
    if (!INT) {
      st_axis(s.bel_ext, qp, ae, al);
      s.bel_ext[qp.v + 20 * kTile] = mu[0];
      s.bel_ext[qp.v + 22 * kTile] = mu[1];
      // bel_ext is single buffered and may still hold cross entries from the robot's last spell in
      // k_iterate: rows a and a + 2 of the precision, the other axis' columns
      const int64_t c0 = qp.m + (1 - 2 * a) * kTile;
      s.bel_ext[c0] = 0.0;
      s.bel_ext[c0 + 2 * kTile] = 0.0;
      s.bel_ext[c0 + 8 * kTile] = 0.0;
      s.bel_ext[c0 + 10 * kTile] = 0.0;
    }
  }
  if (a == 0 && (do_ext || do_int)) s.cov_lazy[vi] = 1;
  if (i == 0 && a == 0) {
    s.iter_factor[r] = itf;
    if (do_int) s.latest[r] = 0;
    else if (do_ext) s.latest[r] = 1;
  }
}

}  // namespace gbp
