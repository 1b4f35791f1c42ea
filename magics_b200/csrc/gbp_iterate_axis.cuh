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
#define GBP_AXIS_MAXREG 96  // 5 warps per scheduler either way from 80 to 96 registers (16384 / (32 * 96) = 5.3)
#endif

static_assert(GBP_TILED == 1, "k_iterate_axis addresses the tiled store");

// Edge heads (neighbour slot, birth epoch, frozen bit, safety distance) a robot's lanes stage in shared
// memory for one another; a robot with more neighbours than this goes to the general kernel.
constexpr int kAxisEdges = 32;
#ifndef GBP_AXIS_BATCH
#define GBP_AXIS_BATCH 2  // edges whose neighbour loads a lane has in flight at once (4: spills, 10 % slower)
#endif
constexpr int kAxisBatch = GBP_AXIS_BATCH;
#ifndef GBP_AXIS_MARK_SHFL
#define GBP_AXIS_MARK_SHFL 0   // 1: the y lane gets "message is Empty" from the x lane instead of loading row 0 itself
                               // (4 registers fewer, but +1 % time: profiles/README.md r02j)
#endif
#ifndef GBP_AXIS_RELOAD_SENT
#define GBP_AXIS_RELOAD_SENT 0 // 1: the last delivered mean is read again in the rare freeze path instead of being kept
                               // (4 registers fewer, +3 % time: r02j)
#endif
#ifndef GBP_AXIS_DYN_PAIR
#define GBP_AXIS_DYN_PAIR 0  // 1: the two Dynamic messages of a variable as one straight-line block (two inverses in flight)
#endif
#ifndef GBP_AXIS_ELL
#define GBP_AXIS_ELL 0  // 1: the first kEll edge heads come from Store::ell_* (a fixed-position copy kept by k_ell_fill) in the
                        // FIRST wave of loads, by cp.async straight into the staging arrays, instead of waiting for
                        // eoff[r]; the frozen bit of an edge is read with the neighbour's data in the third wave.
                        // Bit-identical (110 GPU tests), one dependent load level fewer — and 3.4 % slower: 84 B of
                        // spills at the 96-register cap (profiles/README.md r02v6).  Off; the copy is not even built.
#endif
#ifndef GBP_AXIS_PREFETCH
#define GBP_AXIS_PREFETCH 0  // CTAs ahead whose wave-1 rows this CTA pulls into L2 (592 = 148 SMs x 4 resident CTAs:
                             // +1.5 % DRAM bytes, no time gained — profiles/README.md r02e/r02f; off)
#endif

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
  // at most 512 threads (96 registers each): odd V gives up the 128-byte alignment of a warp's variable slots
  if (2 * V * rpc > 512) rpc = 256 / V > 0 ? 256 / V : 1;
#ifdef GBP_AXIS_RPC
  rpc = GBP_AXIS_RPC;  // tuning builds
#endif
  AxisGeom q;
  q.rpc = rpc;
  q.threads = (2 * V * rpc + 31) / 32 * 32;
  // doubles: xr, xl [6][T], Dynamic-factor constants [4][V], staged safety distances; then ints: xne [T],
  // xbail, xflip [rpc], staged neighbour slots and birth epochs; then bytes: staged frozen bits
  q.smem = (size_t(12) * q.threads + size_t(4) * V + size_t(rpc) * kAxisEdges) * sizeof(double) +
           (size_t(q.threads) + 2 * size_t(rpc) + 2 * size_t(rpc) * kAxisEdges) * sizeof(int) + size_t(rpc) * kAxisEdges;
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
#ifndef GBP_AXIS_STREAM
#define GBP_AXIS_STREAM 0  // 1: evict-first loads / stores for rows nobody else reads in this launch
#endif
GBP_DEV double ld_row(const double *p) {
#if GBP_AXIS_STREAM
  return __ldcs(p);
#else
  return *p;
#endif
}
GBP_DEV void st_row(double *p, double v) {
#if GBP_AXIS_STREAM
  __stcs(p, v);
#else
  *p = v;
#endif
}
GBP_DEV void ld_axis(const double *__restrict__ arr, AxisRows q, double (&e)[2], double (&L)[4]) {
  e[0] = ld_row(arr + q.v);
  e[1] = ld_row(arr + q.v + 2 * kTile);
  L[0] = ld_row(arr + q.m);
  L[1] = ld_row(arr + q.m + 2 * kTile);
  L[2] = ld_row(arr + q.m + 8 * kTile);
  L[3] = ld_row(arr + q.m + 10 * kTile);
}
GBP_DEV void st_axis(double *__restrict__ arr, AxisRows q, const double (&e)[2], const double (&L)[4]) {
  st_row(arr + q.v, e[0]);
  st_row(arr + q.v + 2 * kTile, e[1]);
  st_row(arr + q.m, L[0]);
  st_row(arr + q.m + 2 * kTile, L[1]);
  st_row(arr + q.m + 8 * kTile, L[2]);
  st_row(arr + q.m + 10 * kTile, L[3]);
}

// The loads of a launch come in three dependent waves instead of one per use (the kernel is bound by
// load latency, not by bytes: profiles/README.md r02a):
//   1. everything addressed by (robot, variable) alone — flags, edge range, prior, record, stored Dynamic
//      messages, last delivered mean — issued back to back for every live lane, used or not;
//   2. the robot's edge heads, one edge per lane, into shared memory (one __syncthreads);
//   3. per lane, for kAxisBatch of its edges at once, what hangs off the neighbour slot: radio / idle
//      bits, the neighbour variable's position mean and record epoch.
//
// Which robots a launch covers (multi-GPU): PART 0: slots [0, Nloc); PART 2: the same except
// those flagged in `skip` (the interior robots of a shard); PART 1: the *nlist robots of `list` (the border robots,
// whose records the peers wait for: they run first and their halo travels while the interior launch runs).
// PART is a template parameter: the single-GPU launch (PART 0) must not pay for the other two (measured: 4 %).
template <bool EXT, bool INT, int PART>
__global__ void __maxnreg__(GBP_AXIS_MAXREG)
    k_iterate_axis(const __grid_constant__ Store s, const int p, const uint32_t epoch, const int rpc, const int par,
                   const int32_t *__restrict__ list, const int32_t *__restrict__ nlist,
                   const uint8_t *__restrict__ skip) {
  extern __shared__ double sh[];
  const int T = blockDim.x, V = s.V;
  const int t = threadIdx.x, a = t & 1, slot = t >> 1;
  const int rl = slot / V, i = slot - rl * V;
  const int64_t ridx = int64_t(blockIdx.x) * rpc + rl;
  int64_t r = ridx;
  bool live = rl < rpc && ridx < int64_t(s.Nloc);
  if (PART == 1) {
    live = rl < rpc && ridx < int64_t(*nlist);
    r = live ? int64_t(list[ridx]) : 0;
  }
  // PART 2: a flagged (border) robot is skipped, but nothing waits for its flag: it is read with the other
  // per-robot flags of wave 1 and only decides `work` (gating the first loads on it cost 5 % of the launch)
  const int64_t vi = live ? r * V + i : 0;
  double *const xr = sh;                             // [6][T] variable -> Dynamic factor i   (its right-hand factor)
  double *const xl = sh + 6 * T;                     // [6][T] variable -> Dynamic factor i-1 (its left-hand factor)
  double *const dtab = sh + 12 * T;                  // [V][4] Store::dyn_tab
  double *const hd_dsafe = dtab + 4 * V;             // [rpc][kAxisEdges]
  int *const xne = reinterpret_cast<int *>(hd_dsafe + rpc * kAxisEdges);  // [T] the variable has sent a message at all
  int *const xbail = xne + T;                        // [rpc] some lane of the robot needs the general kernel
  int *const xflip = xbail + rpc;                    // [rpc] some e_frozen bit of the robot has to change
  int *const hd_nbr = xflip + rpc;                   // [rpc][kAxisEdges]
  uint32_t *const hd_birth = reinterpret_cast<uint32_t *>(hd_nbr + rpc * kAxisEdges);
  uint8_t *const hd_frozen = reinterpret_cast<uint8_t *>(hd_birth + rpc * kAxisEdges);
  if (t < 2 * rpc) xbail[t] = 0;
  // Programmatic dependent launch: everything above overlapped the tail of the kernel before this one in the
  // stream; nothing it wrote is touched before this point.  The kernel after this one may start launching at once
  // (it waits for this grid's completion at its own griddepcontrol.wait).  No-ops without the launch attribute.
  pdl_wait();
  pdl_launch_dependents();

  const double *const pubr = s.pub[p];
  double *const pubw = s.pub[1 - p];
  const AxisRows qp = axis_rows<kRec>(s, a, vi), qm = axis_rows<20>(s, a, vi);

#if GBP_AXIS_PREFETCH > 0
  // The CTA that will take this one's place on the SM finds its wave-1 rows in L2 (contiguous launches only: CTAs
  // are dispatched in blockIdx order, GBP_AXIS_PREFETCH of them are resident at a time).
  if (PART != 1 && rl < rpc) {
    const int64_t rn = ridx + int64_t(GBP_AXIS_PREFETCH) * rpc;
    if (rn < s.Nloc) {
      const int64_t vn = rn * V + i;
      const AxisRows pp = axis_rows<kRec>(s, a, vn), pm = axis_rows<20>(s, a, vn);
      const double *const rec = s.pub[p];
      if (INT) {
        prefetch_l2(rec + pp.v);
        prefetch_l2(rec + pp.v + 2 * kTile);
        prefetch_l2(rec + pp.m);
        prefetch_l2(rec + pp.m + 2 * kTile);
        prefetch_l2(rec + pp.m + 8 * kTile);
        prefetch_l2(rec + pp.m + 10 * kTile);
      }
      prefetch_l2(rec + pp.v + 20 * kTile);
      prefetch_l2(rec + pp.v + 22 * kTile);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const double *const m = side ? s.m_dynR[p] : s.m_dynL[p];
        prefetch_l2(m + pm.v);
        prefetch_l2(m + pm.v + 2 * kTile);
        prefetch_l2(m + pm.m);
        prefetch_l2(m + pm.m + 2 * kTile);
        prefetch_l2(m + pm.m + 8 * kTile);
        prefetch_l2(m + pm.m + 10 * kTile);
      }
      prefetch_l2(s.prior_eta + s.at<4>(a, vn));
      prefetch_l2(s.prior_eta + s.at<4>(a + 2, vn));
      if (a == 0) prefetch_l2(s.prior_lam + vn);
      if (EXT) prefetch_l2(s.mu_ext + s.at<2>(a, vn));
    }
  }
#endif
  // ---- wave 1 ---------------------------------------------------------------------------------
  bool was_general = false, idle = true, ant = false, latest = false, skipped = false;
  uint32_t itf = 0u;
  int64_t eo0 = 0, eo1 = 0;
  int32_t nlow = 0;
  bool own_ne = false;
  double pe[2] = {0.0, 0.0}, pl = 0.0, mu[2] = {0.0, 0.0}, pos = 0.0, vel = 0.0;
  double mu_sent[2] = {0.0, 0.0};
#if GBP_AXIS_ELL
  const bool use_ell = 2 * V >= kEll;  // every ELL slot has a lane (any V >= 4)
#else
  const bool use_ell = false;
#endif
  double eRec[2] = {0.0, 0.0}, LRec[4] = {0.0, 0.0, 0.0, 0.0};
  double eL[2] = {0.0, 0.0}, LL[4] = {0.0, 0.0, 0.0, 0.0}, eR[2] = {0.0, 0.0}, LR[4] = {0.0, 0.0, 0.0, 0.0};
#if !GBP_AXIS_MARK_SHFL
  double markL = 0.0, markR = 0.0;  // row 0 of the stored Dynamic messages: the Empty marker sits there
#endif
  if (live) {
    eo0 = s.eoff[r];
    eo1 = s.eoff[r + 1];
    was_general = s.mode[r] != 0;
    if (PART == 2) skipped = skip[r] != 0;
    idle = s.idle[r] != 0;
    ant = s.antenna[r] != 0;
    latest = s.latest[r] != 0;
    itf = s.iter_factor[r];
    nlow = s.nlow[r];
    own_ne = s.pub_epoch[p][vi] > 0u;
    ld_axis(s.m_dynL[p], qm, eL, LL);
    ld_axis(s.m_dynR[p], qm, eR, LR);
#if !GBP_AXIS_MARK_SHFL
    markL = a ? s.m_dynL[p][qm.v - kTile] : eL[0];
    markR = a ? s.m_dynR[p][qm.v - kTile] : eR[0];
#endif
    if (INT) ld_axis(pubr, qp, eRec, LRec);
    pos = pubr[qp.v + 20 * kTile];
    vel = pubr[qp.v + 22 * kTile];
    pe[0] = s.prior_eta[s.at<4>(a, vi)];
    pe[1] = s.prior_eta[s.at<4>(a + 2, vi)];
    pl = s.prior_lam[vi];
    if (EXT) {
      mu_sent[0] = s.mu_ext[s.at<2>(0, vi)];
      mu_sent[1] = s.mu_ext[s.at<2>(1, vi)];
#if GBP_AXIS_ELL
      if (use_ell && 2 * i + a < kEll) {  // straight into the staging arrays: no register waits for them
        const int64_t q = r * kEll + (2 * i + a);
        const int k = rl * kAxisEdges + 2 * i + a;  // slots past the robot's last edge hold -1 and are never read
        cp_async4(hd_nbr + k, s.ell_nbr + q);
        cp_async4(hd_birth + k, s.ell_birth + q);
        cp_async8(hd_dsafe + k, s.ell_dsafe + q);
      }
#endif
    }
  }
  // ---- wave 2: edge heads, Dynamic-factor constants ------------------------------------------------
  const int ne_all = int(eo1 - eo0);
  const int ne = ne_all < kAxisEdges ? ne_all : kAxisEdges;
  if (EXT && live) {
    for (int k = 2 * i + a; k < ne; k += 2 * V) {
      if (use_ell && k < kEll) continue;
      const int64_t e = eo0 + k;
      hd_nbr[rl * kAxisEdges + k] = s.enbr[e];
      hd_birth[rl * kAxisEdges + k] = s.e_birth[e];
      if (!use_ell) hd_frozen[rl * kAxisEdges + k] = s.e_frozen[e];
      hd_dsafe[rl * kAxisEdges + k] = s.e_dsafe[e];
    }
  }
  if (INT && s.dyn_tab)
    for (int k = t; k < 4 * V; k += T) dtab[k] = s.dyn_tab[k];
  // VariableBelief.mean survives in bel_ext after an external half, else in the published record
  mu[0] = pos;
  mu[1] = vel;
  if (live && latest) {
    mu[0] = s.bel_ext[qp.v + 20 * kTile];
    mu[1] = s.bel_ext[qp.v + 22 * kTile];
  }
#if GBP_AXIS_ELL
  if (EXT) {
    cp_async_commit();
    cp_async_wait<0>();
  }
#endif
  __syncthreads();

  const bool work = live && !was_general && !skipped;
  const bool do_ext = EXT && work && !idle && ant;
  const bool do_int = INT && work && !idle;
  // Dynamic factors disabled: whatever they sent while enabled stays in the inbox — general kernel
  bool bad = work && (!s.en_dyn || ne_all > kAxisEdges);

  // ---- wave 3 issued, then the stored Dynamic messages are consumed while it is in flight ----------
  // The pair shares the robot's edges (lane a takes edges a, a + 2, ...): every InterRobot factor must
  // take `skip`; its head is the neighbour's position mean, the record's epoch, radio / idle bits and
  // the staged edge scalars.
  const bool edges = EXT && do_ext && i >= 1 && ne > a;
  const int nmine = edges ? (ne - a + 1) >> 1 : 0;
  double m0[kAxisBatch] = {}, m1[kAxisBatch] = {};
  uint32_t epA[kAxisBatch] = {};
  unsigned onA = 0u;  // bit q: neighbour q of the batch has its radio on and is not idle
  unsigned frzA = 0u; // bit q: edge q of the batch is frozen (use_ell: read here, else staged with the heads)
  auto fetch = [&](int b0) {
    onA = 0u;
    frzA = 0u;
#pragma unroll
    for (int q = 0; q < kAxisBatch; ++q) {
      const int k = (b0 + q < nmine) ? a + 2 * (b0 + q) : a;  // past the end: a valid edge again, result unused
      const int A = hd_nbr[rl * kAxisEdges + k];
      if (use_ell) frzA |= (s.e_frozen[eo0 + k] & 1) ? (1u << q) : 0u;
      const int64_t va = int64_t(A) * V + i;
      m0[q] = pubr[s.at<kRec>(20, va)];
      m1[q] = pubr[s.at<kRec>(21, va)];
      epA[q] = s.pub_epoch[p][va];
      onA |= ((s.antenna[A] != 0) & (s.idle[A] == 0)) ? (1u << q) : 0u;
    }
  };
  if (edges) fetch(0);

  // external inbox sum (variable.rs:263-271; FactorId order prior, dyn(i-1), dyn(i) — mirror, Obstacle and
  // Tracking messages contribute nothing in this regime) and the variable -> factor messages of the previous
  // variable iteration, (record - the factor's own last message) (variable.rs:301-330), handed to the
  // neighbouring variables through shared memory
  double ae[2] = {pe[0], pe[1]}, al[4] = {pl, 0.0, 0.0, pl}, Q[4];
  {
#if GBP_AXIS_MARK_SHFL
    // the Empty marker sits in row 0 of a stored message, which the x lane holds as eL[0] / eR[0]; the y lane asks it
    const unsigned even = (threadIdx.x & 31u) & ~1u;
    const unsigned ne_bits = (is_empty_marker(eL[0]) ? 0u : 1u) | (is_empty_marker(eR[0]) ? 0u : 2u);
    const unsigned pair_bits = __shfl_sync(0xffffffffu, ne_bits, even);
    const bool hasL = work && (pair_bits & 1u), hasR = work && (pair_bits & 2u);
#else
    const bool hasL = work && !is_empty_marker(markL), hasR = work && !is_empty_marker(markR);
#endif
    if (INT && work && idle) {
      // idle robot: its record and messages are carried over to the other buffers unchanged
      st_axis(pubw, qp, eRec, LRec);
      pubw[qp.v + 20 * kTile] = pos;
      pubw[qp.v + 22 * kTile] = vel;
      if (a == 0) s.pub_epoch[1 - p][vi] = s.pub_epoch[p][vi];
      st_axis(s.m_dynL[1 - p], qm, eL, LL);  // row 0 (the marker, if Empty) travels with the x lane
      st_axis(s.m_dynR[1 - p], qm, eR, LR);
    }
    if (EXT && do_ext) {
      if (hasL) {
#pragma unroll
        for (int k = 0; k < 2; ++k) ae[k] = ae[k] + eL[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) al[k] = al[k] + LL[k];
      }
      if (hasR) {
#pragma unroll
        for (int k = 0; k < 2; ++k) ae[k] = ae[k] + eR[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) al[k] = al[k] + LR[k];
      }
    }
    if (INT) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        xl[k * T + t] = hasL ? eRec[k] - eL[k] : eRec[k];
        xr[k * T + t] = hasR ? eRec[k] - eR[k] : eRec[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xl[(2 + k) * T + t] = hasL ? LRec[k] - LL[k] : LRec[k];
        xr[(2 + k) * T + t] = hasR ? LRec[k] - LR[k] : LRec[k];
      }
      xne[t] = own_ne ? 1 : 0;
    }
  }

  double mu_ext_new = 0.0;
  uint32_t frz = 0u;
  bool flip = false;

  // =================== external half ====================================
  if (EXT) {
    for (int b0 = 0; b0 < nmine; b0 += kAxisBatch) {
      if (b0) fetch(b0);
#pragma unroll
      for (int q = 0; q < kAxisBatch; ++q) {
        if (b0 + q >= nmine) break;
        const int k = a + 2 * (b0 + q);
        const bool act = s.en_ir && ((onA >> q) & 1u);
        const uint32_t birth = hd_birth[rl * kAxisEdges + k];
        const bool frozen = use_ell ? ((frzA >> q) & 1u) != 0u : (hd_frozen[rl * kAxisEdges + k] & 1) != 0;
        const double dsafe = hd_dsafe[rl * kAxisEdges + k];
        flip |= frozen == act;  // bit 0 has to end up as !act
        if (act) {
          const bool a_ne = epA[q] > birth;
          const double muA[2] = {a_ne ? m0[q] : 0.0, a_ne ? m1[q] : 0.0};
          double mb[2] = {mu_sent[0], mu_sent[1]};
          if (frozen) {
            const int64_t m = (eo0 + k) * (V - 1) + (i - 1);
            mb[0] = s.mu_frozen[m];
            mb[1] = s.mu_frozen[s.EV + m];
          }
          if (!interrobot_skip(k < nlow, muA, mb, dsafe)) bad = true;
        } else if (!frozen) {
          // undelivered: A's factor keeps the mean it holds while this belief moves on (robot.rs:1851)
          frz |= 1u << (b0 + q);
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
    // ObstacleFactor (obstacle.rs:129-188, factor/mod.rs:102-128): the Jacobian must come out zero.  Same
    // perturb-and-restore sequence as obstacle_update; the pair splits the four lookups — the x lane takes
    // h(x, y) and h(x + d, y), the y lane h(x', y + d) and h(x', y') with x' = (x + d) - d, y' = (y + d) - d.
    bool obs_here = false;
    double h_first = 0.0, h_second = 0.0;
    if (s.en_obs && do_int && i >= 1 && i <= V - 2) {
      obs_here = true;
      const double x = own_ne ? (a ? other_pos : pos) : 0.0, y = own_ne ? (a ? pos : other_pos) : 0.0;
      const double delta = s.jac_delta;
      double px = x + delta, py = y;
      double x1 = x, y1 = y, x2 = px, y2 = y;
      if (a) {
        px -= delta;
        py += delta;
        x1 = px;
        y1 = py;
        py -= delta;
        x2 = px;
        y2 = py;
      }
      h_first = sdf_measure(s, x1, y1);
      h_second = sdf_measure(s, x2, y2);
    }
    // h(x, y) is the x lane's first lookup; J0 = (h(x + d, y) - h0) / d, J1 and J2 = J3 likewise
    const double h0 = __shfl_sync(0xffffffffu, h_first, (threadIdx.x & 31u) & ~1u);
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
#if GBP_AXIS_DYN_PAIR
      {
        // Dynamic factor i-1 -> variable i (slot 1, "L": the other message comes from variable i-1) and Dynamic
        // factor i -> variable i (slot 0, "R": from variable i+1), evaluated together (dyn_message_axis_pair).  The
        // first / last variable has only one of them and evaluates it twice.
        const bool doL = i >= 1, doR = i <= V - 2;
        const double *const x1 = doL ? xr : xl, *const x2 = doR ? xl : xr;
        const int d1 = doL ? -2 : 2, d2 = doR ? 2 : -2;
        const int tn1 = t + d1, tq1 = (t ^ 1) + d1, tn2 = t + d2, tq2 = (t ^ 1) + d2;
        const int f1 = doL ? i - 1 : i, f2 = doR ? i : i - 1;  // factor index of the constants
        const double oe1[2] = {x1[tn1], x1[T + tn1]}, oe2[2] = {x2[tn2], x2[T + tn2]};
        const double oP1[4] = {x1[2 * T + tn1], x1[3 * T + tn1], x1[4 * T + tn1], x1[5 * T + tn1]};
        const double oQ1[4] = {x1[2 * T + tq1], x1[3 * T + tq1], x1[4 * T + tq1], x1[5 * T + tq1]};
        const double oP2[4] = {x2[2 * T + tn2], x2[3 * T + tn2], x2[4 * T + tn2], x2[5 * T + tn2]};
        const double oQ2[4] = {x2[2 * T + tq2], x2[3 * T + tq2], x2[4 * T + tq2], x2[5 * T + tq2]};
        double dc1[4], dc2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dc1[k] = s.dyn_tab ? dtab[4 * f1 + k] : s.dyn_c[s.at<4>(k, vi - i + f1)];
          dc2[k] = s.dyn_tab ? dtab[4 * f2 + k] : s.dyn_c[s.at<4>(k, vi - i + f2)];
        }
        const DynM M1 = dyn_potential_q(dc1[0], dc1[1], dc1[2], dc1[3]);
        const DynM M2 = dyn_potential_q(dc2[0], dc2[1], dc2[2], dc2[3]);
        double e1[2], l1[4], e2[2], l2[4];
        bool ok1, ok2;
        if (doL && doR) {
          dyn_message_axis_pair(a, M1, xne[tn1] != 0, oe1, oP1, oQ1, M2, xne[tn2] != 0, oe2, oP2, oQ2, e1, l1, e2, l2, ok1,
                                ok2);
        } else if (doL) {
          ok1 = dyn_message_axis<1>(a, M1, xne[tn1] != 0, oe1, oP1, oQ1, e1, l1);
          ok2 = true;
        } else {
          ok2 = dyn_message_axis<0>(a, M2, xne[tn2] != 0, oe2, oP2, oQ2, e2, l2);
          ok1 = true;
        }
        if (doL) {
          if (ok1) {
            st_axis(s.m_dynL[1 - p], qm, e1, l1);
#pragma unroll
            for (int k = 0; k < 2; ++k) ae[k] = ae[k] + e1[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) al[k] = al[k] + l1[k];
          } else {
            bad = true;
          }
        }
        if (doR) {
          if (ok2) {
            st_axis(s.m_dynR[1 - p], qm, e2, l2);
#pragma unroll
            for (int k = 0; k < 2; ++k) ae[k] = ae[k] + e2[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) al[k] = al[k] + l2[k];
          } else {
            bad = true;
          }
        }
      }
#else
      if (i >= 1) {  // Dynamic factor i-1 -> variable i (slot 1); the other message comes from variable i-1
        const int tn = t - 2, tx = (t & ~1) - 2, ty = (t | 1) - 2;  // own axis, x lane, y lane of variable i-1
        const double oe[2] = {xr[tn], xr[T + tn]};
        const double oX[4] = {xr[2 * T + tx], xr[3 * T + tx], xr[4 * T + tx], xr[5 * T + tx]};
        const double oY[4] = {xr[2 * T + ty], xr[3 * T + ty], xr[4 * T + ty], xr[5 * T + ty]};
        double dc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = s.dyn_tab ? dtab[4 * (i - 1) + k] : s.dyn_c[s.at<4>(k, vi - 1)];
        const DynM M = dyn_potential_q(dc[0], dc[1], dc[2], dc[3]);
        double ne_[2], nl[4];
        if (dyn_message_axis_xy<1>(a, M, xne[tn] != 0, oe, oX, oY, ne_, nl)) {
          st_axis(s.m_dynL[1 - p], qm, ne_, nl);
#pragma unroll
          for (int k = 0; k < 2; ++k) ae[k] = ae[k] + ne_[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) al[k] = al[k] + nl[k];
        } else {
          bad = true;
        }
      }
      if (i <= V - 2) {  // Dynamic factor i -> variable i (slot 0); the other message comes from variable i+1
        const int tn = t + 2, tx = (t & ~1) + 2, ty = (t | 1) + 2;
        const double oe[2] = {xl[tn], xl[T + tn]};
        const double oX[4] = {xl[2 * T + tx], xl[3 * T + tx], xl[4 * T + tx], xl[5 * T + tx]};
        const double oY[4] = {xl[2 * T + ty], xl[3 * T + ty], xl[4 * T + ty], xl[5 * T + ty]};
        double dc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = s.dyn_tab ? dtab[4 * i + k] : s.dyn_c[s.at<4>(k, vi)];
        const DynM M = dyn_potential_q(dc[0], dc[1], dc[2], dc[3]);
        double ne_[2], nl[4];
        if (dyn_message_axis_xy<0>(a, M, xne[tn] != 0, oe, oX, oY, ne_, nl)) {
          st_axis(s.m_dynR[1 - p], qm, ne_, nl);
#pragma unroll
          for (int k = 0; k < 2; ++k) ae[k] = ae[k] + ne_[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) al[k] = al[k] + nl[k];
        } else {
          bad = true;
        }
      }
#endif
      if (i >= 1 && i <= V - 2) {
        // Tracking factors run from iteration_count.factor >= 10 on (factorgraph.rs:701): general kernel
        if (s.en_trk && itf >= 10u) bad = true;
        if (obs_here) {
          if (!(isfinite(pos) & isfinite(vel))) bad = true;  // v0 = J.x - h has to be finite
          if (!((a ? h_first - h0 == 0.0 : true) & (h_second - h0 == 0.0))) bad = true;
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
  if (live && !skipped && i == 0 && a == 0 && (was_general || bail)) {
    s.gen_list[atomicAdd(&s.gen_count[par], 1)] = int32_t(r);
    if (!was_general) s.mode[r] = 1;
  }
  if (!work || bail) return;

  if (do_ext) {
    if (frz) {
      // the mean these factors keep is the one delivered last (still in mu_ext: read again, rare path)
#if GBP_AXIS_RELOAD_SENT
      const double sent0 = s.mu_ext[s.at<2>(0, vi)], sent1 = s.mu_ext[s.at<2>(1, vi)];
#else
      const double sent0 = mu_sent[0], sent1 = mu_sent[1];
#endif
      while (frz) {
        const int k = __ffs(int(frz)) - 1;
        frz &= frz - 1u;
        const int64_t m = (eo0 + a + 2 * int64_t(k)) * (V - 1) + (i - 1);
        s.mu_frozen[m] = sent0;
        s.mu_frozen[s.EV + m] = sent1;
      }
    }
    // both lanes of the pair (same branch: do_ext is per robot) must have read mu_ext before either overwrites its
    // component
#if GBP_AXIS_RELOAD_SENT
    __syncwarp(3u << ((threadIdx.x & 31u) & ~1u));
#endif
    s.mu_ext[s.at<2>(a, vi)] = mu_ext_new;
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
