// gbp_iterate.cuh — the fused GBP iteration kernel.
//
// One thread per (robot, variable index i); a robot's V threads sit in one warp
// (32/V robots per warp).  Thread i owns variable i, the messages variable i
// holds from Dynamic factors i-1 and i, Obstacle factor i, Tracking factor i and
// from every neighbour's InterRobot factor at i.
//
//   EXT half  = FactorGraph::external_factor_iteration   (factorgraph.rs:719-760)
//             + delivery                                  (robot.rs:1814-1831)
//             + FactorGraph::external_variable_iteration  (factorgraph.rs:794-826)
//             + delivery                                  (robot.rs:1843-1858)
//   INT half  = FactorGraph::internal_factor_iteration    (factorgraph.rs:688-714)
//             + FactorGraph::internal_variable_iteration  (factorgraph.rs:762-790)
//
// The reference pushes every InterRobot message from the factor's owner A to the
// other robot B.  Here B PULLS: the only thing A's factor needs from A is the
// message A's variable sent it, which is A's published belief record (the entry
// for A's own InterRobot factors in A's variable inboxes is permanently Empty,
// factorgraph.rs:745-753, so the variable answers with its full belief,
// variable.rs:305-306).  So B evaluates A's factor itself from pub[p][A] and the
// position mean B last sent it, and the message never leaves B's registers
// except as the "mirror" record B keeps for later sums.  The only cross-robot
// traffic of a sub-step is therefore the read of neighbours' pub records, and a
// launch boundary is needed only between INT(t) and EXT(t): EXT(t) and INT(t+1)
// run fused in one launch.
#pragma once
#include "gbp_math.cuh"
#include "gbp_store.cuh"
#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace gbp {

#ifndef GBP_ITER_BLOCK
#define GBP_ITER_BLOCK 128
#endif
#ifndef GBP_ITER_MIN_BLOCKS
#define GBP_ITER_MIN_BLOCKS 3
#endif
constexpr int kIterBlock = GBP_ITER_BLOCK;

// ---- L2 prefetch -----------------------------------------------------------------------------
// The kernel is a chain of ~16 dependent load phases per thread at 12 warps per SM, so the
// loaded DRAM latency of each phase is exposed.  A thread knows at entry every address it will
// read later in its own records, so it asks L2 for them up front: prefetch.global.L2 costs no
// register and no shared memory, and turns the later own-record phases into L2 hits.  (Measured,
// profiles/README.md r01d: +5 %; prefetching the neighbours' records as well, staging through
// shared memory with cp.async, or parking accumulators in shared memory for 16 warps/SM were
// all slower than this; so was, r01f, letting a robot's lanes split its edges and pull each
// neighbour's position means / epoch / radio bits into L1 ahead of the edge loop: +-0; and sending
// the read-once records around L1 with ld.global.cg: -10 %; staging the heads of 8 edges at once in
// per-thread shared-memory slots (two dependent round trips per robot instead of K): -9 %; loading
// the Dynamic-factor constants and the linearisation mean with the first internal batch: -1.5 %.)
#ifndef GBP_PREFETCH
#define GBP_PREFETCH 1
#endif
#ifndef GBP_MIRROR_MASK
#define GBP_MIRROR_MASK 1
#endif

// Programmatic dependent launch (sm_90+): wait for the grids this one depends on / let the next grid of the stream
// start.  Compiled in with -DGBP_PDL=1 only: measured (profiles/README.md r02q, r02r) it hides 0.7 % of a tick on a
// 125 k-robot shard and nothing on 1 M robots, while the barrier the asm puts into k_iterate_axis' prologue costs
// that kernel 1.5 %.
#ifndef GBP_PDL
#define GBP_PDL 0
#endif
GBP_DEV void pdl_wait() {
#if GBP_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
GBP_DEV void pdl_launch_dependents() {
#if GBP_PDL
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}

GBP_DEV void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// first N components of a P-component per-variable record
template <int P, int N>
GBP_DEV void prefetch_planes(const Store &s, const double *base, int64_t vi) {
#pragma unroll
  for (int k = 0; k < N; ++k) prefetch_l2(base + s.at<P>(k, vi));
}

GBP_DEV double dot2(const double (&a)[2], const double (&b)[2]) { return (0.0 + a[0] * b[0]) + a[1] * b[1]; }
GBP_DEV double norm2(double x, double y) { return sqrt((0.0 + x * x) + y * y); }

// ObstacleFactor::measure (factor/obstacle.rs:141-188).
GBP_DEV double sdf_measure(const Store &s, double x_pos, double y_pos, uint32_t *opx = nullptr,
                           uint32_t *opy = nullptr) {
  // offsets world / 2.0 and scales f64::from(image dimension) / world, evaluated once on the host
  // (refresh_scalars: one IEEE division each, the same bits as evaluating them here per call)
  const double x_offset = s.sdf_xo, y_offset = s.sdf_yo;
  const double x_scale = s.sdf_xs, y_scale = s.sdf_ys;
  const uint32_t xp = sat_u32((x_pos + x_offset) * x_scale);
  const uint32_t yp = sat_u32((-y_pos + y_offset) * y_scale);
  if (opx) *opx = xp;
  if (opy) *opy = yp;
  if (!(xp < uint32_t(s.sdf_w) && yp < uint32_t(s.sdf_h))) return 0.0;
  const uint8_t red = s.sdf[size_t(yp) * size_t(s.sdf_w) + xp];
  return 1.0 - double(red) / 255.0;
}

// Sum of the messages variable i holds from its own non-InterRobot factors, in
// FactorId order dyn(i-1) < dyn(i) < obs(i) < trk(i) (id.rs:25-61; creation order
// robot.rs:1228-1334), added onto (ae, al) (variable.rs:263-271).
// One stored Dynamic-factor message (eta4, Lambda16) added onto (ae, al) unless it is Empty.
// The record is loaded whole before its Empty marker is looked at: the slots of factors a variable
// does not have (dyn(i-1) of variable 0, ...) hold the marker for ever, so no index test is needed
// and no load waits on another load.
GBP_DEV void add_dyn_stored(const Store &s, const double *__restrict__ arr, int64_t vi, double (&ae)[4],
                            double (&al)[16]) {
  double m[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) m[k] = arr[s.at<20>(k, vi)];
  if (!is_empty_marker(m[0])) {
#pragma unroll
    for (int k = 0; k < 4; ++k) ae[k] = ae[k] + m[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) al[k] = al[k] + m[4 + k];
  }
}
// The stored Obstacle and Tracking messages (factored form, expanded on use).
// Returns bit 0: the Obstacle message contributed (non-zero Jacobian or non-finite v0), bit 1: a Tracking
// message is stored — either one couples the variable's x and y chains (k_iterate_axis must not run it).
GBP_DEV unsigned add_unary_stored(const Store &s, int64_t vi, double (&ae)[4], double (&al)[16]) {
  double o[4], t[3];
  unsigned contributed = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = s.m_obs[s.at<4>(k, vi)];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = s.m_trk[s.at<3>(k, vi)];
  // A zero Jacobian (every variable away from obstacle edges) with a finite v0 adds exact zeros to all
  // 20 sums: skipped — also keeps x and y decoupled for the inverse that follows.
  if (!is_empty_marker(o[0]) && !((o[0] == 0.0) & (o[1] == 0.0) & (o[2] == 0.0) & isfinite(o[3]))) {
    const double J[4] = {o[0], o[1], o[2], o[2]};
    unary_add(J, o[3], s.lm_obs, ae, al);
    contributed |= 1u;
  }
  if (!is_empty_marker(t[0])) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double g = t[k] * s.lm_trk;
      ae[k] = ae[k] + g * t[2];
#pragma unroll
      for (int l = 0; l < 2; ++l) al[k * 4 + l] = al[k * 4 + l] + g * t[l];
    }
    contributed |= 2u;
  }
  return contributed;
}
GBP_DEV unsigned add_internal(const Store &s, int p, int64_t vi, double (&ae)[4], double (&al)[16]) {
  add_dyn_stored(s, s.m_dynL[p], vi, ae, al);
  add_dyn_stored(s, s.m_dynR[p], vi, ae, al);
  return add_unary_stored(s, vi, ae, al);
}

// Store::cov_lazy resolved: VariableBelief.covariance_matrix = inv4 of the variable's current precision, which
// k_iterate_axis took and found finite (belief_axis) but did not store.  Out of line: rare (a robot's first
// launch after it left k_iterate_axis, FactorGraph::reset_variables).
__device__ __noinline__ void materialise_cov(const Store &s, int p, int64_t r, int64_t vi) {
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  double lam[16], cov[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lam[k] = src[s.at<kRec>(4 + k, vi)];
    cov[k] = 0.0;
  }
  inv4(lam, cov);
#pragma unroll
  for (int k = 0; k < 16; ++k) s.cov[s.at<16>(k, vi)] = cov[k];
  s.valid[vi] = 1;
  s.cov_lazy[vi] = 0;
}

// Adds the stored mirror message m if there is one; returns whether there was.
GBP_DEV bool add_mirror(const Store &s, int64_t m, double (&ae)[4], double (&al)[16]) {
  double v[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = s.mir[k * s.EV + m];
  if (is_empty_marker(v[0])) return false;
  ae[0] = ae[0] + v[0];
  ae[1] = ae[1] + v[1];
  al[0] = al[0] + v[2];
  al[1] = al[1] + v[3];
  al[4] = al[4] + v[4];
  al[5] = al[5] + v[5];
  return true;
}

// ---- mirror messages streamed through shared memory --------------------------------------------------------
// A variable's inbox sum walks its InterRobot edges in FactorId order; the adds are a serial chain by definition, but
// in a loop of add_mirror calls every add also waits for ITS edge's six loads (profiles/r02y2: 52 % of k_iterate's
// stall samples in a dense swarm sit on the Empty-marker test).  All addresses are known up front (m = e (V - 1) +
// (i - 1)), so the loads run kMirDepth edges ahead as cp.async copies into a per-thread strip of shared memory: no
// registers are held while they are in flight.  A thread reads only what its own copies wrote (cp.async.wait_group).
#ifndef GBP_MIR_STREAM
#define GBP_MIR_STREAM 1
#endif
#ifndef GBP_MIR_DEPTH
#define GBP_MIR_DEPTH 4  // 4: 36.8 ms, 8: 37.3 ms, 12: 44.1 ms per dense tick (r02v3); 24 KiB of shared memory per CTA
#endif
constexpr int kMirDepth = GBP_MIR_DEPTH;
constexpr size_t kIterSmemBytes = GBP_MIR_STREAM ? size_t(kMirDepth) * 6 * GBP_ITER_BLOCK * sizeof(double) : 0;
GBP_DEV void cp_async8(double *smem, const double *gmem) {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem)
               : "memory");
#endif
}
GBP_DEV void cp_async4(void *smem, const void *gmem) {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem)
               : "memory");
#endif
}
GBP_DEV void cp_async_commit() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
GBP_DEV void cp_async_wait() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
// For every edge e of [lo, hi) in order: `pre(e)` (sums that sit between two mirror messages in the inbox order), then,
// unless `skip(e)`, `use(e, v)` with the six stored doubles of the variable's mirror message; `pre(hi)` at the end.
template <class Pre, class Skip, class Use>
GBP_DEV void stream_mirrors(const Store &s, double *strip, int64_t lo, int64_t hi, int Vm1, int im1, Pre &&pre,
                            Skip &&skip, Use &&use) {
#if GBP_MIR_STREAM
  const int T = GBP_ITER_BLOCK;
  static_assert((kMirDepth & (kMirDepth - 1)) == 0, "GBP_MIR_DEPTH must be a power of two");
  auto issue = [&](int64_t e) {
    if (e < hi && !skip(e)) {
      const int slot = int(unsigned(e - lo) & unsigned(kMirDepth - 1));
      const int64_t m = e * Vm1 + im1;
#pragma unroll
      for (int k = 0; k < 6; ++k) cp_async8(strip + (slot * 6 + k) * T, s.mir + k * s.EV + m);
    }
    cp_async_commit();  // an empty group keeps the count of groups in flight the same for every edge
  };
  for (int d = 0; d < kMirDepth; ++d) issue(lo + d);
  for (int64_t e = lo; e < hi; ++e) {
    pre(e);
    cp_async_wait<kMirDepth - 1>();
    if (!skip(e)) {
      const int slot = int(unsigned(e - lo) & unsigned(kMirDepth - 1));
      double v[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = strip[(slot * 6 + k) * T];
      use(e, v);
    }
    issue(e + kMirDepth);  // into the slot just read (the reads above have been consumed)
  }
  cp_async_wait<0>();
  pre(hi);
#else
  for (int64_t e = lo; e < hi; ++e) {
    pre(e);
    if (skip(e)) continue;
    double v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = s.mir[k * s.EV + (e * Vm1 + im1)];
    use(e, v);
  }
  pre(hi);
#endif
}
// add_mirror on values already loaded.
GBP_DEV bool add_mirror_values(const double (&v)[6], double (&ae)[4], double (&al)[16]) {
  if (is_empty_marker(v[0])) return false;
  ae[0] = ae[0] + v[0];
  ae[1] = ae[1] + v[1];
  al[0] = al[0] + v[2];
  al[1] = al[1] + v[3];
  al[4] = al[4] + v[4];
  al[5] = al[5] + v[5];
  return true;
}

// The Dynamic messages are double buffered like the published records (an internal half reads [p] and
// writes [1 - p], so that a robot k_iterate_axis gives up on is still untouched); a variable that does not
// renew them — idle robot, Dynamic factors disabled — carries them over.
GBP_DEV void copy_dyn_messages(const Store &s, int p, int64_t vi) {
#pragma unroll 4
  for (int k = 0; k < 20; ++k) {
    s.m_dynL[1 - p][s.at<20>(k, vi)] = s.m_dynL[p][s.at<20>(k, vi)];
    s.m_dynR[1 - p][s.at<20>(k, vi)] = s.m_dynR[p][s.at<20>(k, vi)];
  }
}

GBP_DEV void load_prior(const Store &s, int64_t vi, double (&ae)[4], double (&al)[16]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) ae[k] = s.prior_eta[s.at<4>(k, vi)];
  const double pl = s.prior_lam[vi];
#pragma unroll
  for (int k = 0; k < 16; ++k) al[k] = (k % 5 == 0) ? pl : 0.0;
}

GBP_DEV void shfl_vec(double (&v)[20], bool &flag, int delta_up, unsigned lane) {
  // delta_up > 0: take from lane - delta; < 0: take from lane + |delta|
#pragma unroll
  for (int k = 0; k < 20; ++k)
    v[k] = delta_up > 0 ? __shfl_up_sync(0xffffffffu, v[k], delta_up)
                        : __shfl_down_sync(0xffffffffu, v[k], -delta_up);
  const int f = flag ? 1 : 0;
  flag = (delta_up > 0 ? __shfl_up_sync(0xffffffffu, f, delta_up)
                       : __shfl_down_sync(0xffffffffu, f, -delta_up)) != 0;
}

// TrackingFactor::skip + measure + jacobian (factor/tracking.rs:171-381) for
// variable vi of robot r at linearisation point x.  Writes the message record.
GBP_DEV void tracking_update(const Store &s, int64_t r, int64_t vi, const double (&x)[4]) {
  bool skip = false;
  int32_t timeout = s.trk_timeout[vi];
  if (timeout >= 0) {
    if (timeout == 0) {
      timeout = -1;
    } else {
      timeout -= 1;
      skip = true;
    }
    s.trk_timeout[vi] = timeout;
  }
  const int32_t w0 = s.wp_off[r];
  const uint32_t npath = uint32_t(s.wp_off[r + 1] - w0);
  uint32_t rec = s.trk_record[vi];
  if (!skip && (npath < 2 || rec >= npath - 1)) skip = true;
  if (skip) {
    s.m_trk[s.at<3>(0, vi)] = empty_marker();
    return;
  }
  const float *wp = s.wp_xy + 2 * size_t(w0);
  const double x_pos[2] = {x[0], x[1]};
  const double cs[2] = {double(wp[2 * rec]), double(wp[2 * rec + 1])};
  const double ce[2] = {double(wp[2 * rec + 2]), double(wp[2 * rec + 3])};
  const double line[2] = {ce[0] - cs[0], ce[1] - cs[1]};
  const double rel[2] = {x_pos[0] - cs[0], x_pos[1] - cs[1]};
  const double t = dot2(rel, line) / dot2(line, line);
  const double cur[2] = {cs[0] + t * line[0], cs[1] + t * line[1]};
  const double d0 = s.trk_switch_padding, d1 = d0 * 0.01;
  const double cur_to_end = norm2(ce[0] - cur[0], ce[1] - cur[1]);
  bool have_prev = false;
  double pp[2] = {0.0, 0.0};
  if (rec > 0) {
    const double ps[2] = {double(wp[2 * rec - 2]), double(wp[2 * rec - 1])};
    const double pe[2] = {cs[0], cs[1]};  // previous_end is the current start (same f32 pair)
    const double pl[2] = {pe[0] - ps[0], pe[1] - ps[1]};
    const double prel[2] = {x_pos[0] - ps[0], x_pos[1] - ps[1]};
    const double tp = dot2(prel, pl) / dot2(pl, pl);
    pp[0] = ps[0] + tp * pl[0];
    pp[1] = ps[1] + tp * pl[1];
    const double cur_to_prev_end = norm2(pe[0] - cur[0], pe[1] - cur[1]);
    const double prev_to_prev_end = norm2(cs[0] - pp[0], cs[1] - pp[1]);
    have_prev = cur_to_prev_end < d0 && cur_to_prev_end > d1 && prev_to_prev_end < d0;
  }
  if (cur_to_end < d0) {  // Tracking::increment_record (tracking.rs:55-65)
    rec = min(rec + 1u, npath - 2u);
    s.trk_record[vi] = rec;
  }
  double mp[2];
  if (have_prev) {
#pragma unroll
    for (int k = 0; k < 2; ++k) mp[k] = x_pos[k] + ((cur[k] - x_pos[k]) + (pp[k] - x_pos[k]));
  } else {
    double ln[2] = {line[0], line[1]};
    const double mag = norm2(line[0], line[1]);
    if (!(mag == 0.0 || isinf(mag))) {
      ln[0] /= mag;
      ln[1] /= mag;
    }
    const double vn = norm2(x[2], x[3]);
#pragma unroll
    for (int k = 0; k < 2; ++k) mp[k] = cur[k] + ln[k] * vn / 5.0;
  }
  const double dist = norm2(mp[0] - x_pos[0], mp[1] - x_pos[1]);
  const double ad = s.trk_attraction;
  const double meas = dist < ad ? dist / ad : 1.0;
  const float lx = float(mp[0]), ly = float(mp[1]);
  s.trk_last[s.at<2>(0, vi)] = lx;
  s.trk_last[s.at<2>(1, vi)] = ly;
  s.trk_value[vi] = meas;
  // jacobian (tracking.rs:171-194) from the measurement just stored
  const double j0 = (1.0 / meas) * (x_pos[0] - double(lx));
  const double j1 = (1.0 / meas) * (x_pos[1] - double(ly));
  const double v0 = ((0.0 + j0 * x[0]) + j1 * x[1]) + (0.0 - meas);
  s.m_trk[s.at<3>(0, vi)] = j0;
  s.m_trk[s.at<3>(1, vi)] = j1;
  s.m_trk[s.at<3>(2, vi)] = v0;
}

// ObstacleFactor update: measure + first_order_jacobian (factor/mod.rs:102-128,
// obstacle.rs:129-188) at linearisation point x; six SDF lookups, perturb and
// restore sequence kept so that x+d-d rounding reaches the same pixels.
GBP_DEV void obstacle_update(const Store &s, int64_t vi, const double (&x)[4]) {
  const double h = sdf_measure(s, x[0], x[1]);
  const double h0 = sdf_measure(s, x[0], x[1]);
  const double delta = s.jac_delta;
  double px = x[0], py = x[1];
  px += delta;
  const double h1 = sdf_measure(s, px, py);
  px -= delta;
  py += delta;
  const double h2 = sdf_measure(s, px, py);
  py -= delta;
  const double h3 = sdf_measure(s, px, py);  // columns 2 and 3 perturb the velocity only
  // three quotients by the same delta; bit-identical to the three divisions.  (A plain `0.0 / delta`
  // — every robot away from obstacles — takes the division's slow path: 2.5 % of the kernel's
  // instructions in the r01f profile.)
  const double num[3] = {h1 - h0, h2 - h0, h3 - h0};
  double quo[3];
  divide_all(num, delta, quo);
  const double j0 = quo[0], j1 = quo[1], j2 = quo[2];
  const double v0 = ((((0.0 + j0 * x[0]) + j1 * x[1]) + j2 * x[2]) + j2 * x[3]) + (0.0 - h);
  s.m_obs[s.at<4>(0, vi)] = j0;
  s.m_obs[s.at<4>(1, vi)] = j1;
  s.m_obs[s.at<4>(2, vi)] = j2;
  s.m_obs[s.at<4>(3, vi)] = v0;
}

// Head of one InterRobot edge (receiver r <- neighbour A) for variable i: everything needed to
// decide whether A's factor sends a message at all — A's position mean, the record's epoch, A's
// radio/idle bits and the edge scalars: 10 small loads with one dependent level (A = enbr[e]).
// A's (eta, Lambda) — 20 doubles — is fetched only when the factor is not skipped
// (InterRobotFactor::skip, interrobot.rs:213-226): in a swarm most robots within comms range
// are outside safety range.  (Loading the head one edge ahead, or two heads at once in front of one
// non-unrolled edge body, was measured slower: registers / spills.)
struct EdgeHead {
  double mu0, mu1, dsafe;
  uint64_t rnum;
  uint32_t epochA, birth;
  int A;
  bool act, frozen;
};
GBP_DEV void load_head(const Store &s, const double *__restrict__ pubr, int p, int64_t e, int A, int V, int i,
                       EdgeHead &h) {
  const int64_t va = int64_t(A) * V + i;
  h.A = A;
  h.mu0 = pubr[s.at<kRec>(20, va)];
  h.mu1 = pubr[s.at<kRec>(21, va)];
  h.epochA = s.pub_epoch[p][va];
  h.act = s.en_ir && s.antenna[A] != 0 && s.idle[A] == 0;
  h.birth = s.e_birth[e];
  h.frozen = (s.e_frozen[e] & 1) != 0;
  h.rnum = s.e_rnum[e];
  h.dsafe = s.e_dsafe[e];
}

// External half, one edge (receiver r <- neighbour A) for variable i: evaluates A's InterRobot
// factor from A's published record (or keeps the stored mirror message when nothing is delivered)
// and adds the message to the running inbox sum (ae, al).
GBP_DEV void ext_edge(const Store &s, const double *__restrict__ pubr, const EdgeHead &h, int64_t e, int64_t e0,
                      int64_t elow, int V, int i, const double (&mu_sent)[2], double (&ae)[4], double (&al)[16],
                      uint64_t &mir_ne) {
  const int64_t m = e * (V - 1) + (i - 1);
  if (h.act) {
    const bool a_ne = h.epochA > h.birth;
    const double muA[2] = {a_ne ? h.mu0 : 0.0, a_ne ? h.mu1 : 0.0};
    double mb[2] = {mu_sent[0], mu_sent[1]};
    if (h.frozen) {  // rare: edge just created, or A's radio was off at the last delivery
      mb[0] = s.mu_frozen[m];
      mb[1] = s.mu_frozen[s.EV + m];
    }
    bool ok = false;
    double me[2], ml[4];
    if (!interrobot_skip(e < elow, muA, mb, h.dsafe)) {
      double rec[20];
      const int64_t va = int64_t(h.A) * V + i;
      const uint64_t rnum = h.rnum;
#pragma unroll
      for (int k = 0; k < 20; ++k) rec[k] = pubr[s.at<kRec>(k, va)];
      const double tiny = s.tiny_scale * double(rnum + uint64_t(i - 1));
      ok = interrobot_message(e < elow, muA, mb, a_ne, rec, h.dsafe, tiny, s.lm_ir, me, ml);
    }
    if (ok) {
      s.mir[m] = me[0];
      s.mir[s.EV + m] = me[1];
      s.mir[2 * s.EV + m] = ml[0];
      s.mir[3 * s.EV + m] = ml[1];
      s.mir[4 * s.EV + m] = ml[2];
      s.mir[5 * s.EV + m] = ml[3];
      ae[0] = ae[0] + me[0];
      ae[1] = ae[1] + me[1];
      al[0] = al[0] + ml[0];
      al[1] = al[1] + ml[1];
      al[4] = al[4] + ml[2];
      al[5] = al[5] + ml[3];
      if (e - e0 < 64) mir_ne |= 1ull << (e - e0);
    } else {
      s.mir[m] = empty_marker();
    }
  } else {
    // undelivered: the variable keeps the old message
    if (add_mirror(s, m, ae, al) && e - e0 < 64) mir_ne |= 1ull << (e - e0);
    // ... and A's factor keeps the mean it already holds from this variable
    // while this variable's belief moves on (robot.rs:1851): freeze it
    if (!h.frozen) {
      s.mu_frozen[m] = mu_sent[0];
      s.mu_frozen[s.EV + m] = mu_sent[1];
    }
  }
}

// ---- InterRobot messages of an external half, one thread per (edge, variable) -------------------------------
// The general kernel below keeps a robot's V variables in V lanes of one warp; evaluating the neighbours' InterRobot
// factors inside it makes every lane walk the robot's K edges one after the other — K dependent load chains and, for
// every factor inside its safety distance, a 20-double record fetch and a Schur complement, at 168 registers and
// 3 warps per scheduler (profiles/r02y: 43 % long-scoreboard, 31 % instruction-fetch stalls in a dense swarm).  The
// messages of one external half do not depend on each other, so they are computed FIRST, by this kernel, with one
// thread per (edge, variable) — K (V - 1) independent threads per robot, a third of the registers, 2 kB of code —
// and stored where the receiver keeps them anyway (Store::mir).  k_iterate then adds the stored messages in inbox
// order; the sums see the same doubles in the same order as when they came from registers.
// Also leaves Store::e_act[e] = "the neighbour's radio is on and it is not idle" (what decides between delivery and
// freeze, robot.rs:1843-1858) so that k_iterate reads one byte per edge instead of chasing the neighbour slot.
// par >= 0: the robots of Store::gen_list (length gen_count[par]); par < 0: every robot of the shard.
#ifndef GBP_EDGE_SPLIT
#define GBP_EDGE_SPLIT 1
#endif
constexpr int kEdgeBlock = 128;
#ifndef GBP_EDGE_MIN_BLOCKS
#define GBP_EDGE_MIN_BLOCKS 4  // 126 registers without a cap: 16 warps / SM
#endif
GBP_DEV void edge_messages_body(const Store &s, const int p, const int par, const int chunks) {
  // `chunks` warps share a robot's (edge, variable) pairs: 1 for a large swarm, up to 32 for a few dozen robots
  // with dozens of neighbours each (the reference's Circle Experiment), where a warp per robot would leave the
  // GPU to 30 warps walking 18 rounds each
  const int V = s.V, Vm1 = V - 1;
  const unsigned lane = threadIdx.x & 31u;
  const int64_t nrob = par >= 0 ? int64_t(s.gen_count[par]) : int64_t(s.Nloc);
  const int64_t wstride = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const double *const pubr = s.pub[p];
  for (int64_t item = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; item < nrob * chunks; item += wstride) {
    const int64_t k = item / chunks;
    const int chunk = int(item - k * chunks);
    const int64_t r = par >= 0 ? int64_t(s.gen_list[k]) : k;
    if (s.idle[r] != 0 || s.antenna[r] == 0) continue;  // no external half for this robot (robot.rs:1800-1812)
    const int64_t e0 = s.eoff[r], e1 = s.eoff[r + 1], elow = e0 + s.nlow[r];
    const int64_t npairs = (e1 - e0) * Vm1;
    for (int64_t q = chunk * 32 + lane; q < npairs; q += 32 * chunks) {
      const int64_t eo = q / Vm1;
      const int i = 1 + int(q - eo * Vm1);
      const int64_t e = e0 + eo, m = e0 * Vm1 + q;  // == e * (V - 1) + (i - 1)
      const int A = s.enbr[e];
      const bool act = s.en_ir && s.antenna[A] != 0 && s.idle[A] == 0;
      if (i == 1) s.e_act[e] = act ? 1 : 0;
      const int64_t va = int64_t(A) * V + i, vi = r * V + i;
      if (!act) {
        // undelivered: the receiver keeps the message it has, and A's factor keeps the mean it already holds from
        // this variable while this variable's belief moves on (robot.rs:1851): freeze it
        if (!(s.e_frozen[e] & 1)) {
          s.mu_frozen[m] = s.mu_ext[s.at<2>(0, vi)];
          s.mu_frozen[s.EV + m] = s.mu_ext[s.at<2>(1, vi)];
        }
        continue;
      }
      const double a0 = pubr[s.at<kRec>(20, va)], a1 = pubr[s.at<kRec>(21, va)];
      const uint32_t epochA = s.pub_epoch[p][va], birth = s.e_birth[e];
      const bool frozen = (s.e_frozen[e] & 1) != 0;
      const double dsafe = s.e_dsafe[e];
      const uint64_t rnum = s.e_rnum[e];
      double mb[2] = {s.mu_ext[s.at<2>(0, vi)], s.mu_ext[s.at<2>(1, vi)]};
      if (frozen) {  // rare: edge just created, or A's radio was off at the last delivery
        mb[0] = s.mu_frozen[m];
        mb[1] = s.mu_frozen[s.EV + m];
      }
      const bool a_ne = epochA > birth;
      const double muA[2] = {a_ne ? a0 : 0.0, a_ne ? a1 : 0.0};
      bool ok = false;
      double me[2], ml[4];
      if (!interrobot_skip(e < elow, muA, mb, dsafe)) {
        double rec[20];
#pragma unroll
        for (int c = 0; c < 20; ++c) rec[c] = pubr[s.at<kRec>(c, va)];
        const double tiny = s.tiny_scale * double(rnum + uint64_t(i - 1));
        ok = interrobot_message(e < elow, muA, mb, a_ne, rec, dsafe, tiny, s.lm_ir, me, ml);
      }
      if (ok) {
        s.mir[m] = me[0];
        s.mir[s.EV + m] = me[1];
        s.mir[2 * s.EV + m] = ml[0];
        s.mir[3 * s.EV + m] = ml[1];
        s.mir[4 * s.EV + m] = ml[2];
        s.mir[5 * s.EV + m] = ml[3];
      } else {
        s.mir[m] = empty_marker();
      }
    }
  }
}
__global__ void __launch_bounds__(kEdgeBlock, GBP_EDGE_MIN_BLOCKS)
    k_edge_messages(const __grid_constant__ Store s, const int p, const int par, const int chunks) {
  edge_messages_body(s, p, par, chunks);
}

#ifdef GBP_ITER_MAXREG  // experiments: exact register cap instead of a CTAs-per-SM target
#define GBP_ITER_BOUNDS __maxnreg__(GBP_ITER_MAXREG)
#else
#define GBP_ITER_BOUNDS __launch_bounds__(kIterBlock, GBP_ITER_MIN_BLOCKS)
#endif
// One warp's robots (32/V of them, lane = rl * V + i) through one launch.  `requalify`: the robots come
// from Store::gen_list; a robot whose state satisfies the invariant of Store::mode 0 again after this
// launch is handed back to k_iterate_axis.
template <bool EXT, bool INT>
GBP_DEV void iterate_warp(const Store &s, const int p, const uint32_t epoch, const int64_t r, const bool live,
                          const int rl, const int i, const unsigned lane, const bool requalify, double *strip) {
  const int V = s.V;
  const int64_t vi = live ? r * V + i : 0;
  if (live && s.cov_lazy[vi]) materialise_cov(s, p, r, vi);
  // light: this lane's variable ends the launch in the decoupled regime (gbp_iterate_axis.cuh)
  bool light = INT && s.en_dyn;

  const double *const pubr = s.pub[p];
  double *const pubw = s.pub[1 - p];
  // One batch of independent loads: the robot's flags and both candidates for the running mean
  // (VariableBelief.mean survives an update that cannot invert, variable.rs:276-297; it lives in
  // bel_ext after an external half, else in the published record).
  bool idle = true, ant = false;
  double mu[4] = {0.0, 0.0, 0.0, 0.0};
  uint32_t itf = 0u;
  int64_t eo0 = 0, eo1 = 0;
  int32_t nlow = 0;
  if (live) {
#if GBP_PREFETCH
    prefetch_planes<20, 20>(s, s.m_dynL[p], vi);
    prefetch_planes<20, 20>(s, s.m_dynR[p], vi);
    prefetch_planes<4, 4>(s, s.m_obs, vi);
    prefetch_planes<3, 3>(s, s.m_trk, vi);
    prefetch_planes<4, 4>(s, s.prior_eta, vi);
    prefetch_l2(s.prior_lam + vi);
    if (INT) {
      prefetch_planes<kRec, 20>(s, pubr, vi);
      prefetch_planes<4, 4>(s, s.dyn_c, vi);
    }
#endif
    const uint8_t f_idle = s.idle[r], f_ant = s.antenna[r], f_latest = s.latest[r];
    itf = s.iter_factor[r];
    eo0 = s.eoff[r];
    eo1 = s.eoff[r + 1];
    nlow = s.nlow[r];
    double ma[4], mb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = pubr[s.at<kRec>(20 + k, vi)];
      mb[k] = s.bel_ext[s.at<24>(20 + k, vi)];
    }
    idle = f_idle != 0;
    ant = f_ant != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) mu[k] = f_latest ? mb[k] : ma[k];
  }
  const bool do_ext = EXT && live && !idle && ant;
  const bool do_int = INT && live && !idle;

  // bit (e - e0): the mirror message of edge e is non-empty after the external half, so the
  // internal half of the same launch need not read the Empty ones back (edges beyond 64: it reads)
  uint64_t mir_ne = 0;

  // =================== external half ====================================
  if (do_ext) {
    double ae[4], al[16];
    load_prior(s, vi, ae, al);
    const int64_t e0 = eo0;
    const int64_t e1 = (i >= 1) ? eo1 : e0;  // variable 0 has no InterRobot factors
    const int64_t elow = e0 + nlow;          // edges [e0, elow) have a lower robot id than r
    // where the own factors' messages sit in the inbox order (id.rs:25-61): after the mirror
    // factors of lower-id robots, before those of higher-id robots
    const int64_t eadd = elow < e1 ? elow : e1;
#if !GBP_EDGE_SPLIT
    double mu_sent[2] = {0.0, 0.0};
    if (e1 > e0) {
      mu_sent[0] = s.mu_ext[s.at<2>(0, vi)];
      mu_sent[1] = s.mu_ext[s.at<2>(1, vi)];
    }
#endif
#if GBP_EDGE_SPLIT
    // the neighbours' factors have been evaluated by k_edge_messages: every edge's message (new, or the one kept
    // because nothing was delivered) sits in Store::mir
    stream_mirrors(
        s, strip, e0, e1, V - 1, i - 1,
        [&](int64_t e) {
          if (e == eadd) add_internal(s, p, vi, ae, al);
        },
        [](int64_t) { return false; },
        [&](int64_t e, const double(&v)[6]) {
          // (an undelivered edge's mean was frozen by k_edge_messages)
          if (add_mirror_values(v, ae, al) && e - e0 < 64) mir_ne |= 1ull << (e - e0);
        });
#else
    int A_next = (e0 < e1) ? s.enbr[e0] : 0;
    for (int64_t e = e0;; ++e) {
      if (e == eadd) add_internal(s, p, vi, ae, al);
      if (e >= e1) break;
      EdgeHead h;
      load_head(s, pubr, p, e, A_next, V, i, h);
      if (e + 1 < e1) A_next = s.enbr[e + 1];
      ext_edge(s, pubr, h, e, e0, elow, V, i, mu_sent, ae, al, mir_ne);
    }
#endif
    double cov[16];
    bool valid = false;
    const bool taken = belief_moments(ae, al, mu, cov, valid);
    if (taken) {
#pragma unroll
      for (int k = 0; k < 16; ++k) s.cov[s.at<16>(k, vi)] = cov[k];
      s.valid[vi] = valid ? 1 : 0;
    }
    light = light && taken && valid && mir_ne == 0ull && e1 - e0 <= 64;
    s.mu_ext[s.at<2>(0, vi)] = mu[0];
    s.mu_ext[s.at<2>(1, vi)] = mu[1];
    if (!INT || !do_int) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s.bel_ext[s.at<24>(k, vi)] = ae[k];
#pragma unroll
      for (int k = 0; k < 16; ++k) s.bel_ext[s.at<24>(4 + k, vi)] = al[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) s.bel_ext[s.at<24>(20 + k, vi)] = mu[k];
    }
    itf += 1;
  }
  if (EXT) {
    __syncwarp();
    if (do_ext) {
      // delivered edges hold mu_ext again; undelivered ones are (stay) frozen; the robot's lanes
      // share the edges (every lane has finished reading e_frozen: __syncwarp above)
      for (int64_t e = eo0 + i; e < eo1; e += V) {
#if GBP_EDGE_SPLIT
        const uint8_t fr = s.e_act[e] ? 0 : 1;
#else
        const int A = s.enbr[e];
        const uint8_t fr = (s.en_ir && s.antenna[A] != 0 && s.idle[A] == 0) ? 0 : 1;
#endif
        const uint8_t cur = s.e_frozen[e];
        if ((cur & 1) != fr) s.e_frozen[e] = uint8_t((cur & 2) | fr);  // bit 1 belongs to the collision monitor
      }
    }
  }

  // =================== internal half ====================================
  if (INT) {
    // ---- variable -> Dynamic factor messages of the previous variable
    // iteration, rebuilt as (record - factor's own last message) (variable.rs:301-330)
    double toR[20], toL[20];
    bool own_ne = false;
    if (do_int) {
      own_ne = s.pub_epoch[p][vi] > 0u;
      double R[20];
#pragma unroll
      for (int k = 0; k < 20; ++k) R[k] = pubr[s.at<kRec>(k, vi)];
#pragma unroll
      for (int k = 0; k < 20; ++k) toR[k] = s.m_dynR[p][s.at<20>(k, vi)];
#pragma unroll
      for (int k = 0; k < 20; ++k) toL[k] = s.m_dynL[p][s.at<20>(k, vi)];
      const bool hasR = !is_empty_marker(toR[0]), hasL = !is_empty_marker(toL[0]);
#pragma unroll
      for (int k = 0; k < 20; ++k) {
        toR[k] = hasR ? R[k] - toR[k] : R[k];
        toL[k] = hasL ? R[k] - toL[k] : R[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 20; ++k) toR[k] = toL[k] = 0.0;
    }
    bool fromL_ne = own_ne, fromR_ne = own_ne;
    shfl_vec(toR, fromL_ne, 1, lane);    // lane i receives var i-1 -> dyn(i-1)
    shfl_vec(toL, fromR_ne, -1, lane);   // lane i receives var i+1 -> dyn(i)
    if (do_int) {
      // Inbox sum of the variable iteration that follows (variable.rs:263-271), in FactorId order:
      // prior, mirror messages of lower-id robots, dyn(i-1), dyn(i), obstacle, tracking, mirror
      // messages of higher-id robots.  The two Dynamic messages are added straight from the registers
      // they were computed in.
      double ae[4], al[16];
      const int64_t e0 = eo0;
      const int64_t e1 = (i >= 1) ? eo1 : e0;
      const int64_t elow = e0 + nlow;
      const int64_t eadd = elow < e1 ? elow : e1;
      bool any_mir = false;
      load_prior(s, vi, ae, al);
      auto known_empty = [&](int64_t e) {
#if GBP_MIRROR_MASK
        return EXT && do_ext && e - e0 < 64 && !((mir_ne >> (e - e0)) & 1ull);
#else
        return false;
#endif
      };
      stream_mirrors(
          s, strip, e0, eadd, V - 1, i - 1, [](int64_t) {}, known_empty,
          [&](int64_t, const double(&v)[6]) { any_mir |= add_mirror_values(v, ae, al); });
      if (s.en_dyn) {
        if (i >= 1) {  // Dynamic factor i-1 -> variable i (slot 1)
          double dcL[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) dcL[k] = s.dyn_c[s.at<4>(k, vi - 1)];
          const DynM M = dyn_potential_q(dcL[0], dcL[1], dcL[2], dcL[3]);
          double ne[4], nl[16];
          if (dyn_message<1>(M, fromL_ne, toR, ne, nl)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s.m_dynL[1 - p][s.at<20>(k, vi)] = ne[k];
#pragma unroll
            for (int k = 0; k < 16; ++k) s.m_dynL[1 - p][s.at<20>(4 + k, vi)] = nl[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) ae[k] = ae[k] + ne[k];
#pragma unroll
            for (int k = 0; k < 16; ++k) al[k] = al[k] + nl[k];
            light = light && (nl[1] == 0.0) & (nl[3] == 0.0) & (nl[4] == 0.0) & (nl[6] == 0.0) & (nl[9] == 0.0) &
                             (nl[11] == 0.0) & (nl[12] == 0.0) & (nl[14] == 0.0);
          } else {
            s.m_dynL[1 - p][s.at<20>(0, vi)] = empty_marker();
            light = false;
          }
        }  // variable 0 has no dyn(i-1): its slot holds the Empty marker for ever
        if (i <= V - 2) {  // Dynamic factor i -> variable i (slot 0)
          double dcR[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) dcR[k] = s.dyn_c[s.at<4>(k, vi)];
          const DynM M = dyn_potential_q(dcR[0], dcR[1], dcR[2], dcR[3]);
          double ne[4], nl[16];
          if (dyn_message<0>(M, fromR_ne, toL, ne, nl)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s.m_dynR[1 - p][s.at<20>(k, vi)] = ne[k];
#pragma unroll
            for (int k = 0; k < 16; ++k) s.m_dynR[1 - p][s.at<20>(4 + k, vi)] = nl[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) ae[k] = ae[k] + ne[k];
#pragma unroll
            for (int k = 0; k < 16; ++k) al[k] = al[k] + nl[k];
            light = light && (nl[1] == 0.0) & (nl[3] == 0.0) & (nl[4] == 0.0) & (nl[6] == 0.0) & (nl[9] == 0.0) &
                             (nl[11] == 0.0) & (nl[12] == 0.0) & (nl[14] == 0.0);
          } else {
            s.m_dynR[1 - p][s.at<20>(0, vi)] = empty_marker();
            light = false;
          }
        }  // likewise the last variable and dyn(i)
      }
      else {  // Dynamic factors disabled: whatever they sent while enabled is still in the inbox
        add_dyn_stored(s, s.m_dynL[p], vi, ae, al);
        add_dyn_stored(s, s.m_dynR[p], vi, ae, al);
        copy_dyn_messages(s, p, vi);
      }
      if (i >= 1 && i <= V - 2 && (s.en_obs || s.en_trk)) {
        // linearisation point = mean of the variable's last message (factor/mod.rs:336-349)
        double x[4], x0[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          x[k] = pubr[s.at<kRec>(20 + k, vi)];
          x0[k] = own_ne ? x[k] : 0.0;  // Obstacle inbox starts Empty (factorgraph.rs:310-322)
        }
        if (s.en_obs) obstacle_update(s, vi, x0);
        // Tracking factors are skipped until iteration_count.factor >= 10
        // (factorgraph.rs:701); their inbox starts with the variable's belief
        if (s.en_trk && itf >= 10u) {
          light = false;
          double xt[4];
          const bool has = own_ne || s.trk_seed[vi] != 0;  // emptied by reset_variables: linearise at zeros
#pragma unroll
          for (int k = 0; k < 4; ++k) xt[k] = has ? x[k] : 0.0;
          tracking_update(s, r, vi, xt);
        }
      }
      itf += 1;

      // ---- belief update + new record (variable.rs:251-297)
      const unsigned unary = add_unary_stored(s, vi, ae, al);
      stream_mirrors(
          s, strip, eadd, e1, V - 1, i - 1, [](int64_t) {}, known_empty,
          [&](int64_t, const double(&v)[6]) { any_mir |= add_mirror_values(v, ae, al); });
      double cov[16];
      bool valid = false;
      const bool taken = belief_moments(ae, al, mu, cov, valid);
      if (taken) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s.cov[s.at<16>(k, vi)] = cov[k];
        s.valid[vi] = valid ? 1 : 0;
      }
      light = light && taken && valid && unary == 0u && !any_mir && (al[1] == 0.0) & (al[3] == 0.0) & (al[4] == 0.0) &
                           (al[6] == 0.0) & (al[9] == 0.0) & (al[11] == 0.0) & (al[12] == 0.0) & (al[14] == 0.0) &
                           isfinite(ae[0]) & isfinite(ae[1]) & isfinite(ae[2]) & isfinite(ae[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) pubw[s.at<kRec>(k, vi)] = ae[k];
#pragma unroll
      for (int k = 0; k < 16; ++k) pubw[s.at<kRec>(4 + k, vi)] = al[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) pubw[s.at<kRec>(20 + k, vi)] = mu[k];
      s.pub_epoch[1 - p][vi] = epoch;
    } else if (live) {
      // idle robot: its record is carried over to the other buffer unchanged
#pragma unroll
      for (int k = 0; k < kRec; ++k) pubw[s.at<kRec>(k, vi)] = pubr[s.at<kRec>(k, vi)];
      s.pub_epoch[1 - p][vi] = s.pub_epoch[p][vi];
      copy_dyn_messages(s, p, vi);
    }
  }
  if (live && i == 0) {
    s.iter_factor[r] = itf;
    if (do_int) s.latest[r] = 0;
    else if (do_ext) s.latest[r] = 1;
  }
  if (INT && requalify) {
    // every variable of the robot must qualify; idle robots keep their mode
    const unsigned all = ((V == 32) ? 0xffffffffu : ((1u << V) - 1u)) << (rl * V);
    const unsigned votes = __ballot_sync(0xffffffffu, live && do_int && light);
    // Two qualifying internal halves in a row: both buffers of the double-buffered records and messages
    // then hold this robot's decoupled state (k_iterate_axis never touches the cross rows).
    if (live && do_int && i == 0) s.mode[r] = (votes & all) == all ? (s.mode[r] == 2 ? 0 : 2) : 1;
  }
}

// par >= 0: the robots of Store::gen_list (length gen_count[par]; the counter of the next launch is
// cleared here) through a grid-stride loop; par < 0: every robot of the shard, no hand-back.
template <bool EXT, bool INT>
__global__ void GBP_ITER_BOUNDS
    k_iterate(const __grid_constant__ Store s, const int p, const uint32_t epoch, const int par) {
  extern __shared__ double iter_smem[];  // kIterSmemBytes: the threads' mirror-message strips (stream_mirrors)
  const int V = s.V;
  const int rpw = 32 / V;
  const unsigned lane = threadIdx.x & 31u;
  const int rl = int(lane) / V;
  const int i = int(lane) - rl * V;
  pdl_wait();
  pdl_launch_dependents();
  int64_t nrob = s.Nloc;
  if (par >= 0) {
    nrob = s.gen_count[par];
    if (blockIdx.x == 0 && threadIdx.x == 0) s.gen_count[par ^ 1] = 0;
  }
  const int64_t nwarps = (nrob + rpw - 1) / rpw;
  const int64_t wstride = (int64_t(gridDim.x) * kIterBlock) >> 5;
  for (int64_t warp = (int64_t(blockIdx.x) * kIterBlock + threadIdx.x) >> 5; warp < nwarps; warp += wstride) {
    const int64_t k = warp * rpw + rl;
    const bool live = rl < rpw && k < nrob;
    const int64_t r = live ? (par >= 0 ? int64_t(s.gen_list[k]) : k) : 0;
    iterate_warp<EXT, INT>(s, p, epoch, r, live, rl, i, lane, par >= 0, iter_smem + threadIdx.x);
  }
}

// ---- a whole iterate_gbp in ONE launch, for swarms that fit the GPU at once ----------------------------------------
// The reference's own scenarios have 10 - 50 robots: there a tick is 30 kernel launches of a few microseconds of work
// each, and launch latency is all there is.  k_tick_fused runs the launches of run_schedule back to back inside one
// cooperative grid — k_edge_messages' body, a grid barrier, k_iterate's body (every robot, as under
// gbp_world_set_iterate_path(general_only)), a grid barrier — with the same device functions, so the bits are those of
// the launch-per-half-step path.  Opt-in (gbp_world_set_iterate_path(w, 2)); the host falls back to separate launches
// when the swarm needs more warps than are resident at once.
struct TickPlan {
  int n;             // launches
  uint8_t ext[64];   // launch j runs an external half
  uint8_t in[64];    // ... and / or an internal half (both: external of sub-step t fused with internal of t + 1)
};
template <bool EXT, bool INT>
GBP_DEV void iterate_all_body(const Store &s, const int p, const uint32_t epoch, double *strip) {
  const int V = s.V;
  const int rpw = 32 / V;
  const unsigned lane = threadIdx.x & 31u;
  const int rl = int(lane) / V;
  const int i = int(lane) - rl * V;
  const int64_t nrob = s.Nloc;
  const int64_t nwarps = (nrob + rpw - 1) / rpw;
  const int64_t wstride = (int64_t(gridDim.x) * kIterBlock) >> 5;
  for (int64_t warp = (int64_t(blockIdx.x) * kIterBlock + threadIdx.x) >> 5; warp < nwarps; warp += wstride) {
    const int64_t k = warp * rpw + rl;
    const bool live = rl < rpw && k < nrob;
    iterate_warp<EXT, INT>(s, p, epoch, live ? k : 0, live, rl, i, lane, false, strip);
  }
}
__global__ void GBP_ITER_BOUNDS k_tick_fused(const __grid_constant__ Store s, int p, uint32_t epoch, const int chunks,
                                             const __grid_constant__ TickPlan plan) {
  extern __shared__ double iter_smem[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  double *const strip = iter_smem + threadIdx.x;
  for (int j = 0; j < plan.n; ++j) {
    epoch += 1;  // group_launch steps the epoch once per launch
    const bool ext = plan.ext[j] != 0, in = plan.in[j] != 0;
    if (ext && s.E > 0) {
      edge_messages_body(s, p, -1, chunks);
      grid.sync();
    }
    if (ext && in) iterate_all_body<true, true>(s, p, epoch, strip);
    else if (ext) iterate_all_body<true, false>(s, p, epoch, strip);
    else iterate_all_body<false, true>(s, p, epoch, strip);
    if (in) p ^= 1;
    grid.sync();
  }
}

}  // namespace gbp
