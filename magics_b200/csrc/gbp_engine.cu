// gbp_engine.cu — world object, launch sequencing and the C ABI (include/gbp_b200.h).
//
// Host-side counterpart of the reference's RobotPlugin FixedUpdate chain
// (planner/robot.rs:85-108) and of the FactorGraph method set it calls; all
// state lives in the device store (gbp_store.cuh).  No CPU fallback exists:
// every entry point that computes launches CUDA kernels on the world's stream.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "../../include/gbp_b200.h"
#include "gbp_collide.cuh"
#include "gbp_comm.cuh"
#include "gbp_iterate.cuh"
#include "gbp_iterate_axis.cuh"
#include "gbp_math.cuh"
#include "gbp_shard.cuh"
#include "gbp_store.cuh"
#include "gbp_sdf.cuh"
#include "gbp_topology.cuh"

using gbp::Store;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(GBP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

template <class T>
cudaError_t dalloc(T *&p, size_t n) {
  p = nullptr;
  if (n == 0) n = 1;
  return cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T));
}

// Re-stride a [planes][old_stride] array to [planes][new_stride], keeping the
// first `used` entries of each plane.  `tiled`: a per-variable array in the tiled layout of
// gbp_store.cuh (tiles of 32 slots x planes, back to back) — growing it appends tiles, the used
// ones move as one block.
template <class T>
cudaError_t regrow(T *&p, int planes, int64_t old_stride, int64_t new_stride, int64_t used,
                   cudaStream_t st, bool tiled = false) {
  T *q = nullptr;
  cudaError_t e = dalloc(q, size_t(planes) * size_t(new_stride));
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(q, 0, size_t(planes) * size_t(new_stride) * sizeof(T), st);
  if (e != cudaSuccess) return e;
  if (p && used > 0) {
    if (tiled && GBP_TILED && planes > 1) {
      const size_t tiles = size_t((used + gbp::kTile - 1) / gbp::kTile);
      e = cudaMemcpyAsync(q, p, tiles * size_t(planes) * gbp::kTile * sizeof(T), cudaMemcpyDeviceToDevice, st);
    } else {
      e = cudaMemcpy2DAsync(q, size_t(new_stride) * sizeof(T), p, size_t(old_stride) * sizeof(T),
                            size_t(used) * sizeof(T), size_t(planes), cudaMemcpyDeviceToDevice, st);
    }
    if (e != cudaSuccess) return e;
  }
  if (p) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    cudaFree(p);
  }
  p = q;
  return cudaSuccess;
}

// Upload host [planes][n] into rows of a device [planes][stride] array at column `at`.
template <class T>
cudaError_t upload_planes(T *dst, int64_t stride, int64_t at, const T *src, int planes, int64_t n,
                          cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  return cudaMemcpy2DAsync(dst + at, size_t(stride) * sizeof(T), src, size_t(n) * sizeof(T),
                           size_t(n) * sizeof(T), size_t(planes), cudaMemcpyHostToDevice, st);
}

inline unsigned blocks_for(int64_t n, int t) { return unsigned((n + t - 1) / t); }

// ---- gbp_schedule crate (host side, tiny) ------------------------------------
// interleave_evenly.rs:40-110
void ie_fill(uint8_t *s, int len, int n) {
  const int max = len, half = max / 2;
  auto cycle = [&](int times) {
    for (int i = 0; i < len; ++i) s[i] = (i % times) == 0 ? 1 : 0;
  };
  if (n == max) {
    std::fill(s, s + len, uint8_t(1));
  } else if (n == 0) {
    std::fill(s, s + len, uint8_t(0));
  } else if (n % 2 == 1 && max % 2 == 1) {
    if (max % n == 0) {
      cycle(max / n);
    } else {
      ie_fill(s, half, n / 2);
      s[half] = 1;
      ie_fill(s + half + 1, len - half - 1, n / 2);
      std::reverse(s + half + 1, s + len);
    }
  } else if (n % 2 == 0 && max % 2 == 1) {
    ie_fill(s, half, n / 2);
    std::reverse(s, s + half);
    s[half] = 0;
    ie_fill(s + half + 1, len - half - 1, n / 2);
  } else if (n % 2 == 0 && max % 2 == 0) {
    if (max % n == 0) {
      cycle(max / n);
    } else {
      ie_fill(s, half, n / 2);
      ie_fill(s + half, len - half, n / 2);
    }
  } else {
    ie_fill(s, half, n / 2 + 1);
    std::reverse(s, s + half);
    ie_fill(s + half, len - half, n / 2);
  }
}
void schedule_half(int kind, int n, int max, uint8_t *out) {
  switch (kind) {
    case GBP_SCHEDULE_CENTERED: {  // centered.rs:12-49
      for (int idx = 0; idx < max; ++idx) {
        if (n == 0 && max == 1) {
          out[idx] = 0;
          continue;
        }
        const int mid = max / 2, hn = n / 2;
        const int start = mid >= hn ? mid - hn : 0;
        const int end = (start + n <= max) ? start + n - 1 : max - 1;
        out[idx] = (idx >= start && idx <= end) ? 1 : 0;
      }
      break;
    }
    case GBP_SCHEDULE_INTERLEAVE_EVENLY: ie_fill(out, max, n); break;
    case GBP_SCHEDULE_SOON_AS_POSSIBLE:  // soon_as_possible.rs:26-49
      for (int i = 0; i < max; ++i) out[i] = i < n ? 1 : 0;
      break;
    case GBP_SCHEDULE_LATE_AS_POSSIBLE:  // late_as_possible.rs:29-50
      for (int i = 0; i < max; ++i) out[i] = (n == max) ? 1 : (n == 0 ? 0 : (i >= max - n ? 1 : 0));
      break;
    case GBP_SCHEDULE_HALF_BEGINNING_HALF_END: {  // half_beginning_half_end.rs:19-45
      const int hn = n / 2, rem = n % 2, sm = hn, em = max - hn - rem;
      for (int i = 0; i < max; ++i) out[i] = (i < sm || i >= em) ? 1 : 0;
      break;
    }
  }
}

// ---- small kernels -------------------------------------------------------------

// VariableNode::new (variable.rs:140-166) for the variables [first, first+count):
// expects prior_lam already uploaded; `mu0` = the initial means and delta_t as [5][count] planes.
__global__ void k_init_vars(Store s, int64_t first, int64_t count, const double *__restrict__ mu0) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int64_t vi = first + t;
  double mu[4], lam[16], cov[16];
  const double pl = s.prior_lam[vi];
#pragma unroll
  for (int k = 0; k < 4; ++k) mu[k] = mu0[k * count + t];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lam[k] = (k % 5 == 0) ? pl : 0.0;
    cov[k] = 0.0;
  }
  gbp::inv4(lam, cov);  // unwrap_or_else(zeros)
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 16; ++k) fin = fin && isfinite(cov[k]);
  for (int b = 0; b < 2; ++b) {
    double *rec = s.pub[b];
#pragma unroll
    for (int k = 0; k < 4; ++k) rec[s.at<gbp::kRec>(k, vi)] = pl * mu[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) rec[s.at<gbp::kRec>(4 + k, vi)] = lam[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) rec[s.at<gbp::kRec>(20 + k, vi)] = mu[k];
    s.pub_epoch[b][vi] = 0u;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s.prior_eta[s.at<4>(k, vi)] = pl * mu[k];
    s.bel_ext[s.at<gbp::kRec>(k, vi)] = pl * mu[k];
    s.bel_ext[s.at<gbp::kRec>(20 + k, vi)] = mu[k];
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    s.bel_ext[s.at<gbp::kRec>(4 + k, vi)] = lam[k];
    s.cov[s.at<16>(k, vi)] = cov[k];
  }
  s.valid[vi] = fin ? 1 : 0;
  s.cov_lazy[vi] = 0;
  s.mu_ext[s.at<2>(0, vi)] = mu[0];
  s.mu_ext[s.at<2>(1, vi)] = mu[1];
  for (int b = 0; b < 2; ++b) {
    s.m_dynL[b][s.at<20>(0, vi)] = gbp::empty_marker();
    s.m_dynR[b][s.at<20>(0, vi)] = gbp::empty_marker();
  }
  s.m_obs[s.at<4>(0, vi)] = gbp::empty_marker();
  s.m_trk[s.at<3>(0, vi)] = gbp::empty_marker();
  s.trk_record[vi] = 0u;
  s.trk_timeout[vi] = -1;
  s.trk_seed[vi] = 1;
  s.trk_last[s.at<2>(0, vi)] = float(mu[0]);  // with_last_measurement (factor/mod.rs:279-283)
  s.trk_last[s.at<2>(1, vi)] = float(mu[1]);
  s.trk_value[vi] = 0.0;
  // DynamicFactor i between variables i and i+1 (robot.rs:1228-1255); the last variable has none
  const double dt = mu0[4 * count + t];
  double q11, q12, q22;
  gbp::dyn_q(dt, s.qs_dyn, q11, q12, q22);
  s.dyn_c[s.at<4>(0, vi)] = dt;
  s.dyn_c[s.at<4>(1, vi)] = q11;
  s.dyn_c[s.at<4>(2, vi)] = q12;
  s.dyn_c[s.at<4>(3, vi)] = q22;
}

// FactorGraph::reset_variables on another shard's robot (global ids `ids`, ascending): the own robots' InterRobot
// factors toward it have lost its message (they will linearise at zeros), i.e. every own edge whose neighbour is in
// the list is frozen at (0, 0) — the remote half of k_reset_reverse_edges.
__global__ void k_freeze_edges_toward(Store s, const int32_t *__restrict__ egid, int n_ids,
                                      const int32_t *__restrict__ ids) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= s.E) return;
  const int32_t a = egid[e];
  int lo = 0, hi = n_ids;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ids[mid] < a) lo = mid + 1;
    else hi = mid;
  }
  if (lo >= n_ids || ids[lo] != a) return;
  const int Vm1 = s.V - 1;
  s.e_frozen[e] = uint8_t(s.e_frozen[e] | 1);
  for (int i = 0; i < Vm1; ++i) {
    s.mu_frozen[e * Vm1 + i] = 0.0;
    s.mu_frozen[s.EV + e * Vm1 + i] = 0.0;
  }
}

// TrackingFactor state read-back (tracking.rs:62-90 `Tracking.record`, `LastMeasurement`): AoS rows per variable.
__global__ void k_gather_tracking(Store s, int64_t nv, int64_t *record, float *pos, double *value) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nv) return;
  record[t] = int64_t(s.trk_record[t]);
  pos[2 * t] = s.trk_last[s.at<2>(0, t)];
  pos[2 * t + 1] = s.trk_last[s.at<2>(1, t)];
  value[t] = s.trk_value[t];
}

// Store::dyn_tab: the Dynamic-factor constants of one robot, by variable index (valid for every robot while
// all of them share one radius, i.e. one delta_t per factor).
__global__ void k_dyn_table(Store s, int64_t robot, double *tab) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i >= s.V) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) tab[4 * i + k] = s.dyn_c[s.at<4>(k, robot * s.V + i)];
}

// VariableNode::change_prior + FactorGraph::change_prior_of_variable for variable
// `var` of robot r with new mean nm (variable.rs:203-230, factorgraph.rs:494-528):
// every factor that holds a message from this variable now holds
// (eta_belief, Lambda_belief, new mean); the variable's inbox is emptied.
__device__ void change_prior_dev(const Store &s, int p, uint32_t epoch, int64_t r, int var,
                                 const double (&nm)[4]) {
  const int64_t vi = r * s.V + var;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const double pl = s.prior_lam[vi];
  double rec[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) rec[k] = src[s.at<gbp::kRec>(k, vi)];
#pragma unroll
  for (int k = 0; k < 4; ++k) s.prior_eta[s.at<4>(k, vi)] = pl * nm[k];
  double *dst[2] = {s.pub[p], s.bel_ext};
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int k = 0; k < 20; ++k) dst[b][s.at<gbp::kRec>(k, vi)] = rec[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[b][s.at<gbp::kRec>(20 + k, vi)] = nm[k];
  }
  s.pub_epoch[p][vi] = epoch;
  s.mu_ext[s.at<2>(0, vi)] = nm[0];
  s.mu_ext[s.at<2>(1, vi)] = nm[1];
  s.m_dynL[p][s.at<20>(0, vi)] = gbp::empty_marker();
  s.m_dynR[p][s.at<20>(0, vi)] = gbp::empty_marker();
  s.m_obs[s.at<4>(0, vi)] = gbp::empty_marker();
  s.m_trk[s.at<3>(0, vi)] = gbp::empty_marker();
  if (var >= 1 && s.eoff)
    for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
      const int64_t m = e * (s.V - 1) + (var - 1);
      s.mir[m] = gbp::empty_marker();
      // external factors receive the new mean whatever the antenna state (robot.rs:2272-2282)
      s.mu_frozen[m] = nm[0];
      s.mu_frozen[s.EV + m] = nm[1];
    }
}

// change_prior_dev spread over the 32 lanes of a warp (one robot per warp): a prior change touches
// ~70 scattered sectors of one robot, so one thread per robot is pure latency; every lane must
// hold the same `nm`, computed from reads that happened before the __syncwarp between the two phases.
// Two phases — everything a change reads (prior_load), then its stores (prior_store) — so that a warp making
// both prior changes of a tick (k_prior_both) has the loads of both in flight together.
struct PriorLoad {
  bool latest;
  double keep, pl, mark, mir0;
  int64_t eo0, eo1;
  uint8_t frz0;
};
__device__ PriorLoad prior_load(const Store &s, int p, int64_t r, int var, unsigned lane) {
  const int64_t vi = r * s.V + var;
  PriorLoad q;
  q.latest = s.latest[r] != 0;
  q.keep = 0.0;
  if (q.latest && lane < 20) q.keep = s.bel_ext[s.at<gbp::kRec>(lane, vi)];
  q.pl = s.prior_lam[vi];
  q.mark = 0.0;
  if (lane == 25) q.mark = s.m_dynL[p][s.at<20>(0, vi)];
  else if (lane == 26) q.mark = s.m_dynR[p][s.at<20>(0, vi)];
  else if (lane == 27) q.mark = s.m_obs[s.at<4>(0, vi)];
  else if (lane == 28) q.mark = s.m_trk[s.at<3>(0, vi)];
  q.eo0 = q.eo1 = 0;
  q.mir0 = 0.0;
  q.frz0 = 0;
  if (var >= 1 && s.eoff) {
    q.eo0 = s.eoff[r];
    q.eo1 = s.eoff[r + 1];
    const int64_t e = q.eo0 + lane;
    if (e < q.eo1) {
      q.mir0 = s.mir[e * (s.V - 1) + (var - 1)];
      q.frz0 = s.e_frozen[e];
    }
  }
  return q;
}
__device__ void prior_store(const Store &s, int p, uint32_t epoch, int64_t r, int var, const double (&nm)[4],
                            unsigned lane, const PriorLoad &q) {
  // Every store below lands in a 32-byte sector of which it fills 8 bytes (a read-modify-write in DRAM), so
  // nothing is written that already holds the value: the (eta, Lambda) rows move only when the current belief
  // lives in bel_ext, bel_ext is left alone while nothing reads it (latest == 0: every reader takes
  // `latest ? bel_ext : pub[p]`), an Empty marker is not stored over an Empty marker, and mu_frozen is
  // only meaningful while the edge's frozen bit is set (otherwise A's factor holds mu_ext).
  const int64_t vi = r * s.V + var;
  if (lane < 20) {
    if (q.latest) s.pub[p][s.at<gbp::kRec>(lane, vi)] = q.keep;
  } else if (lane < 24) {
    const int k = int(lane) - 20;
    s.pub[p][s.at<gbp::kRec>(20 + k, vi)] = nm[k];
    if (q.latest) s.bel_ext[s.at<gbp::kRec>(20 + k, vi)] = nm[k];
    s.prior_eta[s.at<4>(k, vi)] = q.pl * nm[k];
  } else if (lane == 24) {
    s.pub_epoch[p][vi] = epoch;
    s.mu_ext[s.at<2>(0, vi)] = nm[0];
    s.mu_ext[s.at<2>(1, vi)] = nm[1];
  } else if (lane <= 28 && !gbp::is_empty_marker(q.mark)) {
    if (lane == 25) s.m_dynL[p][s.at<20>(0, vi)] = gbp::empty_marker();
    else if (lane == 26) s.m_dynR[p][s.at<20>(0, vi)] = gbp::empty_marker();
    else if (lane == 27) s.m_obs[s.at<4>(0, vi)] = gbp::empty_marker();
    else s.m_trk[s.at<3>(0, vi)] = gbp::empty_marker();
  }
  if (var >= 1 && s.eoff)
    for (int64_t e = q.eo0 + lane; e < q.eo1; e += 32) {
      const int64_t m = e * (s.V - 1) + (var - 1);
      const bool first = e == q.eo0 + lane;
      const double mir = first ? q.mir0 : s.mir[m];
      const uint8_t frz = first ? q.frz0 : s.e_frozen[e];
      if (!gbp::is_empty_marker(mir)) s.mir[m] = gbp::empty_marker();
      // external factors receive the new mean whatever the antenna state (robot.rs:2272-2282)
      if (frz & 1) {
        s.mu_frozen[m] = nm[0];
        s.mu_frozen[s.EV + m] = nm[1];
      }
    }
}
__device__ void change_prior_warp(const Store &s, int p, uint32_t epoch, int64_t r, int var, const double (&nm)[4],
                                  unsigned lane) {
  const PriorLoad q = prior_load(s, p, r, var, lane);
  __syncwarp();
  prior_store(s, p, epoch, r, var, nm, lane, q);
}

// update_prior_of_horizon_state (planner/robot.rs:2182-2283), one warp per robot.
__global__ void k_prior_horizon(Store s, int p, uint32_t epoch, double delta_t, double max_speed,
                                int iterations_internal) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (r >= s.Nloc) return;
  if (s.finished[r] || s.idle[r]) return;
  const int32_t nwp = s.wp_off[r + 1] - s.wp_off[r], k = s.next_wp[r];
  if (k < 0 || k >= nwp) {
    if (lane == 0) s.finished[r] = 1;
    return;
  }
  if (iterations_internal == 0) return;
  const int64_t vi = r * s.V + (s.V - 1);
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const double ex = src[s.at<gbp::kRec>(20, vi)], ey = src[s.at<gbp::kRec>(21, vi)];
  const float *wp = s.wp_xy + 2 * (size_t(s.wp_off[r]) + k);
  const double hx = double(wp[0]) - ex, hy = double(wp[1]) - ey;
  const double dist = gbp::norm2(hx, hy);
  double nx = hx, ny = hy;
  if (!(dist == 0.0 || isinf(dist))) {
    nx /= dist;
    ny /= dist;
  }
  const double sp = fmin(max_speed, dist);
  const double vx = sp * nx, vy = sp * ny;
  const double nm[4] = {ex + vx * delta_t, ey + vy * delta_t, vx, vy};
  change_prior_warp(s, p, epoch, r, s.V - 1, nm, lane);
}

// update_prior_of_current_state_v3 (planner/robot.rs:2286-2338), one warp per robot.
__global__ void k_prior_current(Store s, int p, uint32_t epoch, float delta_t) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (r >= s.Nloc) return;
  if (s.idle[r]) return;
  const int64_t v0 = r * s.V, v1 = v0 + 1;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const float time_scale = __fdiv_rn(delta_t, s.t0[r]);
  double ch[4], nm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double c = src[s.at<gbp::kRec>(20 + k, v0)];
    ch[k] = double(time_scale) * (src[s.at<gbp::kRec>(20 + k, v1)] - c);
    nm[k] = c + ch[k];
  }
  const float px = s.pos[r], pz = s.pos[s.cap + r];
  change_prior_warp(s, p, epoch, r, 0, nm, lane);
  if (lane == 0) {
    s.pos[r] = __fadd_rn(px, float(ch[0]));
    s.pos[s.cap + r] = __fadd_rn(pz, float(ch[1]));
  }
}

// update_prior_of_horizon_state followed by update_prior_of_current_state (the order of a tick, robot.rs:85-108) in
// one launch, one warp per robot: the two changes touch different variables (V - 1 and 0; the second reads the means
// of 0 and 1), so for V >= 3 the result is that of k_prior_horizon then k_prior_current, with the scattered loads of
// both in flight at once (each kernel alone is bound by its chain of dependent loads, not by bytes).
__global__ void k_prior_both(Store s, int p, uint32_t epoch_h, uint32_t epoch_c, double delta_t, double max_speed,
                             int iterations_internal, float delta_t_f) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (r >= s.Nloc) return;
  if (s.idle[r]) return;
  const bool finished = s.finished[r] != 0;
  const int32_t w0 = s.wp_off[r], nwp = s.wp_off[r + 1] - w0, k = s.next_wp[r];
  const float t0 = s.t0[r], px = s.pos[r], pz = s.pos[s.cap + r];
  bool do_h = !finished;
  if (do_h && (k < 0 || k >= nwp)) {
    if (lane == 0) s.finished[r] = 1;
    do_h = false;
  }
  if (iterations_internal == 0) do_h = false;
  const int64_t v0 = r * s.V, v1 = v0 + 1, vh = v0 + (s.V - 1);
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  double ex = 0.0, ey = 0.0, wx = 0.0, wy = 0.0, c0[4], c1[4];
  if (do_h) {
    ex = src[s.at<gbp::kRec>(20, vh)];
    ey = src[s.at<gbp::kRec>(21, vh)];
    const float *wp = s.wp_xy + 2 * (size_t(w0) + k);
    wx = double(wp[0]);
    wy = double(wp[1]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    c0[q] = src[s.at<gbp::kRec>(20 + q, v0)];
    c1[q] = src[s.at<gbp::kRec>(20 + q, v1)];
  }
  PriorLoad qh{};
  if (do_h) qh = prior_load(s, p, r, s.V - 1, lane);
  const PriorLoad qc = prior_load(s, p, r, 0, lane);
  // horizon (k_prior_horizon)
  const double hx = wx - ex, hy = wy - ey;
  const double dist = gbp::norm2(hx, hy);
  double nx = hx, ny = hy;
  if (!(dist == 0.0 || isinf(dist))) {
    nx /= dist;
    ny /= dist;
  }
  const double sp = fmin(max_speed, dist);
  const double vx = sp * nx, vy = sp * ny;
  const double nmh[4] = {ex + vx * delta_t, ey + vy * delta_t, vx, vy};
  // current state (k_prior_current)
  const float time_scale = __fdiv_rn(delta_t_f, t0);
  double ch[4], nmc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    ch[q] = double(time_scale) * (c1[q] - c0[q]);
    nmc[q] = c0[q] + ch[q];
  }
  __syncwarp();
  if (do_h) prior_store(s, p, epoch_h, r, s.V - 1, nmh, lane, qh);
  prior_store(s, p, epoch_c, r, 0, nmc, lane, qc);
  if (lane == 0) {
    s.pos[r] = __fadd_rn(px, float(ch[0]));
    s.pos[s.cap + r] = __fadd_rn(pz, float(ch[1]));
  }
}

__global__ void k_change_prior_list(Store s, int p, uint32_t epoch, int var, int m,
                                    const int32_t *robots, const double *means) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const double nm[4] = {means[4 * t], means[4 * t + 1], means[4 * t + 2], means[4 * t + 3]};
  change_prior_dev(s, p, epoch, robots[t], var, nm);
}

// FactorGraph::reset_variables (factorgraph.rs:1541-1564) for the listed robots, one thread per
// (robot, variable): VariableNode::reset (variable.rs:350-360) — belief mean and precision replaced, the
// information vector kept, every inbox entry emptied — and FactorNode::empty_inbox (factor/mod.rs:480-483)
// for the own factors, whose messages FROM this variable are the published record: epoch 0 = "none".
__global__ void k_reset_variables(Store s, int p, int m, const int32_t *__restrict__ robots,
                                  const double *__restrict__ means, double first_last_sigma, double inbetween_sigma) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int V = s.V;
  if (t >= int64_t(m) * V) return;
  const int k = int(t / V), i = int(t - int64_t(k) * V);
  const int64_t r = robots[k], vi = r * V + i;
  const double sigma = (i == 0 || i == V - 1) ? first_last_sigma : inbetween_sigma;
  // VariableNode::reset leaves the covariance alone: resolve a lazy one before the precision it derives from changes
  if (s.cov_lazy[vi]) gbp::materialise_cov(s, p, r, vi);
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  double eta[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) eta[a] = src[s.at<gbp::kRec>(a, vi)];
  double *dst[2] = {s.pub[p], s.bel_ext};
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int a = 0; a < 4; ++a) dst[b][s.at<gbp::kRec>(a, vi)] = eta[a];
#pragma unroll
    for (int a = 0; a < 16; ++a) dst[b][s.at<gbp::kRec>(4 + a, vi)] = (a % 5 == 0) ? sigma : 0.0;  // from_diag_elem
#pragma unroll
    for (int a = 0; a < 4; ++a) dst[b][s.at<gbp::kRec>(20 + a, vi)] = means[4 * t + a];
  }
  s.pub_epoch[p][vi] = 0u;
  s.m_dynL[p][s.at<20>(0, vi)] = gbp::empty_marker();
  s.m_dynR[p][s.at<20>(0, vi)] = gbp::empty_marker();
  s.m_obs[s.at<4>(0, vi)] = gbp::empty_marker();
  s.m_trk[s.at<3>(0, vi)] = gbp::empty_marker();
  s.trk_seed[vi] = 0;
  if (i >= 1 && s.eoff)
    for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) s.mir[e * (V - 1) + (i - 1)] = gbp::empty_marker();
}

// ... and the robot's own InterRobot factors lose the message they held from the NEIGHBOUR's variable.
// The neighbour A evaluates that factor itself (gbp_iterate.cuh) with the mean it last sent: until A
// delivers again the factor linearises A's side at zeros, i.e. the edge (A <- r) is frozen at (0, 0).
// One thread per edge of a reset robot; A's list is sorted by id.
__global__ void k_reset_reverse_edges(Store s, int m, const int32_t *__restrict__ robots) {
  const int k = blockIdx.x;
  if (k >= m) return;
  const int32_t r = robots[k];
  const int Vm1 = s.V - 1;
  for (int64_t e = s.eoff[r] + threadIdx.x; e < s.eoff[r + 1]; e += blockDim.x) {
    const int32_t A = s.enbr[e];
    if (A >= s.Nloc) continue;
    // rows are ordered by GLOBAL id (on a shard the ghost slots of lower-id robots come first)
    const int32_t gr = s.gid[r];
    int64_t lo = s.eoff[A], hi = s.eoff[A + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (s.gid[s.enbr[mid]] < gr) lo = mid + 1;
      else hi = mid;
    }
    // every factor set toward r (strict_reference_quirks can leave more than one)
    for (; lo < s.eoff[A + 1] && s.enbr[lo] == r; ++lo) {
      s.e_frozen[lo] = uint8_t(s.e_frozen[lo] | 1);
      for (int i = 0; i < Vm1; ++i) {
        s.mu_frozen[lo * Vm1 + i] = 0.0;
        s.mu_frozen[s.EV + lo * Vm1 + i] = 0.0;
      }
    }
  }
}

// FactorGraph::reset_tracking_factors (factorgraph.rs:1566-1590): set_timeout(10) on the tracking factor
// of every variable but the first and the last.
__global__ void k_reset_tracking(Store s, int m, const int32_t *__restrict__ robots) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int V = s.V;
  if (t >= int64_t(m) * V) return;
  const int k = int(t / V), i = int(t - int64_t(k) * V);
  if (i >= 1 && i <= V - 2) s.trk_timeout[int64_t(robots[k]) * V + i] = 10;
}

__global__ void k_set_next_wp(Store s, int m, const int32_t *__restrict__ robots, int32_t value) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) s.next_wp[robots[t]] = value;
}

// Gather VariableBelief of every variable into the ABI's array-of-structs layout.
__global__ void k_gather_beliefs(Store s, int p, double *eta, double *lam, double *mean, double *cov,
                                 uint8_t *valid) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= int64_t(s.Nloc) * s.V) return;
  const int64_t r = t / s.V;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  if (eta)
    for (int k = 0; k < 4; ++k) eta[4 * t + k] = src[s.at<gbp::kRec>(k, t)];
  if (lam)
    for (int k = 0; k < 16; ++k) lam[16 * t + k] = src[s.at<gbp::kRec>(4 + k, t)];
  if (mean)
    for (int k = 0; k < 4; ++k) mean[4 * t + k] = src[s.at<gbp::kRec>(20 + k, t)];
  if (s.cov_lazy[t]) {
    // not stored: the covariance is inv4 of the current precision, taken and finite (gbp_iterate_axis.cuh)
    if (cov) {
      double l[16], c[16];
      for (int k = 0; k < 16; ++k) {
        l[k] = src[s.at<gbp::kRec>(4 + k, t)];
        c[k] = 0.0;
      }
      gbp::inv4(l, c);
      for (int k = 0; k < 16; ++k) cov[16 * t + k] = c[k];
    }
    if (valid) valid[t] = 1;
    return;
  }
  if (cov)
    for (int k = 0; k < 16; ++k) cov[16 * t + k] = s.cov[s.at<16>(k, t)];
  if (valid) valid[t] = s.valid[t];
}

__global__ void k_sdf_lookup(Store s, int m, const double *xy, uint32_t *px, uint32_t *py, double *val) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  val[t] = gbp::sdf_measure(s, xy[2 * t], xy[2 * t + 1], px + t, py + t);
}

// Store::ell_*: the first kEll edge heads of every own robot at fixed positions (gbp_store.cuh).
__global__ void k_ell_fill(Store s) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= int64_t(s.Nloc) * gbp::kEll) return;
  const int64_t r = t / gbp::kEll, e = s.eoff[r] + (t - r * gbp::kEll);
  const bool ok = e < s.eoff[r + 1];
  s.ell_nbr[t] = ok ? s.enbr[e] : -1;
  if (ok) {
    s.ell_birth[t] = s.e_birth[e];
    s.ell_dsafe[t] = s.e_dsafe[e];
  }
}

__global__ void k_set_dsafe(Store s, double mult) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) s.e_dsafe[e] = mult * double(s.radius[s.enbr[e]]);
}

// reached_waypoint (planner/robot.rs:2080-2176), one thread per robot; f32 arithmetic without
// contraction, as glam's Vec2::distance_squared evaluates it.
__global__ void k_reached_waypoint(Store s, int p, gbp_reached_when_t task, gbp_reached_when_t fin, uint8_t *out) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  if (out) out[r] = 0;
  if (s.gone[r] != 0.0f) return;  // despawned: the system's query no longer yields the entity
  const int32_t nwp = s.wp_off[r + 1] - s.wp_off[r], k = s.next_wp[r];
  if (k < 0 || k >= nwp) return;  // mission.next_waypoint() is None
  const gbp_reached_when_t c = (k == nwp - 1) ? fin : task;
  const int V = s.V;
  const int var = c.intersects_with == GBP_INTERSECTS_CURRENT
                      ? 0
                      : (c.intersects_with == GBP_INTERSECTS_HORIZON
                             ? V - 1
                             : (c.variable_index >= 0 && c.variable_index < V ? c.variable_index : V - 1));
  const int64_t vi = r * V + var;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const float ex = float(src[s.at<gbp::kRec>(20, vi)]), ey = float(src[s.at<gbp::kRec>(21, vi)]);
  const float rad = s.radius[r];
  const float dsq = c.distance == GBP_DISTANCE_ROBOT_RADIUS ? __fmul_rn(rad, rad) : __fmul_rn(c.meter, c.meter);
  const float *wp = s.wp_xy + 2 * (size_t(s.wp_off[r]) + k);
  const float dx = __fsub_rn(ex, wp[0]), dy = __fsub_rn(ey, wp[1]);
  const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  if (d2 < dsq) {
    s.next_wp[r] = k + 1;
    if (out) out[r] = 1;
  }
}

// A few result words from device memory straight into page-locked host memory (device-accessible
// under unified addressing).  The per-tick size read-back of the topology pass used to be a small
// cudaMemcpyAsync D2H; DMA transfers of one direction are served in order, so it queued behind the
// 320 MB read-back of the previous tick's means and the host — and with it the next launches — waited
// 5 ms for it (scripts/e2e_probe.py).  Stores from a kernel do not pass through that queue.
__global__ void k_words_to_host(int64_t *__restrict__ host_dst, const int64_t *__restrict__ src, int n) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) host_dst[k] = src[k];
  __threadfence_system();
}

// ---- MessageCount accounting (factorgraph/mod.rs:103-137; read by export.rs:434-439) ---------------------------
// The reference counts a message as SENT per inbox key of the node that updates (factor/mod.rs:353-367, 410-452;
// variable.rs:299-332; NOT in change_prior, variable.rs:208-229) and as RECEIVED in receive_message_from
// (variable.rs:176-190; factor/mod.rs:307-318, after the `enabled` test), internal / external by graph id.  These
// kernels reproduce the counts from who takes part in a half-step — no message is looked at.  One thread per robot,
// launched BEFORE the work they account for (they read iteration counts and mission flags as that work finds them).
// cnt[k * cap + r], k = sent internal, sent external, received internal, received external: the robot's variables
// and non-InterRobot factors; emsg[k * ecap + e]: its own InterRobot factors toward neighbour e (see EdgeSet::e_msg).
struct MsgShape {
  int V;
  unsigned long long keys_own;  // same-graph non-InterRobot inbox keys of a robot's variables: 2(V-1) + 2(V-2)
  __host__ __device__ explicit MsgShape(int v) : V(v), keys_own(2ull * (v - 1) + 2ull * (v - 2)) {}
  // edges variable <-> enabled factor of each kind (factor_receive is a no-op for a disabled factor)
  __device__ unsigned long long enabled_edges(const Store &s, bool tracking_runs) const {
    return (s.en_dyn ? 2ull * (V - 1) : 0ull) + (s.en_obs ? 1ull * (V - 2) : 0ull) + (tracking_runs ? 1ull * (V - 2) : 0ull);
  }
};
__global__ void k_msg_init_robots(Store s, int64_t first, int64_t count, int64_t cap, unsigned long long *cnt) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int64_t r = first + t;
  const MsgShape sh(s.V);
  // add_internal_edge (factorgraph.rs:304-330) for every own factor: the variable receives Empty, the factor - if
  // enabled - the variable's message
  cnt[0 * cap + r] = 0;
  cnt[1 * cap + r] = 0;
  cnt[2 * cap + r] = sh.keys_own + sh.enabled_edges(s, s.en_trk != 0);
  cnt[3 * cap + r] = 0;
}
template <bool EXT, bool INT>
__global__ void k_msg_half(Store s, int64_t cap, unsigned long long *cnt, int64_t ecap, uint32_t *emsg) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  const bool idle = s.idle[r] != 0, ant = s.antenna[r] != 0;
  const bool do_ext = EXT && !idle && ant, do_int = INT && !idle;
  if (!do_ext && !do_int) return;
  const MsgShape sh(s.V);
  const unsigned long long Vm1 = sh.V - 1;
  const int64_t e0 = s.eoff ? s.eoff[r] : 0, e1 = s.eoff ? s.eoff[r + 1] : 0;
  const unsigned long long K = (unsigned long long)(e1 - e0);
  unsigned long long si = 0, se = 0, ri = 0, re = 0;
  uint32_t itf = s.iter_factor[r];
  if (do_ext) {
    // external_factor_iteration (factorgraph.rs:719-760): every enabled own InterRobot factor sends one internal and
    // one external message; the neighbours' factors do the same and theirs reach my variables if they ran
    // (robot.rs:1814-1831).  external_variable_iteration (factorgraph.rs:794-826): one response per inbox key; the
    // neighbours' responses reach my InterRobot factors if the neighbour ran (robot.rs:1843-1858).
    si += sh.keys_own + K * Vm1;
    se += K * Vm1;
    for (int64_t e = e0; e < e1; ++e) {
      const int A = s.enbr[e];
      const bool a_runs = s.antenna[A] != 0 && s.idle[A] == 0;
      if (s.en_ir) {
        emsg[0 * ecap + e] += uint32_t(Vm1);
        if (a_runs) {
          re += Vm1;
          emsg[2 * ecap + e] += uint32_t(Vm1);
        }
      }
    }
    itf += 1;
  }
  if (do_int) {
    // internal_factor_iteration (factorgraph.rs:688-714): enabled non-InterRobot factors, Tracking from
    // iteration_count.factor >= 10 on; internal_variable_iteration (:762-790): responses to same-graph factors are
    // delivered (to the enabled ones), the others only counted as sent.
    const unsigned long long f = sh.enabled_edges(s, s.en_trk && itf >= 10u);
    si += f;
    ri += f;
    si += sh.keys_own + K * Vm1;
    se += K * Vm1;
    ri += sh.enabled_edges(s, s.en_trk != 0);
    if (s.en_ir)
      for (int64_t e = e0; e < e1; ++e) emsg[1 * ecap + e] += uint32_t(Vm1);
  }
  cnt[0 * cap + r] += si;
  cnt[1 * cap + r] += se;
  cnt[2 * cap + r] += ri;
  cnt[3 * cap + r] += re;
}
// Will update_prior_of_horizon_state change robot r's horizon prior in the call that follows? (robot.rs:2190-2215)
__device__ __forceinline__ bool horizon_update_runs(const Store &s, int64_t r, int iterations_internal) {
  if (s.finished[r] || s.idle[r]) return false;
  const int32_t nwp = s.wp_off[r + 1] - s.wp_off[r], k = s.next_wp[r];
  return k >= 0 && k < nwp && iterations_internal != 0;
}
// update_prior_of_horizon_state: change_prior_of_variable(V-1) delivers to dyn(V-2), to the own InterRobot factors of
// variable V-1 and to the neighbours' (factorgraph.rs:494-528, robot.rs:2266-2282: no antenna test); nothing is
// counted as sent.
__global__ void k_msg_prior_horizon(Store s, int iterations_internal, int64_t cap, unsigned long long *cnt, int64_t ecap,
                                    uint32_t *emsg) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  const bool mine = horizon_update_runs(s, r, iterations_internal);
  if (mine && s.en_dyn) cnt[2 * cap + r] += 1;
  if (!s.en_ir || !s.eoff) return;
  for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
    if (mine) emsg[1 * ecap + e] += 1u;
    const int A = s.enbr[e];
    if (A < s.Nloc && s.gone[r] == 0.0f && horizon_update_runs(s, A, iterations_internal)) emsg[2 * ecap + e] += 1u;
  }
}
__global__ void k_msg_prior_current(Store s, int64_t cap, unsigned long long *cnt) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  if (!s.idle[r] && s.en_dyn) cnt[2 * cap + r] += 1;  // dyn(0) receives (robot.rs:2286-2338)
}
// FactorGraph::change_prior_of_variable for variable `var` of the listed robots (gbp_world_change_prior_of_variable):
// the same-graph factors of the variable receive (dyn(var-1), dyn(var), obs, trk as they exist and are enabled; the own
// InterRobot factors for var >= 1), and so does every neighbour's InterRobot factor toward it — counted on the
// neighbour's edge, found by its global id.
__global__ void k_msg_change_prior(Store s, int var, int m, const int32_t *__restrict__ robots, int64_t cap,
                                   unsigned long long *cnt, int64_t ecap, uint32_t *emsg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int32_t r = robots[t];
  const int V = s.V;
  unsigned long long ri = 0;
  if (s.en_dyn) ri += (var >= 1) + (var <= V - 2);
  if (var >= 1 && var <= V - 2) ri += (s.en_obs ? 1 : 0) + (s.en_trk ? 1 : 0);
  atomicAdd(&cnt[2 * cap + r], ri);
  if (var < 1 || !s.en_ir || !s.eoff) return;
  const int32_t gr = s.gid[r];
  for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
    atomicAdd(&emsg[1 * ecap + e], 1u);
    const int32_t A = s.enbr[e];
    if (A >= s.Nloc || s.gone[A] != 0.0f) continue;
    int64_t lo = s.eoff[A], hi = s.eoff[A + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (s.gid[s.enbr[mid]] < gr) lo = mid + 1;
      else hi = mid;
    }
    // one delivery per factor set the neighbour holds toward r; with several sets (strict_reference_quirks) the
    // k-th set of r's row pairs with the k-th of the neighbour's: count on the matching position
    int64_t k = 0;
    for (int64_t q = s.eoff[r]; q < e; ++q) k += s.enbr[q] == A;
    if (lo + k < s.eoff[A + 1] && s.enbr[lo + k] == r) atomicAdd(&emsg[2 * ecap + lo + k], 1u);
  }
}
// Topology change: counters of surviving edges move with them; a new connection is a handshake — per factor the own
// variable and (if enabled) the factor receive internally (add_internal_edge), the neighbour's variable and (if
// enabled) the factor receive externally (add_external_edge factorgraph.rs:340-353, robot.rs:1557-1585).
__global__ void k_msg_edges(Store s, int32_t n, const int64_t *__restrict__ noff, const int64_t *__restrict__ map,
                            int64_t ocap, const uint32_t *__restrict__ omsg, int64_t ncap, uint32_t *nmsg, int64_t cap,
                            unsigned long long *cnt) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t Vm1 = uint32_t(s.V - 1);
  unsigned long long fresh = 0;
  for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
    const int64_t old = map[e];
    if (old >= 0) {
      for (int k = 0; k < 3; ++k) nmsg[k * ncap + e] = omsg[k * ocap + old];
    } else {
      nmsg[0 * ncap + e] = 0u;
      nmsg[1 * ncap + e] = s.en_ir ? Vm1 : 0u;
      nmsg[2 * ncap + e] = s.en_ir ? Vm1 : 0u;
      ++fresh;
    }
  }
  cnt[2 * cap + r] += fresh * Vm1;
  cnt[3 * cap + r] += fresh * Vm1;
}
// FactorGraph::messages_sent / messages_received (factorgraph.rs:876-890): out[4 r + k]; zeros for a despawned robot.
__global__ void k_msg_gather(Store s, int64_t cap, const unsigned long long *__restrict__ cnt, int64_t ecap,
                             const uint32_t *__restrict__ emsg, int64_t *out) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  unsigned long long v[4] = {0, 0, 0, 0};
  if (s.gone[r] == 0.0f) {
    for (int k = 0; k < 4; ++k) v[k] = cnt[k * cap + r];
    if (s.eoff)
      for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
        v[0] += emsg[0 * ecap + e];
        v[1] += emsg[0 * ecap + e];
        v[2] += emsg[1 * ecap + e];
        v[3] += emsg[2 * ecap + e];
      }
  }
  for (int k = 0; k < 4; ++k) out[4 * r + k] = int64_t(v[k]);
}

// gbp_world_remove_robots: the entity is despawned (robot.rs:2171-2172 RobotDespawned, despawn_entity_after).
__global__ void k_remove_robots(Store s, int m, const int32_t *__restrict__ robots) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int32_t r = robots[k];
  s.gone[r] = 1.0f;
  s.idle[r] = 1;  // never iterated again; k_keep_gone_idle re-applies it after every gbp_world_set_comms
}
__global__ void k_keep_gone_idle(Store s) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < s.Nloc && s.gone[r] != 0.0f) s.idle[r] = 1;
}

// bounding box of nothing, candidate counter 0, and the cell-less dummy robot the padded candidate list points at
__global__ void k_box_init(int32_t *box, int32_t *cx_dummy, int32_t *cz_dummy) {
  if (threadIdx.x == 0) {
    box[0] = box[1] = INT32_MAX;
    box[2] = box[3] = INT32_MIN;
    box[4] = 0;
    *cx_dummy = *cz_dummy = gbp::kNoCell;
  }
}

__global__ void k_iota_gid(int32_t *gid, int32_t g0, int32_t n) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) gid[r] = g0 + r;
}

}  // namespace

// ---------------------------------------------------------------------------
// The shards that iterate one swarm together.  A plain single-GPU world is a group of one.
struct gbp_group {
  int ws = 1;
  bool nccl = false;                 // transport: NCCL send/recv (one member) or in-process copies
  gbp::NcclApi::comm_t comm = nullptr;
  std::vector<gbp_world *> members;  // shards living in this process, indexed by rank when !nccl
  bool committed = false;            // global ids fixed (gbp_world_commit_shards)
  bool halo_stale = true;            // some published record changed since the last halo exchange
  // gbp_world_{internal,external}_factor_iteration has been called and the matching *_variable_iteration has not:
  // the factor half runs fused with the variable half, so the state "after factors, before variables" is never
  // materialised; reads and state-changing calls are refused while a pair is open (pair_open)
  bool pending_internal_factor = false, pending_external_factor = false;
  // The topology pass of the NEXT tick depends on positions only, and those are final once
  // update_prior_of_current_state has run: the search is started right there on a side stream (group_topology_early)
  // and runs next to iterate_gbp; gbp_world_update_topology then only reads its sizes and applies it.
  bool early_pending = false;        // a search is in flight / done on the topo streams
  uint64_t topo_version = 0;         // bumped by whatever changes the search's inputs (positions, robots added / removed)
  uint64_t early_version = 0;        // topo_version the pending search was started from
  bool halo_pending = false;         // an exchange started right after the border robots' internal half is in flight
                                     // on the comm streams (group_launch); the next external half waits for it
  cudaStream_t shared_stream = nullptr;
  int refs = 0;
};

struct gbp_world {
  gbp_config_t cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Store s{};
  int p = 0;
  uint32_t epoch = 1;
  uint64_t robot_number = 1;  // RobotNumberGenerator (robot.rs:121-144)
  int64_t launches = 0;
  int64_t Ecap = 0;
  std::vector<uint32_t> timesteps;
  std::vector<int32_t> wp_off{0};
  std::vector<float> wp_xy;
  uint8_t *sdf_dev = nullptr;
  // two edge sets: the live one and a spare the next topology change is built into
  struct EdgeSet {
    int32_t *enbr = nullptr;   // neighbour slot
    int32_t *egid = nullptr;   // neighbour global id, ascending (== enbr, same allocation, when ws == 1)
    uint64_t *e_own = nullptr; // robot_number of the receiver's OWN factor toward the neighbour (i = 1)
    double *e_dsafe = nullptr;
    uint64_t *e_rnum = nullptr;
    uint32_t *e_birth = nullptr;
    uint8_t *e_frozen = nullptr;
    uint8_t *e_act = nullptr;  // scratch of one sub-step (k_edge_messages -> k_iterate)
    double *mir = nullptr;
    double *mu_frozen = nullptr;
    int64_t *map = nullptr;
    // MessageCount of the receiver's OWN InterRobot factors toward the neighbour, summed over the V-1 factors of the
    // edge: [0] sent (internal == external: every update sends one of each), [1] received internal, [2] received
    // external.  They go when the edge goes, like the counters of a deleted factor node.
    uint32_t *e_msg = nullptr;  // [3][cap]
    int64_t cap = 0;
  } edges[2];
  int cur = 0;
  int32_t *t_nlow = nullptr;
  int64_t *t_result_dev = nullptr, *t_result_host = nullptr;
  // topology scratch
  int32_t *t_cx = nullptr, *t_cz = nullptr, *t_idx = nullptr, *t_idx_sorted = nullptr;
  // strict_reference_quirks (single GPU): robots within range as their own CSR, largest lost neighbour, zombie flags
  int32_t *q_wnbr = nullptr, *q_maxlost = nullptr;
  int64_t *q_woff = nullptr;
  uint8_t *q_zombie = nullptr;
  int64_t q_wcap = 0, q_zcap = 0, q_ncap = 0;
  int32_t *t_park = nullptr;     // neighbour ids found by the counting pass, kNbrPark per own robot
  int32_t *t_box = nullptr;      // sharded neighbour search: own bounding box [0..3], candidates found [4]
  int64_t cand_cap = 0;          // entries the hash of a sharded world sorts (grows when a pass finds more)
  uint32_t *t_keys = nullptr, *t_keys_sorted = nullptr;
  int64_t *t_cnt = nullptr, *t_off = nullptr, *t_newcnt = nullptr, *t_newoff = nullptr;
  void *t_cub = nullptr;
  size_t t_cub_bytes = 0;
  int64_t t_cap = 0;
  // grow-only device scratch for read-backs / small uploads (no cudaMalloc per call)
  void *rb_dev = nullptr;
  size_t rb_bytes = 0;
  // asynchronous read-back: the AoS gather runs on the engine's stream, the device->host copies on
  // a second stream, so the next tick's kernels overlap the PCIe transfer
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_gathered = nullptr, ev_copied = nullptr;
  // host -> device inputs of a tick (gbp_world_set_comms, gbp_world_set_waypoint_index): staged through
  // double-buffered device slots on their own stream so that the call returns when the HOST buffer is
  // free again, not when the previous tick's kernels have drained (staged_upload)
  static constexpr int kUpKinds = 3;  // antenna, idle, next_wp
  cudaStream_t up_stream = nullptr;
  void *up_stage[kUpKinds][2] = {};
  size_t up_bytes[kUpKinds] = {};
  cudaEvent_t ev_up_done[kUpKinds][2] = {}, ev_up_used[kUpKinds][2] = {};
  int up_gen[kUpKinds] = {};
  bool copy_pending = false;
  // optional per-launch CUDA-event timing (bench.py roofline leg)
  struct Span {
    int kind;
    cudaEvent_t a, b;
  };
  bool profiling = false;
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[GBP_PROFILE_KINDS] = {0};
  int64_t prof_count[GBP_PROFILE_KINDS] = {0};
  // ---- sharding (gbp_shard.cuh) ------------------------------------------------------
  gbp_group *grp = nullptr;
  bool owns_stream = true;
  bool smem_opted_in[4] = {false, false, false, false};  // k_iterate<EXT,INT> dynamic shared memory opt-in
  bool axis_opted_in[12] = {};  // k_iterate_axis<EXT,INT,PART> likewise
  bool general_only = false;  // gbp_world_set_iterate_path: every robot through k_iterate
  bool fused_tick = false;    // gbp_world_set_iterate_path(w, 2): a whole iterate_gbp in one cooperative launch (k_tick_fused)
  int fused_occ = 0;          // resident CTAs per SM of k_tick_fused
  bool halo_overlap = true;   // sharded: border robots first, their halo behind the interior launch (GBP_HALO_OVERLAP=0:
                              // one launch for all robots, the exchange before the next external half)
  bool use_pdl = true;        // iterate kernels launched with programmatic stream serialization (GBP_PDL=0: off)
  double *dyn_tab_dev = nullptr;  // Store::dyn_tab while every robot added so far has the same t0 (radius)
  bool t0_seen = false, t0_uniform = true;
  float t0_first = 0.0f;
  int sm_count = 148;
  int edge_occ = 0;                // resident CTAs per SM of k_edge_messages
  bool iter_opted_in[4] = {false, false, false, false};
  int iter_occ[4] = {0, 0, 0, 0};  // resident CTAs per SM of k_iterate<EXT, INT>, by (EXT ? 2 : 0) + (INT ? 1 : 0)
  int par = 0;                // launch parity: which Store::gen_count the current launch appends to
  unsigned long long *coll_totals = nullptr;              // [0] Hit events so far, [1] pairs colliding now
  // robot-environment collisions (gbp_collide.cuh): the Colliders resource, CollisionHistory bits per (robot, collider)
  gbp::ColliderDev *env_cols = nullptr;
  float *env_verts = nullptr;
  int32_t env_ncol = 0, env_words = 0;
  uint32_t *env_state = nullptr, *env_hits = nullptr;  // [robots][words], [robots]
  int64_t env_robots = 0;                               // robots the two arrays are sized for
  unsigned long long *env_totals = nullptr;             // [0] Hit events so far, [1] pairs colliding now
  // position / velocity sample buffers (planner/tracking.rs)
  gbp::TrackRings trk{};
  uint64_t trk_duration_ns = 0;
  int32_t trk_seen = 0;  // robots that existed when the trackers last ran (later ones count as Changed<Transform>)
  gbp::ShardInfo sh{};          // ws, rank, gfirst
  int32_t Ntot = 0;             // robots of the whole swarm
  int32_t nghost = 0;
  bool force_rebuild = false;
  float *gpos = nullptr;        // [4][Ntot] x, z, radius, despawned flag of every robot by global id (ws > 1)
  bool any_gone = false;        // some own robot has been removed: set_comms keeps it idle
  // gbp_world_set_message_counting: MessageCount (factorgraph/mod.rs:103-137) kept by small accounting kernels
  // next to every half-step / prior update / topology change; off by default (the hot kernels never see it)
  bool count_messages = false;
  unsigned long long *msg_cnt = nullptr;  // [4][cap]: variables + non-InterRobot factors of each robot
  int64_t msg_cap = 0;
  float max_radius = 0.0f;         // over the own robots: the collision monitor only tests connected pairs
  std::vector<uint8_t> gone_host;  // host mirror of Store::gone
  int64_t n_gone = 0;
  int64_t gpos_cap = 0;
  int32_t *t_gflag = nullptr, *t_gslot = nullptr, *t_sflag = nullptr, *t_soff = nullptr;
  int64_t *t_ccnt = nullptr, *t_coff = nullptr;
  int64_t t_cap_tot = 0, t_cap_peer = 0;
  int32_t *t_err = nullptr;
  int32_t *sendlist = nullptr;
  int64_t sendlist_cap = 0;
  uint64_t *ckeys_s = nullptr, *cvals_s = nullptr, *ckeys_r = nullptr, *cvals_r = nullptr;
  int64_t cross_cap = 0;
  int64_t *hdr_send = nullptr, *hdr_recv = nullptr, *hdr_host = nullptr;  // 4 int64 per shard
  double *halo_send = nullptr, *halo_recv = nullptr;
  // halo / compute overlap: the robots of the send lists ("border") run first, their records travel on
  // comm_stream while the interior robots run on `stream`
  cudaStream_t comm_stream = nullptr;  // == stream for in-process shards (one device, one stream)
  cudaStream_t topo_stream = nullptr;    // the early topology search (== stream for in-process shards)
  cudaEvent_t ev_prior = nullptr, ev_topo = nullptr;
  bool topo_early = true;                // GBP_TOPO_EARLY=0: search inside gbp_world_update_topology as in round 1
  cudaStream_t border_stream = nullptr;  // highest priority: the border robots' launches run CONCURRENTLY with the
                                         // interior launch and finish first (== stream for in-process shards)
  cudaEvent_t ev_start = nullptr;        // everything before this half-step pair is done (border_stream waits for it)
  int32_t *border_gen_list = nullptr, *border_gen_count = nullptr;  // the border launches' own hand-over list
  int border_par = 0;
  cudaEvent_t ev_border = nullptr, ev_halo = nullptr, ev_fence = nullptr;
  uint32_t *border_words = nullptr;    // Store-side flags, one byte per own robot, as words for atomicOr
  int32_t *border_list = nullptr, *border_count = nullptr;
  int64_t border_cap = 0, border_words_cap = 0, border_gen_cap = 0;
  int64_t n_border_max = 0;            // length of the send lists = upper bound of *border_count
  int64_t halo_send_cap = 0, halo_recv_cap = 0;  // in doubles
  gbp::PeerOffsets ghost_po{}, send_po{};        // live halo layout (records per peer block)
  // results of the current topology pass, applied only when some shard's connectivity changed
  struct TopoPass {
    int64_t E1 = 0, total_new = 0, nghost = 0;
    bool changed = false;
    gbp::PeerOffsets ghost_po{}, send_po{}, cross_po{};
  } tp;
};

namespace {

int set_device(gbp_world *w) {
  CK(cudaSetDevice(w->device));
  return 0;
}

// State-changing and reading entry points call this first: between a *_factor_iteration call and its
// *_variable_iteration the reference has already run the factor half (factorgraph.rs:688-714, 719-760); the engine
// has not (it fuses the two), so it refuses to show or change that state rather than return something else.
int pair_open(const gbp_world *w, const char *what) {
  if (w->grp && (w->grp->pending_internal_factor || w->grp->pending_external_factor))
    return fail(GBP_ERR_STATE, std::string(what) + ": a factor half-iteration is open on this world (call the matching "
                                                   "*_variable_iteration first)");
  return 0;
}

void refresh_scalars(gbp_world *w) {
  const gbp_config_t &c = w->cfg;
  Store &s = w->s;
  auto inv_sq = [](float sigma) {
    const double d = double(sigma);
    return 1.0 / (d * d);
  };
  s.qs_dyn = inv_sq(c.sigma_factor_dynamics);
  s.lm_ir = inv_sq(c.sigma_factor_interrobot);
  s.lm_obs = inv_sq(c.sigma_factor_obstacle);
  s.lm_trk = inv_sq(c.sigma_factor_tracking);
  s.tiny_scale = double(1e-6f);  // InterRobotFactor::TINY_OFFSET_SCALE (interrobot.rs:52)
  s.trk_switch_padding = double(c.tracking_switch_padding);
  s.trk_attraction = double(c.tracking_attraction_distance);
  s.en_dyn = c.enable_dynamic;
  s.en_ir = c.enable_interrobot;
  s.en_obs = c.enable_obstacle;
  s.en_trk = c.enable_tracking;
  s.world_w = c.world_width;
  s.world_h = c.world_height;
  s.jac_delta = (c.world_width / double(uint32_t(s.sdf_w)) + c.world_height / double(uint32_t(s.sdf_h))) / 2.0;
  // ObstacleFactor::measure (obstacle.rs:141-160): offsets and scales are functions of the world and image size only
  s.sdf_xo = s.world_w / 2.0;
  s.sdf_yo = s.world_h / 2.0;
  s.sdf_xs = double(uint32_t(s.sdf_w)) / s.world_w;
  s.sdf_ys = double(uint32_t(s.sdf_h)) / s.world_h;
}

cudaEvent_t take_event(gbp_world *w) {
  if (!w->ev_pool.empty()) {
    cudaEvent_t e = w->ev_pool.back();
    w->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct ProfileScope {
  gbp_world *w;
  cudaStream_t st;
  gbp_world::Span sp{};
  ProfileScope(gbp_world *w_, int kind, cudaStream_t st_ = nullptr) : w(w_), st(st_ ? st_ : w_->stream) {
    if (!w->profiling) return;
    sp.kind = kind;
    sp.a = take_event(w);
    sp.b = take_event(w);
    cudaEventRecord(sp.a, st);
  }
  ~ProfileScope() {
    if (!w->profiling) return;
    cudaEventRecord(sp.b, st);
    w->spans.push_back(sp);
  }
};
int drain_profile(gbp_world *w) {
  if (w->spans.empty()) return 0;
  CK(cudaStreamSynchronize(w->stream));
  for (auto &sp : w->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      w->prof_ms[sp.kind] += ms;
      w->prof_count[sp.kind] += 1;
    }
    w->ev_pool.push_back(sp.a);
    w->ev_pool.push_back(sp.b);
  }
  w->spans.clear();
  return 0;
}

using EdgeSet = gbp_world::EdgeSet;

void mark_halo_stale(gbp_world *w) {
  if (w->grp) w->grp->halo_stale = true;
}

// ---- transport ---------------------------------------------------------------------
// Runs one exchange for every shard of the group living in this process (gbp_comm.cuh).
int exchange(gbp_group *g, std::vector<gbp::XferPlan> &plans, bool on_comm_stream = false, bool home_is_topo = false) {
  if (g->ws == 1) return 0;
  if (g->nccl) {
    gbp::NcclApi &api = gbp::nccl_api();
    gbp_world *w = g->members[0];
    // Every NCCL call of the communicator is enqueued on ONE stream (comm_stream), so the order of its
    // operations is the order of the calls on every rank; a caller working on w->stream is fenced in and out.
    cudaStream_t nst = w->comm_stream;
    cudaStream_t home = home_is_topo ? w->topo_stream : w->stream;
    if (!on_comm_stream) {
      CK(cudaEventRecord(w->ev_fence, home));
      CK(cudaStreamWaitEvent(nst, w->ev_fence, 0));
    }
    int rc = api.GroupStart();
    if (rc) return fail(GBP_ERR_NCCL, std::string("ncclGroupStart: ") + api.GetErrorString(rc));
    for (const gbp::Xfer &x : plans[0].sends)
      if (x.bytes && (rc = api.Send(x.ptr, x.bytes, gbp::kNcclUint8, x.peer, g->comm, nst)))
        return fail(GBP_ERR_NCCL, std::string("ncclSend: ") + api.GetErrorString(rc));
    for (const gbp::Xfer &x : plans[0].recvs)
      if (x.bytes && (rc = api.Recv(x.ptr, x.bytes, gbp::kNcclUint8, x.peer, g->comm, nst)))
        return fail(GBP_ERR_NCCL, std::string("ncclRecv: ") + api.GetErrorString(rc));
    rc = api.GroupEnd();
    if (rc) return fail(GBP_ERR_NCCL, std::string("ncclGroupEnd: ") + api.GetErrorString(rc));
    if (!on_comm_stream) {
      CK(cudaEventRecord(w->ev_fence, nst));
      CK(cudaStreamWaitEvent(home, w->ev_fence, 0));
    }
    return 0;
  }
  // in-process shards on one device and one stream: the k-th send a -> b is copied into the
  // k-th receive of b from a
  const int ws = g->ws;
  for (int a = 0; a < ws; ++a) {
    std::vector<size_t> next(size_t(ws), 0);
    for (const gbp::Xfer &sx : plans[a].sends) {
      const int b = sx.peer;
      const std::vector<gbp::Xfer> &rv = plans[b].recvs;
      size_t &k = next[b];
      while (k < rv.size() && rv[k].peer != a) ++k;
      if (k >= rv.size() || rv[k].bytes != sx.bytes)
        return fail(GBP_ERR_STATE, "local exchange: send/receive lists of two shards do not pair up");
      if (sx.bytes) CK(cudaMemcpyAsync(rv[k].ptr, sx.ptr, sx.bytes, cudaMemcpyDeviceToDevice, g->shared_stream));
      ++k;
    }
  }
  return 0;
}

template <class T>
int ensure_buf(gbp_world *w, T *&p, int64_t &cap, int64_t need) {
  if (need <= cap) return 0;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(p);
  p = nullptr;
  cap = 0;
  const int64_t ncap = need + need / 4 + 1024;
  CK(dalloc(p, size_t(ncap)));
  cap = ncap;
  return 0;
}

// ---- per-sub-step halo: published records of border robots -> ghost slots of the peers ----
// ahead == false: the exchange an external half needs, on the shards' own streams, from / into pub[p].
// ahead == true: started by group_launch right after the border robots' internal half, while the interior
// robots are still running: from / into pub[1 - p] (the buffer that half writes and the NEXT external half
// reads), on the comm streams, between the events ev_border (border records written) and ev_halo
// (ghost slots filled).  Every rank posts the same sequence of exchanges either way.
int group_halo(gbp_group *g, bool ahead = false) {
  if (g->ws == 1) return 0;
  const int ws = g->ws;
  std::vector<gbp::XferPlan> plans(g->members.size());
  for (size_t m = 0; m < g->members.size(); ++m) {
    gbp_world *w = g->members[m];
    CK(cudaSetDevice(w->device));
    Store &s = w->s;
    cudaStream_t st = ahead ? w->comm_stream : w->stream;
    const int pb = ahead ? 1 - w->p : w->p;
    if (ahead && w->comm_stream != w->stream) CK(cudaStreamWaitEvent(w->comm_stream, w->ev_border, 0));
    const int64_t hd = gbp::halo_doubles_per_robot(s.V);
    const int64_t nsend = w->send_po.start[ws];
    if (nsend > 0) {
      gbp::k_halo_pack<<<blocks_for(nsend * (s.V - 1), 256), 256, 0, st>>>(s, pb, ws, w->send_po, w->sendlist,
                                                                        w->halo_send);
      CK(cudaGetLastError());
      w->launches += 1;
    }
    for (int q = 0; q < ws; ++q) {
      if (q == w->sh.rank) continue;
      const int64_t cs = w->send_po.start[q + 1] - w->send_po.start[q];
      const int64_t cr = w->ghost_po.start[q + 1] - w->ghost_po.start[q];
      if (cs) plans[m].sends.push_back({q, w->halo_send + w->send_po.start[q] * hd, size_t(cs * hd * 8)});
      if (cr) plans[m].recvs.push_back({q, w->halo_recv + w->ghost_po.start[q] * hd, size_t(cr * hd * 8)});
    }
  }
  {
    gbp_world *w0 = g->members[0];
    gbp_world::Span sp{};
    if (w0->profiling) {  // the exchange alone, timed on the stream it runs on
      sp.kind = GBP_PROFILE_HALO;
      sp.a = take_event(w0);
      sp.b = take_event(w0);
      cudaEventRecord(sp.a, ahead ? w0->comm_stream : w0->stream);
    }
    if (int rc = exchange(g, plans, ahead)) return rc;
    if (w0->profiling) {
      cudaEventRecord(sp.b, ahead ? w0->comm_stream : w0->stream);
      w0->spans.push_back(sp);
    }
  }
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    cudaStream_t st = ahead ? w->comm_stream : w->stream;
    const int pb = ahead ? 1 - w->p : w->p;
    if (w->nghost > 0) {
      gbp::k_halo_unpack<<<blocks_for(int64_t(w->nghost) * (w->s.V - 1), 256), 256, 0, st>>>(w->s, pb, ws, w->ghost_po,
                                                                                           w->halo_recv);
      CK(cudaGetLastError());
      w->launches += 1;
    }
    if (ahead && w->comm_stream != w->stream) CK(cudaEventRecord(w->ev_halo, w->comm_stream));
  }
  g->halo_stale = false;
  g->halo_pending = ahead;
  return 0;
}

// Work queued on the shards' own streams from here on sees the exchange that was started ahead (if any) finished:
// called before anything that reads ghost slots or rebuilds what the exchange uses (send lists, buffers).
int group_halo_join(gbp_group *g) {
  if (!g->halo_pending) return 0;
  for (gbp_world *w : g->members)
    if (w->comm_stream != w->stream) {
      CK(cudaSetDevice(w->device));
      CK(cudaStreamWaitEvent(w->stream, w->ev_halo, 0));
    }
  g->halo_pending = false;
  return 0;
}

// The external half that follows needs current ghost records: an exchange started ahead is waited for; one is
// made now if records changed since (or none was started).
int group_halo_ready(gbp_group *g) {
  if (g->ws == 1) return 0;
  if (int rc = group_halo_join(g)) return rc;
  // Always exchanged before an external half otherwise (never skipped on a per-shard "nothing changed"
  // guess: every rank must post the same sequence of transfers).
  if (g->halo_stale) return group_halo(g, false);
  return 0;
}

// Kernel launch with programmatic stream serialization (PDL): the kernel may begin while the previous kernel of the
// stream is draining; it orders itself with griddepcontrol.wait (gbp_iterate.cuh pdl_wait).
template <class... KArgs, class... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                       Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && GBP_PDL) ? 1 : 0;  // without the kernels' griddepcontrol.wait the attribute would be unsafe
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Warps that share one robot's (edge, variable) pairs in k_edge_messages: enough of them to fill the GPU twice when
// the swarm is small (at most `nrob` robots can be on the list), one per robot when the swarm alone does.
int edge_chunks(const gbp_world *w, int64_t nrob) {
  const int64_t want = int64_t(w->sm_count) * 16 * 2;  // 16 resident warps per SM
  int chunks = 1;
  while (chunks < 32 && nrob * chunks < want) chunks *= 2;
  return chunks;
}

// One iterate launch pair (k_iterate_axis, then k_iterate over what it handed over) for one part of a shard:
// part 0 = every own robot, 1 = the border robots (send lists), 2 = the others.
template <bool EXT, bool INT>
int launch_iterate(gbp_world *w, int part) {
  // The border part runs on its own high-priority stream with its own hand-over list, concurrently with the
  // interior part on w->stream (group_launch orders them with events).
  Store s = w->s;
  cudaStream_t st = w->stream;
  int *par = &w->par;
  if (part == 1) {
    st = w->border_stream;
    s.gen_list = w->border_gen_list;
    s.gen_count = w->border_gen_count;
    par = &w->border_par;
  }
  const int which = (EXT ? 2 : 0) + (INT ? 1 : 0);
  if (s.Nloc == 0) return 0;
  if (!w->iter_opted_in[which]) {
    CK(cudaFuncSetAttribute(gbp::k_iterate<EXT, INT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            int(gbp::kIterSmemBytes)));
    w->iter_opted_in[which] = true;
  }
  const int rpw = 32 / s.V;
  const int wpb = gbp::kIterBlock / 32;
  const int kind = EXT ? (INT ? GBP_PROFILE_ITERATE_EXT_INT : GBP_PROFILE_ITERATE_EXT) : GBP_PROFILE_ITERATE_INT;
  if (w->general_only) {
    const int64_t warps = (int64_t(s.Nloc) + rpw - 1) / rpw;
    const unsigned grid = unsigned((warps + wpb - 1) / wpb);
    ProfileScope ps(w, kind);
#if GBP_EDGE_SPLIT
    if (EXT && s.E > 0) {  // the neighbours' InterRobot factors first, one thread per (edge, variable)
      const int chunks = edge_chunks(w, s.Nloc);
      const unsigned eg = unsigned(std::min<int64_t>((int64_t(s.Nloc) * chunks + 3) / 4, int64_t(w->sm_count) * 16));
      gbp::k_edge_messages<<<eg, gbp::kEdgeBlock, 0, st>>>(s, w->p, -1, chunks);
      w->launches += 1;
    }
#endif
    gbp::k_iterate<EXT, INT><<<grid, gbp::kIterBlock, gbp::kIterSmemBytes, st>>>(s, w->p, w->epoch, -1);
    w->launches += 1;
  } else {
    // the decoupled robots (two lanes per variable), then whatever that kernel handed over
    const gbp::AxisGeom q = gbp::axis_geom(s.V);
    if (part == 2 && !w->border_words) part = 0;  // no topology pass has flagged anything yet
    auto kernel = part == 0 ? gbp::k_iterate_axis<EXT, INT, 0>
                            : (part == 1 ? gbp::k_iterate_axis<EXT, INT, 1> : gbp::k_iterate_axis<EXT, INT, 2>);
    bool &opted = w->axis_opted_in[which * 3 + part];
    if (!opted) {
      CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(q.smem)));
      opted = true;
    }
    const int64_t nrob = part == 1 ? w->n_border_max : int64_t(s.Nloc);
    if (nrob > 0) {
      {
        ProfileScope ps(w, part == 1 ? int(GBP_PROFILE_ITERATE_BORDER) : kind, st);
        CK(launch_pdl(w->use_pdl, kernel, blocks_for(nrob, q.rpc), unsigned(q.threads), q.smem, st, s, w->p, w->epoch,
                      q.rpc, *par, static_cast<const int32_t *>(w->border_list),
                      static_cast<const int32_t *>(w->border_count),
                      reinterpret_cast<const uint8_t *>(w->border_words)));
      }
      CK(cudaGetLastError());
      const int64_t warps = (nrob + rpw - 1) / rpw;
      // a persistent grid of exactly the CTAs that fit at once: the kernel walks the hand-over list with a grid
      // stride, so a grid of 4 CTAs per SM at 3 resident ran a second, one-third-full wave (r02y: SMs busy 75 %)
      int &occ = w->iter_occ[which];
      if (occ <= 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gbp::k_iterate<EXT, INT>, gbp::kIterBlock,
                                                         gbp::kIterSmemBytes));
        if (occ <= 0) occ = 1;
        if (const char *e = std::getenv("GBP_ITER_GRID_PER_SM")) occ = std::max(1, atoi(e));
      }
      const unsigned grid = unsigned(std::min<int64_t>((warps + wpb - 1) / wpb, int64_t(w->sm_count) * occ));
      ProfileScope pg(w, GBP_PROFILE_ITERATE_GENERAL, st);
#if GBP_EDGE_SPLIT
      if (EXT && s.E > 0) {
        // InterRobot factors of the handed-over robots first, one thread per (edge, variable); a warp per robot, a
        // persistent grid (the list's length is only known on the device)
        if (w->edge_occ <= 0) {
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->edge_occ, gbp::k_edge_messages, gbp::kEdgeBlock, 0));
          if (w->edge_occ <= 0) w->edge_occ = 1;
        }
        const int chunks = edge_chunks(w, nrob);
        const unsigned eg = unsigned(std::min<int64_t>((nrob * chunks + 3) / 4, int64_t(w->sm_count) * w->edge_occ));
        gbp::k_edge_messages<<<eg, gbp::kEdgeBlock, 0, st>>>(s, w->p, *par, chunks);
        w->launches += 1;
      }
#endif
      CK(launch_pdl(w->use_pdl, gbp::k_iterate<EXT, INT>, grid, unsigned(gbp::kIterBlock), gbp::kIterSmemBytes, st, s, w->p,
                    w->epoch, *par));
      *par ^= 1;
      w->launches += 2;
    }
  }
  CK(cudaGetLastError());
  if (w->spans.size() > 4096) {
    if (int rc = drain_profile(w)) return rc;
  }
  return 0;
}

// One half-step pair for every shard of the group: the external half reads the neighbours'
// published records, so the halo must be current before it; the internal half republishes.
// Sharded worlds run an internal half in two parts — border robots, then the rest — and start the
// exchange of the new border records in between, so that it overlaps the interior launch
// (the two delivery loops robot.rs:1814-1831, :1843-1858 for pairs split over two GPUs).
template <bool EXT, bool INT>
int group_launch(gbp_group *g) {
  if (EXT) {
    if (int rc = group_halo_ready(g)) return rc;
  }
  for (gbp_world *w : g->members)
    if (w->count_messages && w->s.Nloc > 0) {
      CK(cudaSetDevice(w->device));
      k_msg_half<EXT, INT><<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(w->s, w->msg_cap, w->msg_cnt,
                                                                             w->edges[w->cur].cap, w->edges[w->cur].e_msg);
      CK(cudaGetLastError());
      w->launches += 1;
    }
  bool split = INT && g->ws > 1 && g->members[0]->halo_overlap;
  for (gbp_world *w : g->members) {
    w->epoch += 1;  // every shard steps its epoch, with or without robots
    split = split && !w->general_only;
  }
  if (!split) {
    for (gbp_world *w : g->members) {
      CK(cudaSetDevice(w->device));
      if (int rc = launch_iterate<EXT, INT>(w, 0)) return rc;
    }
    if (INT) g->halo_stale = true;
  } else {
    // border robots on the high-priority stream (behind everything queued so far), the other robots on the shard's own
    // stream at the same time, the halo of the new border records on the comm stream behind the border launches.
    // The interior launch is enqueued BEFORE the host builds the exchange (NCCL group calls take tens of
    // microseconds of host time, during which the shard's own stream would otherwise sit empty: r02t8).
    for (gbp_world *w : g->members) {
      CK(cudaSetDevice(w->device));
      if (w->border_stream != w->stream) {
        CK(cudaEventRecord(w->ev_start, w->stream));
        CK(cudaStreamWaitEvent(w->border_stream, w->ev_start, 0));
      }
      if (int rc = launch_iterate<EXT, INT>(w, 1)) return rc;
      if (w->border_stream != w->stream) CK(cudaEventRecord(w->ev_border, w->border_stream));
      if (int rc = launch_iterate<EXT, INT>(w, 2)) return rc;
      // what follows on the shard's own stream sees the border robots' results too
      if (w->border_stream != w->stream) CK(cudaStreamWaitEvent(w->stream, w->ev_border, 0));
    }
    if (int rc = group_halo(g, true)) return rc;
  }
  if (INT)
    for (gbp_world *w : g->members) w->p ^= 1;
  return 0;
}

// iterate_gbp_v2 (robot.rs:1769-1861): flatten the schedule into half-steps
// I (internal factor+variable) and E (external factor+variable); an E directly
// followed by an I runs as one fused launch.
// The launches of run_schedule as one cooperative launch (k_tick_fused) when the world asked for it and fits:
// one GPU, every robot's warp resident at once, at most 64 launches, no per-launch accounting wanted.
// Returns 1 if it ran, 0 if the caller has to launch half-step by half-step, < 0 on error.
int try_fused_tick(gbp_group *g, const std::vector<char> &ph) {
  gbp_world *w = g->members[0];
  if (!w->fused_tick || g->ws != 1 || w->count_messages || w->profiling || w->s.Nloc == 0) return 0;
  gbp::TickPlan plan{};
  int n_int = 0;
  for (size_t k = 0; k < ph.size();) {
    if (plan.n >= 64) return 0;
    const bool fused = ph[k] == 'E' && k + 1 < ph.size() && ph[k + 1] == 'I';
    plan.ext[plan.n] = ph[k] == 'E';
    plan.in[plan.n] = fused || ph[k] == 'I';
    n_int += plan.in[plan.n];
    plan.n += 1;
    k += fused ? 2 : 1;
  }
  if (plan.n == 0) return 1;
  CK(cudaSetDevice(w->device));
  if (w->fused_occ <= 0) {
    CK(cudaFuncSetAttribute(gbp::k_tick_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, int(gbp::kIterSmemBytes)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->fused_occ, gbp::k_tick_fused, gbp::kIterBlock,
                                                     gbp::kIterSmemBytes));
    if (w->fused_occ <= 0) return 0;
  }
  Store s = w->s;
  const int rpw = 32 / s.V, wpb = gbp::kIterBlock / 32;
  const int64_t warps = (int64_t(s.Nloc) + rpw - 1) / rpw;
  const int64_t resident = int64_t(w->sm_count) * w->fused_occ;
  const int64_t need = (warps + wpb - 1) / wpb;
  if (need > resident) return 0;  // a grid barrier needs every CTA resident, and one pass per half-step
  // the edge kernel's body shares this grid: as many warps per robot as the grid has to spare
  int chunks = 1;
  while (chunks < 32 && int64_t(s.Nloc) * chunks * 2 <= resident * wpb) chunks *= 2;
  const unsigned grid = unsigned(std::max<int64_t>(need, std::min<int64_t>(resident, (int64_t(s.Nloc) * chunks + wpb - 1) / wpb)));
  int p = w->p;
  uint32_t epoch = w->epoch;
  void *args[] = {&s, &p, &epoch, &chunks, &plan};
  CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(gbp::k_tick_fused), dim3(grid), dim3(gbp::kIterBlock), args,
                                 gbp::kIterSmemBytes, w->stream));
  w->epoch += uint32_t(plan.n);
  if (n_int & 1) w->p ^= 1;
  w->launches += 1;
  return 1;
}

int run_schedule(gbp_group *g, int n, const uint8_t *internal, const uint8_t *external) {
  std::vector<char> ph;
  ph.reserve(size_t(n) * 2);
  for (int i = 0; i < n; ++i) {
    if (internal[i]) ph.push_back('I');
    if (external[i]) ph.push_back('E');
  }
  if (int rc = try_fused_tick(g, ph)) return rc < 0 ? rc : 0;
  for (size_t k = 0; k < ph.size();) {
    int rc;
    if (ph[k] == 'E' && k + 1 < ph.size() && ph[k + 1] == 'I') {
      rc = group_launch<true, true>(g);
      k += 2;
    } else if (ph[k] == 'E') {
      rc = group_launch<true, false>(g);
      k += 1;
    } else {
      rc = group_launch<false, true>(g);
      k += 1;
    }
    if (rc) return rc;
  }
  return 0;
}

int ensure_scratch(gbp_world *w, size_t bytes) {
  if (w->copy_pending) {  // an asynchronous read-back is still copying out of the scratch
    CK(cudaEventSynchronize(w->ev_copied));
    w->copy_pending = false;
  }
  if (bytes <= w->rb_bytes) return 0;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(w->rb_dev);
  w->rb_dev = nullptr;
  w->rb_bytes = 0;
  bytes += bytes / 8 + 4096;
  CK(cudaMalloc(&w->rb_dev, bytes));
  w->rb_bytes = bytes;
  return 0;
}

void bind_edge_set(gbp_world *w) {
  const EdgeSet &e = w->edges[w->cur];
  Store &s = w->s;
  s.enbr = e.enbr;
  s.e_dsafe = e.e_dsafe;
  s.e_rnum = e.e_rnum;
  s.e_birth = e.e_birth;
  s.e_frozen = e.e_frozen;
  s.e_act = e.e_act;
  s.mir = e.mir;
  s.mu_frozen = e.mu_frozen;
  s.EV = e.cap * (s.V - 1);
}

// Store::ell_* follows the live CSR: after every change of the edge lists, of the neighbour slots or of the safety
// distances.
int refresh_ell(gbp_world *w) {
  Store &s = w->s;
  if (!s.ell_nbr || s.Nloc == 0 || !s.eoff) return 0;
  k_ell_fill<<<blocks_for(int64_t(s.Nloc) * gbp::kEll, 256), 256, 0, w->stream>>>(s);
  CK(cudaGetLastError());
  w->launches += 1;
  return 0;
}

void free_edge_set(gbp_world *w, EdgeSet *e) {
  if (e->egid != e->enbr) cudaFree(e->egid);
  cudaFree(e->enbr); cudaFree(e->e_own); cudaFree(e->e_dsafe); cudaFree(e->e_rnum); cudaFree(e->e_birth);
  cudaFree(e->e_frozen); cudaFree(e->e_act); cudaFree(e->mir); cudaFree(e->map); cudaFree(e->mu_frozen); cudaFree(e->e_msg);
  *e = EdgeSet();
}

int grow_edge_set(gbp_world *w, EdgeSet *e, int64_t cap) {
  CK(cudaStreamSynchronize(w->stream));
  free_edge_set(w, e);
  const int Vm1 = w->s.V - 1;
  CK(dalloc(e->enbr, size_t(cap)));
  if (w->sh.ws > 1) CK(dalloc(e->egid, size_t(cap)));
  else e->egid = e->enbr;  // one GPU: global id == slot
  CK(dalloc(e->e_own, size_t(cap)));
  CK(dalloc(e->e_dsafe, size_t(cap)));
  CK(dalloc(e->e_rnum, size_t(cap)));
  CK(dalloc(e->e_birth, size_t(cap)));
  CK(dalloc(e->e_frozen, size_t(cap)));
  CK(dalloc(e->e_act, size_t(cap)));
  CK(dalloc(e->mu_frozen, size_t(2) * size_t(cap) * Vm1));
  CK(dalloc(e->map, size_t(cap)));
  CK(dalloc(e->mir, size_t(6) * size_t(cap) * Vm1));
  CK(dalloc(e->e_msg, size_t(3) * size_t(cap)));
  e->cap = cap;
  return 0;
}

constexpr int kResultWords = 4 + 3 * (gbp::kMaxShards + 1) + 1;

int ensure_topology_scratch(gbp_world *w) {
  const int64_t nloc = w->s.Nloc, ntot = w->Ntot, ws = w->sh.ws;
  if (!w->t_result_dev) {
    CK(dalloc(w->t_box, 5));  // bounding box of the own robots (4 ordered ints) + candidate count
    CK(dalloc(w->t_result_dev, kResultWords));
    CK(cudaMallocHost(reinterpret_cast<void **>(&w->t_result_host), kResultWords * sizeof(int64_t)));
    CK(dalloc(w->t_err, 1));
    CK(cudaMemset(w->t_err, 0, sizeof(int32_t)));
    CK(dalloc(w->hdr_send, 4));
    CK(dalloc(w->hdr_recv, 4 * gbp::kMaxShards));
    CK(cudaMallocHost(reinterpret_cast<void **>(&w->hdr_host), 4 * (gbp::kMaxShards + 1) * sizeof(int64_t)));
  }
  bool regrow_cub = false;
  if (!w->t_cnt || nloc > w->t_cap) {
    cudaFree(w->t_nlow); cudaFree(w->t_cnt); cudaFree(w->t_off); cudaFree(w->t_newcnt); cudaFree(w->t_newoff);
    cudaFree(w->t_park);
    CK(dalloc(w->t_park, size_t(nloc) * gbp::kNbrPark + 1));
    CK(dalloc(w->t_nlow, nloc));
    CK(dalloc(w->t_cnt, nloc + 1)); CK(dalloc(w->t_off, nloc + 1));
    CK(dalloc(w->t_newcnt, nloc + 1)); CK(dalloc(w->t_newoff, nloc + 1));
    w->t_cap = nloc;
    regrow_cub = true;
  }
  if (!w->t_cx || ntot > w->t_cap_tot) {
    cudaFree(w->t_cx); cudaFree(w->t_cz); cudaFree(w->t_idx); cudaFree(w->t_idx_sorted);
    cudaFree(w->t_keys); cudaFree(w->t_keys_sorted); cudaFree(w->t_gflag); cudaFree(w->t_gslot);
    CK(dalloc(w->t_cx, ntot + 1)); CK(dalloc(w->t_cz, ntot + 1));  // slot ntot: the dummy of the padded candidate list
    CK(dalloc(w->t_idx, ntot)); CK(dalloc(w->t_idx_sorted, ntot));
    CK(dalloc(w->t_keys, ntot)); CK(dalloc(w->t_keys_sorted, ntot));
    CK(dalloc(w->t_gflag, ntot + 1)); CK(dalloc(w->t_gslot, ntot + 1));
    w->t_cap_tot = ntot;
    regrow_cub = true;
  }
  if (ws > 1 && (!w->t_sflag || ws * nloc > w->t_cap_peer)) {
    cudaFree(w->t_sflag); cudaFree(w->t_soff); cudaFree(w->t_ccnt); cudaFree(w->t_coff);
    CK(dalloc(w->t_sflag, ws * nloc + 1)); CK(dalloc(w->t_soff, ws * nloc + 1));
    CK(dalloc(w->t_ccnt, ws * nloc + 1)); CK(dalloc(w->t_coff, ws * nloc + 1));
    w->t_cap_peer = ws * nloc;
    regrow_cub = true;
  }
  if (regrow_cub || !w->t_cub) {
    size_t b[5] = {0, 0, 0, 0, 0};
    cub::DeviceRadixSort::SortPairs(nullptr, b[0], w->t_keys, w->t_keys_sorted, w->t_idx, w->t_idx_sorted,
                                    int(w->t_cap_tot));
    cub::DeviceScan::ExclusiveSum(nullptr, b[1], w->t_cnt, w->t_off, int(w->t_cap + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, b[2], w->t_gflag, w->t_gslot, int(w->t_cap_tot + 1));
    if (ws > 1) {
      cub::DeviceScan::ExclusiveSum(nullptr, b[3], w->t_sflag, w->t_soff, int(w->t_cap_peer + 1));
      cub::DeviceScan::ExclusiveSum(nullptr, b[4], w->t_ccnt, w->t_coff, int(w->t_cap_peer + 1));
    }
    size_t need = *std::max_element(b, b + 5) + 256;
    if (need > w->t_cub_bytes) {
      CK(cudaStreamSynchronize(w->stream));
      cudaFree(w->t_cub);
      w->t_cub = nullptr;
      CK(cudaMalloc(&w->t_cub, need));
      w->t_cub_bytes = need;
    }
  }
  return 0;
}

// Re-stride every per-variable / per-robot plane to `newcap` robot slots, keeping the own robots.
int group_topology_early_wait(gbp_group *g);
// Whatever changes what a neighbour search reads (positions, despawned flags, the robot count, the arrays
// themselves) calls this first: a search in flight is waited for and its result is not used.
int topology_inputs_change(gbp_world *w) {
  w->grp->topo_version += 1;
  return group_topology_early_wait(w->grp);
}

int reserve_robots(gbp_world *w, int64_t newcap) {
  Store &s = w->s;
  if (newcap <= s.cap) return 0;
  if (int rc = topology_inputs_change(w)) return rc;  // the arrays move
  cudaStream_t st = w->stream;
  const int V = s.V;
  // whole tiles (gbp_store.cuh)
  const int64_t oldNV = s.NV, newNV = (newcap * V + gbp::kTile - 1) / gbp::kTile * gbp::kTile, used = int64_t(s.Nloc) * V, oldcap = s.cap, keep = s.Nloc;
  CK(regrow(s.prior_eta, 4, oldNV, newNV, used, st, true));
  CK(regrow(s.prior_lam, 1, oldNV, newNV, used, st));
  CK(regrow(s.pub[0], gbp::kRec, oldNV, newNV, used, st, true));
  CK(regrow(s.pub[1], gbp::kRec, oldNV, newNV, used, st, true));
  CK(regrow(s.pub_epoch[0], 1, oldNV, newNV, used, st));
  CK(regrow(s.pub_epoch[1], 1, oldNV, newNV, used, st));
  CK(regrow(s.bel_ext, gbp::kRec, oldNV, newNV, used, st, true));
  CK(regrow(s.mu_ext, 2, oldNV, newNV, used, st, true));
  CK(regrow(s.cov, 16, oldNV, newNV, used, st, true));
  CK(regrow(s.valid, 1, oldNV, newNV, used, st));
  CK(regrow(s.cov_lazy, 1, oldNV, newNV, used, st));
  for (int b = 0; b < 2; ++b) {
    CK(regrow(s.m_dynL[b], 20, oldNV, newNV, used, st, true));
    CK(regrow(s.m_dynR[b], 20, oldNV, newNV, used, st, true));
  }
  CK(regrow(s.m_obs, 4, oldNV, newNV, used, st, true));
  CK(regrow(s.m_trk, 3, oldNV, newNV, used, st, true));
  CK(regrow(s.dyn_c, 4, oldNV, newNV, used, st, true));
  CK(regrow(s.trk_record, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_timeout, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_seed, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_last, 2, oldNV, newNV, used, st, true));
  CK(regrow(s.trk_value, 1, oldNV, newNV, used, st));
  CK(regrow(s.radius, 1, oldcap, newcap, keep, st));
  CK(regrow(s.t0, 1, oldcap, newcap, keep, st));
  CK(regrow(s.pos, 2, oldcap, newcap, keep, st));
  CK(regrow(s.antenna, 1, oldcap, newcap, keep, st));
  CK(regrow(s.idle, 1, oldcap, newcap, keep, st));
  CK(regrow(s.finished, 1, oldcap, newcap, keep, st));
  CK(regrow(s.gone, 1, oldcap, newcap, keep, st));
  if (w->count_messages) {
    CK(regrow(w->msg_cnt, 4, w->msg_cap, newcap, keep, st));
    w->msg_cap = newcap;
  }
  CK(regrow(s.latest, 1, oldcap, newcap, keep, st));
  CK(regrow(s.iter_factor, 1, oldcap, newcap, keep, st));
  CK(regrow(s.mode, 1, oldcap, newcap, keep, st));
  CK(regrow(s.gen_list, 1, oldcap, newcap, 0, st));
  if (!s.gen_count) {
    CK(dalloc(s.gen_count, 2));
    CK(cudaMemsetAsync(s.gen_count, 0, 2 * sizeof(int32_t), st));
  }
  CK(regrow(s.gid, 1, oldcap, newcap, keep, st));
  CK(regrow(s.next_wp, 1, oldcap, newcap, keep, st));
  CK(regrow(s.coll_hits, 1, oldcap, newcap, keep, st));
  CK(regrow(s.nlow, 1, oldcap, newcap, keep, st));
#if GBP_AXIS_ELL
  CK(regrow(s.ell_nbr, 1, oldcap * gbp::kEll, newcap * gbp::kEll, keep * gbp::kEll, st));
  CK(regrow(s.ell_birth, 1, oldcap * gbp::kEll, newcap * gbp::kEll, keep * gbp::kEll, st));
  CK(regrow(s.ell_dsafe, 1, oldcap * gbp::kEll, newcap * gbp::kEll, keep * gbp::kEll, st));
#endif
  s.NV = newNV;
  s.cap = newcap;
  // ghost slots were dropped by the re-stride: the next topology pass rebuilds them
  s.N = s.Nloc;
  if (w->nghost > 0) w->force_rebuild = true;
  mark_halo_stale(w);
  return 0;
}

// ---- topology, phase 1 (per shard): neighbour search, diff against the live CSR, ghost and
// send lists; one host sync reads the sizes.
// mode 0: the whole pass on the shard's own stream, sizes read back before returning.
// mode 1: enqueue only, on the topo stream (group_topology_early); mode 2: the pass enqueued by mode 1 has
// completed (the caller waited for ev_topo) — read its sizes; if a buffer turns out too small the affected part
// runs again on the shard's own stream.
int topo_search(gbp_world *w, int mode = 0) {
  CK(cudaSetDevice(w->device));
  Store &s = w->s;
  const int32_t n = s.Nloc, ws = w->sh.ws;
  if (ws == 1) {  // a plain world: global id == slot
    w->sh.rank = 0;
    w->sh.gfirst[0] = 0;
    w->sh.gfirst[1] = n;
    w->Ntot = n;
  }
  const int32_t ntot = w->Ntot, g0 = w->sh.gfirst[w->sh.rank];
  if (mode != 2) w->tp = gbp_world::TopoPass();
  if (ntot == 0) return 0;
  cudaStream_t st = mode == 1 ? w->topo_stream : w->stream;
  if (mode != 2)
    if (int rc = ensure_topology_scratch(w)) return rc;
  const float *gx = ws > 1 ? w->gpos : s.pos;
  const float *gz = ws > 1 ? w->gpos + ntot : s.pos + s.cap;
  const float *ggone = ws > 1 ? w->gpos + 3 * int64_t(ntot) : s.gone;
  const int T = 128;
  const float R = w->cfg.comms_radius;
  const double cell = double(R) * 1.001;
  // entries in the hash: every robot on one GPU; on a shard only the robots near its own ones (k_cell_keys_near),
  // padded to cand_cap with a sentinel key pointing at a cell-less dummy
  int32_t nall = ntot;
  if (ws > 1) {
    if (w->cand_cap <= 0 || w->cand_cap > ntot) w->cand_cap = ntot;
    nall = int32_t(w->cand_cap);
  }
  size_t cb = w->t_cub_bytes;
  if (mode != 2) {
  if (ws == 1) {
    gbp::k_cell_keys<<<blocks_for(ntot, T), T, 0, st>>>(ntot, gx, gz, ggone, cell, w->t_cx, w->t_cz, w->t_keys, w->t_idx);
  } else {
    k_box_init<<<1, 32, 0, st>>>(w->t_box, w->t_cx + ntot, w->t_cz + ntot);
    if (n > 0) gbp::k_own_bbox<<<blocks_for(n, 256), 256, 0, st>>>(g0, n, gx, gz, ggone, w->t_box);
    gbp::k_fill_u32<<<blocks_for(nall, 256), 256, 0, st>>>(w->t_keys, w->t_idx, nall, 0xFFFFFFFFu, ntot);
    gbp::k_cell_keys_near<<<blocks_for(ntot, T), T, 0, st>>>(ntot, gx, gz, ggone, cell, R * 1.01f, w->t_box, w->t_cx,
                                                            w->t_cz, w->t_keys, w->t_idx, nall, w->t_box + 4);
    w->launches += 3;
  }
  cb = w->t_cub_bytes;
  CK(cub::DeviceRadixSort::SortPairs(w->t_cub, cb, w->t_keys, w->t_keys_sorted, w->t_idx, w->t_idx_sorted, nall, 0, 32, st));
  if (n > 0)
    gbp::k_neighbours_find<<<blocks_for(n, T), T, 0, st>>>(nall, g0, n, gx, gz, w->t_cx, w->t_cz, w->t_keys_sorted,
                                                           w->t_idx_sorted, R, w->t_cnt, w->t_park);
  CK(cudaMemsetAsync(w->t_cnt + n, 0, sizeof(int64_t), st));
  cb = w->t_cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_cnt, w->t_off, n + 1, st));
  w->launches += 4;
  }  // mode != 2
  // The new CSR is written straight into the spare edge set.  If the spare set is too small the
  // guarded kernels did nothing: grow it and run them again.
  EdgeSet *spare = &w->edges[1 - w->cur];
  const EdgeSet *live = &w->edges[w->cur];
  const bool quirk = w->cfg.strict_reference_quirks != 0;
  if (quirk && mode != 0) return fail(GBP_ERR_STATE, "strict_reference_quirks: the early topology search is not available");
  if (quirk && n > 0) {
    // delete_interrobot_factors as written: the robots within range become their own CSR (q_woff, q_wnbr); the new
    // rows are the old edges that survive the lossy deletion merged with fresh edges (k_quirk_rows).  A test mode:
    // two extra size read-backs per tick instead of the guarded relaunch scheme.
    int64_t nw = 0;
    CK(cudaMemcpyAsync(&nw, w->t_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (int rc = ensure_buf(w, w->q_wnbr, w->q_wcap, std::max<int64_t>(nw, 1))) return rc;
    if (n > w->q_ncap) {
      cudaFree(w->q_maxlost);
      cudaFree(w->q_woff);
      CK(dalloc(w->q_maxlost, size_t(n)));
      CK(dalloc(w->q_woff, size_t(n) + 1));
      w->q_ncap = n;
    }
    gbp::k_neighbours_fill<<<blocks_for(n, T), T, 0, st>>>(nall, g0, n, gx, gz, w->t_cx, w->t_cz, w->t_keys_sorted,
                                                           w->t_idx_sorted, R, w->t_off, w->t_park, w->q_wnbr, w->q_wcap);
    CK(cudaMemcpyAsync(w->q_woff, w->t_off, size_t(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    const int64_t *ooff = s.E > 0 ? s.eoff : nullptr;
    gbp::k_quirk_lost_max<<<blocks_for(n, T), T, 0, st>>>(n, ooff, live->egid, live->e_frozen, s.gone, w->q_woff,
                                                          w->q_wnbr, w->q_maxlost);
    gbp::k_quirk_rows<false><<<blocks_for(n, T), T, 0, st>>>(n, ooff, live->egid, live->e_frozen, s.gone, w->q_woff,
                                                             w->q_wnbr, w->q_maxlost, w->t_cnt, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr);
    CK(cudaMemsetAsync(w->t_cnt + n, 0, sizeof(int64_t), st));
    cb = w->t_cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_cnt, w->t_off, n + 1, st));
    int64_t e1 = 0;
    CK(cudaMemcpyAsync(&e1, w->t_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (e1 > spare->cap)
      if (int rc = grow_edge_set(w, spare, e1 + e1 / 4 + 1024)) return rc;
    if (int rc = ensure_buf(w, w->q_zombie, w->q_zcap, std::max<int64_t>(e1, 1))) return rc;
    gbp::k_quirk_rows<true><<<blocks_for(n, T), T, 0, st>>>(n, ooff, live->egid, live->e_frozen, s.gone, w->q_woff,
                                                            w->q_wnbr, w->q_maxlost, w->t_off, spare->egid, spare->map,
                                                            w->q_zombie, w->t_newcnt, w->t_nlow);
    CK(cudaGetLastError());
    w->launches += 6;
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool enqueue = !(mode == 2 && attempt == 0);  // mode 2: the first attempt was enqueued by mode 1
    if (enqueue) {
    if (n > 0 && !quirk) {
      gbp::k_neighbours_fill<<<blocks_for(n, T), T, 0, st>>>(nall, g0, n, gx, gz, w->t_cx, w->t_cz, w->t_keys_sorted,
                                                             w->t_idx_sorted, R, w->t_off, w->t_park, spare->egid,
                                                             spare->cap);
      gbp::k_edge_diff<<<blocks_for(n, T), T, 0, st>>>(n, g0, w->t_off, spare->egid, s.eoff, live->egid, n, spare->map,
                                                       w->t_newcnt, w->t_nlow, spare->cap);
    }
    CK(cudaMemsetAsync(w->t_newcnt + n, 0, sizeof(int64_t), st));
    cb = w->t_cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_newcnt, w->t_newoff, n + 1, st));
    w->launches += 3;
    if (ws > 1) {
      CK(cudaMemsetAsync(w->t_gflag, 0, size_t(ntot + 1) * sizeof(int32_t), st));
      CK(cudaMemsetAsync(w->t_sflag + int64_t(ws) * n, 0, sizeof(int32_t), st));
      CK(cudaMemsetAsync(w->t_ccnt + int64_t(ws) * n, 0, sizeof(int64_t), st));
      if (n > 0)
        gbp::k_cross_mark<<<blocks_for(n, T), T, 0, st>>>(w->sh, n, w->t_off, spare->egid, spare->cap, w->t_gflag,
                                                          w->t_sflag, w->t_ccnt);
      cb = w->t_cub_bytes;
      CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_gflag, w->t_gslot, ntot + 1, st));
      cb = w->t_cub_bytes;
      CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_sflag, w->t_soff, ws * n + 1, st));
      cb = w->t_cub_bytes;
      CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_ccnt, w->t_coff, ws * n + 1, st));
      w->launches += 4;
    }
    gbp::k_shard_result<<<1, 32, 0, st>>>(w->sh, n, w->t_off, w->t_newoff, ws > 1 ? w->t_gslot : nullptr, w->t_soff,
                                          w->t_coff, w->t_err, ws > 1 ? w->t_box + 4 : nullptr, w->t_result_dev);
    CK(cudaGetLastError());
    k_words_to_host<<<1, 64, 0, st>>>(w->t_result_host, w->t_result_dev, int(kResultWords));
    CK(cudaGetLastError());
    w->launches += 1;
    if (mode == 1) return 0;  // the sizes are read by the mode-2 call
    CK(cudaStreamSynchronize(st));
    w->launches += 1;
    }  // enqueue
    if (w->t_result_host[3] != 0)
      return fail(GBP_ERR_STATE, "sharded topology: a cross-shard robot_number lookup failed in an earlier pass "
                                 "(neighbour lists of two shards disagree)");
    if (ws > 1) {
      const int64_t ncand = w->t_result_host[4 + 3 * (gbp::kMaxShards + 1)];
      if (ncand > nall) {
        // more robots near this shard than the padded list holds: the lists just built miss some — size the list
        // for what was found and run the whole search again
        w->cand_cap = std::min<int64_t>(ntot, ncand + ncand / 4 + 1024);
        return topo_search(w);
      }
      // shrink toward the need (with slack) so that the sort stays proportional to the shard, not to the swarm
      w->cand_cap = std::min<int64_t>(ntot, std::max<int64_t>(ncand + ncand / 4 + 1024, 4096));
    }
    w->tp.E1 = w->t_result_host[0];
    w->tp.total_new = w->t_result_host[1];
    if (w->tp.E1 <= spare->cap) break;
    if (attempt == 1) return fail(GBP_ERR_STATE, "topology: spare edge set still too small after growing");
    if (int rc = grow_edge_set(w, spare, w->tp.E1 + w->tp.E1 / 4 + 1024)) return rc;
  }
  if (ws > 1) {
    w->tp.nghost = w->t_result_host[2];
    for (int q = 0; q <= ws; ++q) {
      w->tp.ghost_po.start[q] = w->t_result_host[4 + q];
      w->tp.send_po.start[q] = w->t_result_host[4 + (ws + 1) + q];
      w->tp.cross_po.start[q] = w->t_result_host[4 + 2 * (ws + 1) + q];
    }
  }
  // quirk mode: an edge can turn into a zombie without any edge appearing or disappearing — always rebuild
  w->tp.changed = w->tp.total_new != 0 || w->tp.E1 != s.E || w->force_rebuild || quirk;
  return 0;
}

// ---- topology, phase 2 (per shard, ws > 1): what the other shards need from this one: the
// header (new pair count, changed flag, epoch) and the numbers of the own cross-shard factors.
int topo_pack(gbp_world *w, gbp::XferPlan &plan) {
  CK(cudaSetDevice(w->device));
  Store &s = w->s;
  const int ws = w->sh.ws, n = s.Nloc;
  cudaStream_t st = w->stream;
  const int64_t ncross = w->tp.cross_po.start[ws];
  if (ncross > w->cross_cap) {
    CK(cudaStreamSynchronize(st));
    cudaFree(w->ckeys_s); cudaFree(w->cvals_s); cudaFree(w->ckeys_r); cudaFree(w->cvals_r);
    const int64_t cap = ncross + ncross / 4 + 1024;
    CK(dalloc(w->ckeys_s, cap)); CK(dalloc(w->cvals_s, cap)); CK(dalloc(w->ckeys_r, cap)); CK(dalloc(w->cvals_r, cap));
    w->cross_cap = cap;
  }
  int64_t *h = w->hdr_host + 4 * gbp::kMaxShards;  // staging row for the own header
  h[0] = w->tp.total_new;
  h[1] = w->tp.changed ? 1 : 0;
  h[2] = int64_t(w->epoch);
  h[3] = w->tp.E1;
  CK(cudaMemcpyAsync(w->hdr_send, h, 4 * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (n > 0 && ncross > 0) {
    const EdgeSet *spare = &w->edges[1 - w->cur], *live = &w->edges[w->cur];
    gbp::k_cross_pack<<<blocks_for(n, 128), 128, 0, st>>>(w->sh, n, w->t_off, spare->egid, spare->map, w->t_newoff,
                                                        live->e_own, w->t_coff, w->ckeys_s, w->cvals_s);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  for (int q = 0; q < ws; ++q) {
    if (q == w->sh.rank) continue;
    const int64_t c0 = w->tp.cross_po.start[q], cnt = w->tp.cross_po.start[q + 1] - c0;
    plan.sends.push_back({q, w->hdr_send, 4 * sizeof(int64_t)});
    plan.recvs.push_back({q, w->hdr_recv + 4 * q, 4 * sizeof(int64_t)});
    if (cnt) {
      plan.sends.push_back({q, w->ckeys_s + c0, size_t(cnt) * 8});
      plan.sends.push_back({q, w->cvals_s + c0, size_t(cnt) * 8});
      plan.recvs.push_back({q, w->ckeys_r + c0, size_t(cnt) * 8});
      plan.recvs.push_back({q, w->cvals_r + c0, size_t(cnt) * 8});
    }
  }
  return 0;
}

// ---- topology, phase 3 (per shard): apply the new CSR.  hdr[q] = {new pairs, changed, epoch, E}
// of every shard (own row included).
int topo_apply(gbp_world *w, const int64_t (*hdr)[4]) {
  CK(cudaSetDevice(w->device));
  Store &s = w->s;
  const int ws = w->sh.ws, n = s.Nloc, rank = w->sh.rank;
  bool changed = false;
  int64_t sum_new = 0, max_epoch = 0;
  gbp::ShardBases bases{};
  for (int q = 0; q < ws; ++q) {
    bases.base[q] = sum_new;
    sum_new += hdr[q][0];
    changed = changed || hdr[q][1] != 0;
    max_epoch = std::max(max_epoch, hdr[q][2]);
  }
  if (!changed) return 0;  // connectivity unchanged everywhere: keep the store as is
  cudaStream_t st = w->stream;
  const int T = 128;
  const int32_t ntot = w->Ntot, g0 = w->sh.gfirst[rank];
  const int64_t E1 = w->tp.E1;
  // every shard moves to the same epoch, above anything any shard has written so far
  w->epoch = uint32_t(max_epoch) + 1;
  const int Vm1 = s.V - 1;
  if (ws > 1) {
    const int64_t hd = gbp::halo_doubles_per_robot(s.V);
    if (int64_t(n) + w->tp.nghost > s.cap) {
      if (int rc = reserve_robots(w, int64_t(n) + w->tp.nghost + w->tp.nghost / 4 + 256)) return rc;
    }
    if (int rc = ensure_buf(w, w->sendlist, w->sendlist_cap, w->tp.send_po.start[ws])) return rc;
    if (int rc = ensure_buf(w, w->halo_send, w->halo_send_cap, w->tp.send_po.start[ws] * hd)) return rc;
    if (int rc = ensure_buf(w, w->halo_recv, w->halo_recv_cap, w->tp.nghost * hd)) return rc;
  }
  EdgeSet *spare = &w->edges[1 - w->cur];
  const EdgeSet *live = &w->edges[w->cur];
  const float *gradius = ws > 1 ? w->gpos + 2 * int64_t(ntot) : s.radius;
  if (n > 0) {
    gbp::k_edge_assign_own<<<blocks_for(n, T), T, 0, st>>>(
        n, s.V, w->t_off, spare->egid, spare->map, w->t_newoff, gradius, double(w->cfg.safety_distance_multiplier),
        w->robot_number, bases.base[rank], w->epoch, live->e_own, live->e_rnum, live->e_birth, live->e_frozen,
        spare->e_own, spare->e_dsafe, spare->e_rnum, spare->e_birth, spare->e_frozen,
        w->cfg.strict_reference_quirks ? w->q_zombie : nullptr);
    gbp::k_edge_pull<<<blocks_for(n, T), T, 0, st>>>(w->sh, n, s.V, w->t_off, spare->egid, spare->map, spare->e_own,
                                                     w->robot_number, bases, w->tp.cross_po, w->ckeys_r, w->cvals_r,
                                                     spare->e_rnum, w->t_err);
    w->launches += 2;
  }
  if (w->count_messages && n > 0) {
    k_msg_edges<<<blocks_for(n, T), T, 0, st>>>(s, n, w->t_off, spare->map, live->cap, live->e_msg, spare->cap,
                                                spare->e_msg, w->msg_cap, w->msg_cnt);
    w->launches += 1;
  }
  if (E1 > 0) {
    gbp::k_mirror_move<<<blocks_for(E1 * Vm1, 256), 256, 0, st>>>(s, w->p, E1 * Vm1, Vm1, w->t_off, spare->map, s.mir,
                                                                  s.mu_frozen, s.EV, spare->mir, spare->mu_frozen,
                                                                  spare->cap * Vm1);
    w->launches += 1;
  }
  if (ws > 1) {
    gbp::k_ghost_fill<<<blocks_for(ntot, 256), 256, 0, st>>>(ntot, n, w->t_gflag, w->t_gslot, gradius, s.gid, s.radius);
    if (E1 > 0)
      gbp::k_edge_slots<<<blocks_for(E1, 256), 256, 0, st>>>(E1, g0, n, spare->egid, w->t_gslot, spare->enbr);
    if (n > 0)
      gbp::k_sendlist_fill<<<blocks_for(int64_t(ws) * n, 256), 256, 0, st>>>(ws, n, w->t_sflag, w->t_soff, w->sendlist);
    w->launches += 3;
    // border robots = the union of the send lists (a robot next to two peers is listed twice)
    const int64_t nsend = w->tp.send_po.start[ws];
    if (int rc = ensure_buf(w, w->border_list, w->border_cap, std::max<int64_t>(nsend, 1))) return rc;
    if (int rc = ensure_buf(w, w->border_words, w->border_words_cap, (int64_t(s.cap) + 3) / 4 + 1)) return rc;
    if (!w->border_count) CK(dalloc(w->border_count, 1));
    if (!w->border_gen_count) {
      CK(dalloc(w->border_gen_count, 2));
      CK(cudaMemsetAsync(w->border_gen_count, 0, 2 * sizeof(int32_t), st));
    }
    if (int rc = ensure_buf(w, w->border_gen_list, w->border_gen_cap, std::max<int64_t>(nsend, 1))) return rc;
    CK(cudaMemsetAsync(w->border_words, 0, size_t(w->border_words_cap) * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(w->border_count, 0, sizeof(int32_t), st));
    if (nsend > 0) {
      gbp::k_border_list<<<blocks_for(nsend, 256), 256, 0, st>>>(nsend, w->sendlist, w->border_words, w->border_list,
                                                               w->border_count);
      w->launches += 1;
    }
    w->n_border_max = nsend;
  }
  CK(cudaGetLastError());
  if (n > 0) {
    CK(cudaMemcpyAsync(s.eoff, w->t_off, size_t(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(s.nlow, w->t_nlow, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  w->cur = 1 - w->cur;
  bind_edge_set(w);
  s.E = E1;
  w->robot_number += uint64_t(Vm1) * uint64_t(sum_new);
  w->nghost = int32_t(w->tp.nghost);
  s.N = n + w->nghost;
  w->ghost_po = w->tp.ghost_po;
  w->send_po = w->tp.send_po;
  w->force_rebuild = false;
  mark_halo_stale(w);
  return refresh_ell(w);
}

// Every shard learns every robot's Transform (x, z), radius and despawned flag: 16 bytes per robot per tick.
// `early`: on the topo streams (group_topology_early), else on the shards' own streams.
int group_exchange_positions(gbp_group *g, bool early) {
  const int ws = g->ws;
  const size_t nm = g->members.size();
  std::vector<gbp::XferPlan> plans(nm);
  for (size_t m = 0; m < nm; ++m) {
    gbp_world *w = g->members[m];
    CK(cudaSetDevice(w->device));
    Store &s = w->s;
    cudaStream_t st = early ? w->topo_stream : w->stream;
    const int rank = w->sh.rank, ntot = w->Ntot, n = s.Nloc, g0 = w->sh.gfirst[rank];
    const float *src[4] = {s.pos, s.pos + s.cap, s.radius, s.gone};
    for (int k = 0; k < 4; ++k) {
      if (n > 0)
        CK(cudaMemcpyAsync(w->gpos + int64_t(k) * ntot + g0, src[k], size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
      for (int q = 0; q < ws; ++q) {
        if (q == rank) continue;
        const int nq = w->sh.gfirst[q + 1] - w->sh.gfirst[q];
        if (n > 0) plans[m].sends.push_back({q, const_cast<float *>(src[k]), size_t(n) * 4});
        if (nq > 0) plans[m].recvs.push_back({q, w->gpos + int64_t(k) * ntot + w->sh.gfirst[q], size_t(nq) * 4});
      }
    }
  }
  // sends are grouped by plane then peer on both sides, so the k-th send to a peer pairs with its k-th receive
  return exchange(g, plans, false, early);
}

// Host and the shards' own streams wait for a search that group_topology_early started (if any).
int group_topology_early_wait(gbp_group *g) {
  if (!g->early_pending) return 0;
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    CK(cudaEventSynchronize(w->ev_topo));
    if (w->topo_stream != w->stream) CK(cudaStreamWaitEvent(w->stream, w->ev_topo, 0));
  }
  g->early_pending = false;
  return 0;
}

// The next tick's neighbour search, started as soon as the positions it reads are final (the end of
// update_prior_of_current_state): it runs on the topo streams next to iterate_gbp, which neither reads what it writes
// (scratch, the spare edge set) nor writes what it reads (positions, the live edge lists).
int group_topology_early(gbp_group *g) {
  gbp_world *w0 = g->members[0];
  if (!w0->topo_early || w0->cfg.strict_reference_quirks || (g->ws > 1 && !g->committed)) return 0;
  if (int rc = group_topology_early_wait(g)) return rc;  // an earlier one nobody used
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    if (w->topo_stream != w->stream) {
      CK(cudaEventRecord(w->ev_prior, w->stream));
      CK(cudaStreamWaitEvent(w->topo_stream, w->ev_prior, 0));
    }
  }
  if (g->ws > 1) {
    ProfileScope pp(w0, GBP_PROFILE_TOPO_POSITIONS, w0->topo_stream);
    if (int rc = group_exchange_positions(g, true)) return rc;
  }
  {
    ProfileScope pq(w0, GBP_PROFILE_TOPO_SEARCH, w0->topo_stream);
    for (gbp_world *w : g->members)
      if (int rc = topo_search(w, 1)) return rc;
  }
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    CK(cudaEventRecord(w->ev_topo, w->topo_stream));
  }
  g->early_pending = true;
  g->early_version = g->topo_version;
  return 0;
}

// update_robot_neighbours + delete_interrobot_factors + create_interrobot_factors for every
// shard of the group (robot.rs:1362-1586).
int group_update_topology(gbp_group *g) {
  const int ws = g->ws;
  if (ws > 1 && !g->committed) return fail(GBP_ERR_STATE, "sharded world: call gbp_world_commit_shards first");
  if (int rc = group_halo_join(g)) return rc;
  ProfileScope ps(g->members[0], GBP_PROFILE_TOPOLOGY);
  const size_t nm = g->members.size();
  std::vector<gbp::XferPlan> plans(nm);
  // a search started early is used if nothing it read has changed since; either way it is waited for first
  const bool early = g->early_pending && g->early_version == g->topo_version;
  if (int rc = group_topology_early_wait(g)) return rc;
  if (ws > 1 && !early) {
    ProfileScope pp(g->members[0], GBP_PROFILE_TOPO_POSITIONS);
    if (int rc = group_exchange_positions(g, false)) return rc;
  }
  if (early) {  // timed where it ran (group_topology_early); here the host only reads its sizes
    for (gbp_world *w : g->members)
      if (int rc = topo_search(w, 2)) return rc;
  } else {
    ProfileScope pq(g->members[0], GBP_PROFILE_TOPO_SEARCH);
    for (gbp_world *w : g->members)
      if (int rc = topo_search(w, 0)) return rc;
  }
  ProfileScope pa(g->members[0], GBP_PROFILE_TOPO_APPLY);
  int64_t hdr[gbp::kMaxShards][4];
  if (ws == 1) {
    gbp_world *w = g->members[0];
    hdr[0][0] = w->tp.total_new;
    hdr[0][1] = w->tp.changed ? 1 : 0;
    hdr[0][2] = int64_t(w->epoch);
    hdr[0][3] = w->tp.E1;
    return topo_apply(w, hdr);
  }
  for (size_t m = 0; m < nm; ++m) {
    plans[m].clear();
    if (int rc = topo_pack(g->members[m], plans[m])) return rc;
  }
  if (int rc = exchange(g, plans)) return rc;
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    k_words_to_host<<<1, 64, 0, w->stream>>>(w->hdr_host, w->hdr_recv, 4 * ws);
    CK(cudaGetLastError());
    w->launches += 1;
    CK(cudaStreamSynchronize(w->stream));
  }
  for (gbp_world *w : g->members) {
    for (int q = 0; q < ws; ++q)
      for (int k = 0; k < 4; ++k)
        hdr[q][k] = q == w->sh.rank ? w->hdr_host[4 * gbp::kMaxShards + k] : w->hdr_host[4 * q + k];
    if (int rc = topo_apply(w, hdr)) return rc;
  }
  return 0;
}

// Fix the global robot ids: shard q owns [gfirst[q], gfirst[q+1]) in rank order.
int group_commit(gbp_group *g) {
  const int ws = g->ws;
  const size_t nm = g->members.size();
  std::vector<gbp::XferPlan> plans(nm);
  if (int rc = topology_inputs_change(g->members[0])) return rc;
  for (size_t m = 0; m < nm; ++m) {
    gbp_world *w = g->members[m];
    CK(cudaSetDevice(w->device));
    if (int rc = ensure_topology_scratch(w)) return rc;
    int64_t *h = w->hdr_host + 4 * gbp::kMaxShards;
    h[0] = w->s.Nloc;
    h[1] = h[2] = h[3] = 0;
    CK(cudaMemcpyAsync(w->hdr_send, h, 4 * sizeof(int64_t), cudaMemcpyHostToDevice, w->stream));
    for (int q = 0; q < ws; ++q) {
      if (q == w->sh.rank) continue;
      plans[m].sends.push_back({q, w->hdr_send, 4 * sizeof(int64_t)});
      plans[m].recvs.push_back({q, w->hdr_recv + 4 * q, 4 * sizeof(int64_t)});
    }
  }
  if (int rc = exchange(g, plans)) return rc;
  for (gbp_world *w : g->members) {
    CK(cudaSetDevice(w->device));
    CK(cudaMemcpyAsync(w->hdr_host, w->hdr_recv, 4 * size_t(gbp::kMaxShards) * sizeof(int64_t), cudaMemcpyDeviceToHost,
                       w->stream));
    CK(cudaStreamSynchronize(w->stream));
    int64_t tot = 0;
    for (int q = 0; q < ws; ++q) {
      w->sh.gfirst[q] = int32_t(tot);
      tot += q == w->sh.rank ? int64_t(w->s.Nloc) : w->hdr_host[4 * q];
    }
    if (tot > INT32_MAX) return fail(GBP_ERR_BAD_ARGUMENT, "more than 2^31 robots");
    w->sh.gfirst[ws] = int32_t(tot);
    w->Ntot = int32_t(tot);
    if (ws > 1 && 4 * tot > w->gpos_cap) {
      cudaFree(w->gpos);
      CK(dalloc(w->gpos, size_t(4 * tot)));
      w->gpos_cap = 4 * tot;
    }
    if (w->s.Nloc > 0) {
      k_iota_gid<<<blocks_for(w->s.Nloc, 256), 256, 0, w->stream>>>(w->s.gid, w->sh.gfirst[w->sh.rank], w->s.Nloc);
      CK(cudaGetLastError());
      w->launches += 1;
    }
    w->force_rebuild = true;
  }
  g->committed = true;
  g->halo_stale = true;
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" {

const char *gbp_last_error(void) { return g_err.c_str(); }

int gbp_schedule(int32_t kind, uint8_t internal, uint8_t external, uint8_t *oi, uint8_t *oe) {
  if (kind < 0 || kind > 4 || !oi || !oe) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_schedule: bad argument");
  const int max = std::max(internal, external);
  schedule_half(kind, internal, max, oi);
  schedule_half(kind, external, max, oe);
  return max;
}

int gbp_variable_timesteps(uint32_t h, uint32_t m, uint32_t *out, int32_t cap) {
  if (!out || m == 0) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_variable_timesteps: bad argument");
  // utils.rs:35-75, f32 arithmetic; mul_add is a fused multiply-add
  const uint32_t n = 1u + uint32_t(0.5f * (-1.0f + sqrtf(1.0f + 8.0f * float(h) / float(m))));
  int cnt = 0;
  for (uint32_t i = 0; i < m * (n + 1); ++i) {
    const uint32_t section = i / m;
    const float f = fmaf(float(m) / 2.0f, float(section), fmaf(float(section), -float(m), float(i))) *
                    (float(section) + 1.0f);
    if (cnt >= cap) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_variable_timesteps: capacity too small");
    if (f >= float(h)) {
      out[cnt++] = h;
      break;
    }
    out[cnt++] = uint32_t(f);
  }
  return cnt;
}

namespace {
gbp_world *make_world(const gbp_config_t *cfg, int32_t device, cudaStream_t shared) {
  if (!cfg || cfg->num_variables < 2 || cfg->num_variables > 32) {
    fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_create: num_variables must be in [2, 32]");
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    fail(GBP_ERR_CUDA, "gbp_world_create: no usable CUDA device (this engine has no CPU path)");
    return nullptr;
  }
  gbp_world *w = new gbp_world();
  w->cfg = *cfg;
  w->device = device;
  bool ok = cudaSetDevice(device) == cudaSuccess;
  if (ok && shared) {
    w->stream = shared;
    w->owns_stream = false;
  } else if (ok) {
    ok = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) == cudaSuccess;
  }
  if (!ok || cudaEventCreate(&w->ev0) != cudaSuccess || cudaEventCreate(&w->ev1) != cudaSuccess) {
    fail(GBP_ERR_CUDA, "gbp_world_create: stream/event creation failed");
    delete w;
    return nullptr;
  }
  w->s.V = cfg->num_variables;
  cudaDeviceGetAttribute(&w->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (const char *e = std::getenv("GBP_GENERAL_ONLY")) w->general_only = e[0] == '1';
  if (const char *e = std::getenv("GBP_PDL")) w->use_pdl = e[0] != '0';
  if (const char *e = std::getenv("GBP_HALO_OVERLAP")) w->halo_overlap = e[0] != '0';
  // default SDF: a single white pixel (empty environment)
  const uint8_t white = 255;
  if (cudaMalloc(&w->sdf_dev, 1) != cudaSuccess ||
      cudaMemcpy(w->sdf_dev, &white, 1, cudaMemcpyHostToDevice) != cudaSuccess) {
    fail(GBP_ERR_CUDA, "gbp_world_create: sdf alloc failed");
    delete w;
    return nullptr;
  }
  w->s.sdf = w->sdf_dev;
  w->s.sdf_w = 1;
  w->s.sdf_h = 1;
  w->sh.ws = 1;
  w->sh.rank = 0;
  refresh_scalars(w);
  return w;
}

void join_group(gbp_world *w, gbp_group *g, int rank) {
  // NCCL groups: a second stream carries every transfer; in-process shards share one stream for everything
  w->comm_stream = w->stream;
  if (g->nccl) {
    cudaSetDevice(w->device);
    // Highest priority, like the border stream: the pack / send-recv / unpack kernels of a halo exchange are a few
    // CTAs that must slip in between the 15 000 CTAs of the interior launch queued before them; at the default
    // priority they waited for that grid to drain and the exchange landed on the critical path after all
    // (r02z8: ~80 us of every 330 us sub-step at 8 GPUs).  GBP_COMM_PRIORITY=0: default priority.
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char *e = std::getenv("GBP_COMM_PRIORITY");
    if (e && e[0] == '0') cudaStreamCreateWithFlags(&w->comm_stream, cudaStreamNonBlocking);
    else cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, hi);
  }
  cudaEventCreateWithFlags(&w->ev_border, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&w->ev_halo, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&w->ev_fence, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&w->ev_start, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&w->ev_prior, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&w->ev_topo, cudaEventDisableTiming);
  w->topo_stream = w->stream;
  if (g->nccl || g->ws == 1) {
    cudaSetDevice(w->device);
    cudaStreamCreateWithFlags(&w->topo_stream, cudaStreamNonBlocking);
  }
  if (const char *e = std::getenv("GBP_TOPO_EARLY")) w->topo_early = e[0] != '0';
  w->border_stream = w->stream;
  if (g->nccl) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
    cudaStreamCreateWithPriority(&w->border_stream, cudaStreamNonBlocking, hi);
  }
  w->grp = g;
  w->sh.ws = g->ws;
  w->sh.rank = rank;
  g->members.push_back(w);
  g->refs += 1;
}
}  // namespace

gbp_world_t *gbp_world_create(const gbp_config_t *cfg, int32_t device) {
  gbp_world *w = make_world(cfg, device, nullptr);
  if (!w) return nullptr;
  gbp_group *g = new gbp_group();
  g->committed = true;
  join_group(w, g, 0);
  return w;
}

int gbp_comm_unique_id(uint8_t *id) {
  if (!id) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_comm_unique_id: null output");
  gbp::NcclApi &api = gbp::nccl_api();
  if (!api.load()) return fail(GBP_ERR_NCCL, api.error);
  gbp::NcclApi::unique_id uid;
  const int rc = api.GetUniqueId(&uid);
  if (rc) return fail(GBP_ERR_NCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(rc));
  static_assert(sizeof(uid) == GBP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(id, &uid, sizeof(uid));
  return 0;
}

gbp_world_t *gbp_world_create_shard(const gbp_config_t *cfg, int32_t device, int32_t rank, int32_t world_size,
                                    const uint8_t *id) {
  if (world_size < 1 || world_size > gbp::kMaxShards || rank < 0 || rank >= world_size || (world_size > 1 && !id)) {
    fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_create_shard: need 0 <= rank < world_size <= 16 and a communicator id");
    return nullptr;
  }
  if (world_size == 1) return gbp_world_create(cfg, device);
  if (cfg && cfg->strict_reference_quirks) {
    fail(GBP_ERR_BAD_ARGUMENT, "strict_reference_quirks is available on single-GPU worlds only (a pair is deleted or kept "
                               "depending on the lost neighbours of BOTH robots)");
    return nullptr;
  }
  gbp::NcclApi &api = gbp::nccl_api();
  if (!api.load()) {
    fail(GBP_ERR_NCCL, api.error);
    return nullptr;
  }
  gbp_world *w = make_world(cfg, device, nullptr);
  if (!w) return nullptr;
  gbp_group *g = new gbp_group();
  g->ws = world_size;
  g->nccl = true;
  join_group(w, g, rank);
  gbp::NcclApi::unique_id uid;
  std::memcpy(&uid, id, sizeof(uid));
  const int rc = api.CommInitRank(&g->comm, world_size, uid, rank);
  if (rc) {
    fail(GBP_ERR_NCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(rc));
    gbp_world_destroy(w);
    return nullptr;
  }
  return w;
}

int gbp_world_create_local_shards(const gbp_config_t *cfg, int32_t device, int32_t world_size, gbp_world_t **out) {
  if (!out || world_size < 1 || world_size > gbp::kMaxShards)
    return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_create_local_shards: need 1 <= world_size <= 16");
  if (cfg && cfg->strict_reference_quirks && world_size > 1)
    return fail(GBP_ERR_BAD_ARGUMENT, "strict_reference_quirks is available on single-GPU worlds only");
  gbp_group *g = new gbp_group();
  g->ws = world_size;
  g->committed = world_size == 1;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g->shared_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete g;
    return fail(GBP_ERR_CUDA, "gbp_world_create_local_shards: no usable CUDA device (this engine has no CPU path)");
  }
  for (int r = 0; r < world_size; ++r) {
    gbp_world *w = make_world(cfg, device, g->shared_stream);
    if (!w) {
      for (int k = 0; k < r; ++k) gbp_world_destroy(out[k]);
      if (r == 0) {
        cudaStreamDestroy(g->shared_stream);
        delete g;
      }
      return GBP_ERR_CUDA;
    }
    join_group(w, g, r);
    out[r] = w;
  }
  return 0;
}

namespace {
void free_track_rings(gbp_world *w);
}
void gbp_world_destroy(gbp_world_t *w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->stream);
  Store &s = w->s;
  free_edge_set(w, &w->edges[0]);
  free_edge_set(w, &w->edges[1]);
  cudaFree(w->dyn_tab_dev);
  if (w->comm_stream && w->comm_stream != w->stream) {
    cudaStreamSynchronize(w->comm_stream);
    cudaStreamDestroy(w->comm_stream);
  }
  if (w->border_stream && w->border_stream != w->stream) {
    cudaStreamSynchronize(w->border_stream);
    cudaStreamDestroy(w->border_stream);
  }
  if (w->topo_stream && w->topo_stream != w->stream) {
    cudaStreamSynchronize(w->topo_stream);
    cudaStreamDestroy(w->topo_stream);
  }
  for (cudaEvent_t e : {w->ev_prior, w->ev_topo})
    if (e) cudaEventDestroy(e);
  cudaFree(w->env_cols);
  cudaFree(w->env_verts);
  cudaFree(w->env_state);
  cudaFree(w->env_hits);
  cudaFree(w->env_totals);
  free_track_rings(w);
  cudaFree(w->msg_cnt);
  cudaFree(w->border_gen_list);
  cudaFree(w->border_gen_count);
  for (cudaEvent_t e : {w->ev_border, w->ev_halo, w->ev_fence, w->ev_start})
    if (e) cudaEventDestroy(e);
  cudaFree(w->border_list);
  cudaFree(w->border_words);
  cudaFree(w->border_count);
  void *ptrs[] = {s.prior_eta, s.prior_lam, s.pub[0], s.pub[1], s.pub_epoch[0], s.pub_epoch[1], s.bel_ext,
                  s.mu_ext, s.cov, s.valid, s.cov_lazy, s.mode, s.gen_list, s.gen_count, s.m_dynL[0], s.m_dynL[1], s.m_dynR[0], s.m_dynR[1], s.m_obs, s.m_trk, s.dyn_c,
                  s.trk_record, s.trk_timeout, s.trk_seed, s.trk_last, s.trk_value, s.radius, s.t0, s.pos, s.antenna,
                  s.idle, s.finished, s.gone, s.latest, s.iter_factor, s.gid, s.next_wp, s.coll_hits, s.ell_nbr, s.ell_birth, s.ell_dsafe, w->coll_totals, s.wp_off, s.wp_xy, s.eoff,
                  s.nlow, w->t_nlow, w->t_result_dev, w->sdf_dev, w->t_cx,
                  w->t_cz, w->t_box, w->t_park, w->q_wnbr, w->q_maxlost, w->q_woff, w->q_zombie, w->t_idx, w->t_idx_sorted, w->t_keys, w->t_keys_sorted, w->t_cnt, w->t_off,
                  w->t_newcnt, w->t_newoff, w->t_cub, w->rb_dev, w->gpos, w->t_gflag, w->t_gslot, w->t_sflag,
                  w->t_soff, w->t_ccnt, w->t_coff, w->t_err, w->sendlist, w->ckeys_s, w->cvals_s, w->ckeys_r,
                  w->cvals_r, w->hdr_send, w->hdr_recv, w->halo_send, w->halo_recv};
  for (void *q : ptrs) cudaFree(q);
  for (auto &sp : w->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (cudaEvent_t e : w->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(w->ev0);
  cudaEventDestroy(w->ev1);
  if (w->up_stream) {
    cudaStreamSynchronize(w->up_stream);
    cudaStreamDestroy(w->up_stream);
    for (int q = 0; q < gbp_world::kUpKinds; ++q)
      for (int g = 0; g < 2; ++g) {
        cudaFree(w->up_stage[q][g]);
        if (w->ev_up_done[q][g]) cudaEventDestroy(w->ev_up_done[q][g]);
        if (w->ev_up_used[q][g]) cudaEventDestroy(w->ev_up_used[q][g]);
      }
  }
  if (w->copy_stream) {
    cudaStreamSynchronize(w->copy_stream);
    cudaStreamDestroy(w->copy_stream);
    cudaEventDestroy(w->ev_gathered);
    cudaEventDestroy(w->ev_copied);
  }
  if (w->owns_stream) cudaStreamDestroy(w->stream);
  if (w->t_result_host) cudaFreeHost(w->t_result_host);
  if (w->hdr_host) cudaFreeHost(w->hdr_host);
  if (gbp_group *g = w->grp) {
    g->members.erase(std::remove(g->members.begin(), g->members.end(), w), g->members.end());
    if (--g->refs == 0) {
      if (g->comm) gbp::nccl_api().CommDestroy(g->comm);
      if (g->shared_stream) cudaStreamDestroy(g->shared_stream);
      delete g;
    }
  }
  delete w;
}

int gbp_world_commit_shards(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  gbp_group *g = w->grp;
  if (g->ws == 1) return 0;
  if (!g->nccl && int(g->members.size()) != g->ws)
    return fail(GBP_ERR_STATE, "local shard group: a member has been destroyed");
  return group_commit(g);
}

int64_t gbp_world_first_global_id(const gbp_world_t *w) { return w ? w->sh.gfirst[w->sh.rank] : 0; }
int64_t gbp_world_num_robots_global(const gbp_world_t *w) {
  if (!w) return 0;
  return w->sh.ws == 1 ? w->s.Nloc : w->Ntot;
}
int32_t gbp_world_num_ghosts(const gbp_world_t *w) { return w ? w->nghost : 0; }

int gbp_world_set_sdf(gbp_world_t *w, const uint8_t *rgb8, int32_t width, int32_t height) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!rgb8 || width <= 0 || height <= 0) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_set_sdf: bad image");
  if (set_device(w)) return GBP_ERR_CUDA;
  // only the red channel is read (obstacle.rs:178)
  std::vector<uint8_t> red(size_t(width) * height);
  for (size_t k = 0; k < red.size(); ++k) red[k] = rgb8[3 * k];
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(w->sdf_dev);
  CK(cudaMalloc(&w->sdf_dev, red.size()));
  CK(cudaMemcpy(w->sdf_dev, red.data(), red.size(), cudaMemcpyHostToDevice));
  w->s.sdf = w->sdf_dev;
  w->s.sdf_w = width;
  w->s.sdf_h = height;
  refresh_scalars(w);
  return 0;
}

namespace {
// image 0.25.1 `imageops::sample::gaussian` in f32 (host side: the tap weights depend on sigma and the
// integer tap offset only; the kernels normalise them per window)
float env_gaussian(float x, float r) {
  return (1.0f / (std::sqrt(2.0f * 3.14159265358979323846f) * r)) * std::exp(-(x * x) / (2.0f * (r * r)));
}

// Host-side preparation of the placeable obstacles: `shape.expanded(expansion as f64)`, the vertices each
// shape's `inside()` would recompute per point, and the rotation quaternion of is_placeable_obstacle
// (env_to_png/src/lib.rs:303-324).  libm sinf / cosf / sin / cos, as the reference process calls them.
int env_prepare_shapes(const gbp_environment_t *env, std::vector<gbp::EnvShape> &shapes, std::vector<double> &verts) {
  const float kPi = 3.14159265358979323846f, kHalfPi = 1.57079632679489661923f;
  const double expansion = double(env->expansion);
  for (int k = 0; k < env->n_obstacles; ++k) {
    const gbp_obstacle_t &o = env->obstacles[k];
    gbp::EnvShape s{};
    s.kind = o.kind;
    s.tile_row = o.tile_row;
    s.tile_col = o.tile_col;
    s.tx = float(o.tx);
    s.ty = float(o.ty);
    float rotation_offset = kHalfPi;
    switch (o.kind) {
      case GBP_SHAPE_CIRCLE: {  // Circle::expanded / inside (gbp_environment/src/lib.rs:128-141)
        const double r = o.radius + expansion;
        s.r2 = float(r * r);
        break;
      }
      case GBP_SHAPE_TRIANGLE: {  // Triangle::expanded / points (:171-210), f32
        const float radius = float(o.radius + expansion);
        const float a = float(o.angle_a), b = float(o.angle_b);
        const float c = kPi - (a + b);
        const float ah = radius / std::sin(a), bh = radius / std::sin(b), ch = radius / std::sin(c);
        const float aa = kPi + a / 2.0f, ba = -b / 2.0f, ca = kPi - b - c / 2.0f;
        s.tri[0] = std::cos(aa) * ah;
        s.tri[1] = std::sin(aa) * ah;
        s.tri[2] = std::cos(ba) * bh;
        s.tri[3] = std::sin(ba) * bh;
        s.tri[4] = std::cos(ca) * ch;
        s.tri[5] = std::sin(ca) * ch;
        break;
      }
      case GBP_SHAPE_REGULAR_POLYGON: {  // RegularPolygon::expanded / point_at (:250-287), f64
        if (o.sides < 1) return fail(GBP_ERR_BAD_ARGUMENT, "environment: regular polygon without sides");
        const double radius = o.radius + expansion * 2.0;
        s.n = o.sides;
        s.poff = int64_t(verts.size() / 2);
        for (int i = 0; i < o.sides; ++i) {
          const double angle = 2.0 * 3.14159265358979323846 / double(o.sides) * double(i) + 0.78539816339744830962;
          verts.push_back(std::cos(angle) * radius);
          verts.push_back(std::sin(angle) * radius);
        }
        rotation_offset = kHalfPi + kHalfPi + ((o.sides % 2 != 0) ? kPi / float(o.sides) : 0.0f);
        break;
      }
      case GBP_SHAPE_POLYGON: {  // Polygon::expanded (:374-401)
        if (o.n_points < 1 || !env->polygon_points)
          return fail(GBP_ERR_BAD_ARGUMENT, "environment: polygon without points");
        const double *p = env->polygon_points + 2 * o.point_offset;
        double ax = 0.0, ay = 0.0;
        for (int i = 0; i < o.n_points; ++i) {
          ax = ax + p[2 * i];
          ay = ay + p[2 * i + 1];
        }
        const double cx = ax / double(o.n_points), cy = ay / double(o.n_points);
        s.n = o.n_points;
        s.poff = int64_t(verts.size() / 2);
        for (int i = 0; i < o.n_points; ++i) {
          const double dx = p[2 * i] - cx, dy = p[2 * i + 1] - cy;
          verts.push_back(p[2 * i] + dx * 4.0 * expansion);
          verts.push_back(p[2 * i + 1] + dy * 4.0 * expansion);
        }
        rotation_offset = 0.0f;
        break;
      }
      case GBP_SHAPE_RECTANGLE: {  // Rectangle::expanded / inside (:335-359)
        s.hw = (o.width + expansion * 2.0) / 4.0;
        s.hh = (o.height + expansion * 2.0) / 4.0;
        break;
      }
      default: return fail(GBP_ERR_BAD_ARGUMENT, "environment: unknown obstacle shape");
    }
    const float angle = float(o.rotation) + rotation_offset;
    s.qs = std::sin(angle * 0.5f);
    s.qw = std::cos(angle * 0.5f);
    shapes.push_back(s);
  }
  return 0;
}

// env_to_png::env_to_sdf_image (crates/env_to_png/src/lib.rs:149-163) on the current device: leaves the
// single-channel image in *d_gray (cudaMalloc'ed, caller owns it).
int env_to_sdf_device(const gbp_environment_t *env, cudaStream_t st, uint8_t **d_gray, uint32_t *Wo, uint32_t *Ho,
                      int64_t *launches) {
  if (!env || !env->tiles || env->nrows <= 0 || env->ncols <= 0 || env->resolution == 0)
    return fail(GBP_ERR_BAD_ARGUMENT, "environment: empty tile grid or zero resolution");
  // Percentage::new asserts (env_to_png/src/lib.rs:53-57, Sub/Add :126-142)
  const float pw = env->path_width - env->expansion;
  if (!(env->path_width >= 0.0f && env->path_width <= 1.0f && env->expansion >= 0.0f && env->expansion <= 1.0f &&
        env->blur >= 0.0f && env->blur <= 1.0f && pw >= 0.0f))
    return fail(GBP_ERR_BAD_ARGUMENT, "environment: path-width, expansion and blur are percentages in [0, 1], "
                                      "path-width >= expansion");
  if (env->n_obstacles < 0 || (env->n_obstacles > 0 && !env->obstacles))
    return fail(GBP_ERR_BAD_ARGUMENT, "environment: bad obstacle list");
  std::vector<gbp::EnvShape> shapes;
  std::vector<double> verts;
  if (int rc = env_prepare_shapes(env, shapes, verts)) return rc;
  gbp::EnvParams e{env->nrows, env->ncols, env->resolution, env->tile_size, env->path_width, env->expansion};
  const uint64_t W64 = uint64_t(env->ncols) * env->resolution, H64 = uint64_t(env->nrows) * env->resolution;
  if (W64 > 65535u * 64u || H64 > 65535u) return fail(GBP_ERR_BAD_ARGUMENT, "environment: image too large");
  const uint32_t W = uint32_t(W64), H = uint32_t(H64);
  const size_t npx = size_t(W) * H, ntile = size_t(env->nrows) * env->ncols;
  uint32_t *d_tiles = nullptr;
  uint8_t *d_img = nullptr;
  CK(dalloc(d_tiles, ntile));
  CK(dalloc(d_img, npx));
  CK(cudaMemcpyAsync(d_tiles, env->tiles, ntile * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  gbp::EnvShape *d_shapes = nullptr;
  double *d_verts = nullptr;
  CK(dalloc(d_shapes, shapes.size()));
  CK(dalloc(d_verts, verts.size()));
  if (!shapes.empty())
    CK(cudaMemcpyAsync(d_shapes, shapes.data(), shapes.size() * sizeof(gbp::EnvShape), cudaMemcpyHostToDevice, st));
  if (!verts.empty())
    CK(cudaMemcpyAsync(d_verts, verts.data(), verts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  const dim3 grid((W + 255) / 256, H);
  gbp::k_env_raster<<<grid, 256, 0, st>>>(e, d_tiles, d_shapes, int(shapes.size()), d_verts, d_img);
  CK(cudaGetLastError());
  *launches += 1;
  const float blur_pixels = env->blur * float(env->resolution);
  if (!(blur_pixels < 1.0f)) {
    const float sigma = blur_pixels <= 0.0f ? 1.0f : blur_pixels;
    const float support = 2.0f * sigma;
    const int D = int(std::ceil(support)) + 2;
    std::vector<float> wt(size_t(2 * D + 1));
    for (int d = -D; d <= D; ++d) wt[size_t(d + D)] = env_gaussian(float(d), sigma);
    float *d_w = nullptr, *d_tmp = nullptr;
    CK(dalloc(d_w, wt.size()));
    CK(dalloc(d_tmp, npx));
    CK(cudaMemcpyAsync(d_w, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    gbp::k_blur_rows<<<grid, 256, 0, st>>>(d_img, d_tmp, W, H, support, d_w, D);
    gbp::k_blur_cols<<<grid, 256, 0, st>>>(d_tmp, d_img, W, H, support, d_w, D);
    CK(cudaGetLastError());
    *launches += 2;
    CK(cudaStreamSynchronize(st));  // wt goes out of scope
    cudaFree(d_w);
    cudaFree(d_tmp);
  }
  CK(cudaStreamSynchronize(st));
  cudaFree(d_tiles);
  cudaFree(d_shapes);
  cudaFree(d_verts);
  *d_gray = d_img;
  *Wo = W;
  *Ho = H;
  return 0;
}
}  // namespace

int gbp_env_to_sdf_image(const gbp_environment_t *env, int32_t device, uint8_t *rgb8) {
  if (!rgb8) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_env_to_sdf_image: null output");
  CK(cudaSetDevice(device));
  uint8_t *d_gray = nullptr, *d_rgb = nullptr;
  uint32_t W = 0, H = 0;
  int64_t launches = 0;
  if (int rc = env_to_sdf_device(env, nullptr, &d_gray, &W, &H, &launches)) return rc;
  const size_t npx = size_t(W) * H;
  CK(dalloc(d_rgb, 3 * npx));
  gbp::k_gray_to_rgb<<<blocks_for(int64_t(npx), 256), 256>>>(d_gray, d_rgb, npx);
  CK(cudaGetLastError());
  CK(cudaMemcpy(rgb8, d_rgb, 3 * npx, cudaMemcpyDeviceToHost));
  cudaFree(d_gray);
  cudaFree(d_rgb);
  return 0;
}

int gbp_world_set_sdf_from_environment(gbp_world_t *w, const gbp_environment_t *env) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  uint8_t *d_gray = nullptr;
  uint32_t W = 0, H = 0;
  if (int rc = env_to_sdf_device(env, w->stream, &d_gray, &W, &H, &w->launches)) return rc;
  cudaFree(w->sdf_dev);
  w->sdf_dev = d_gray;
  w->s.sdf = w->sdf_dev;
  w->s.sdf_w = int32_t(W);
  w->s.sdf_h = int32_t(H);
  refresh_scalars(w);
  return 0;
}

int gbp_world_add_robots(gbp_world_t *w, int32_t n, const float *radii, const uint32_t *timesteps,
                         const double *init_means, const float *positions, const int32_t *wp_offsets,
                         const float *wp_xy) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (n < 0 || (n > 0 && (!radii || !timesteps || !init_means || !positions || !wp_offsets || !wp_xy)))
    return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_add_robots: null input");
  if (n == 0) return 0;
  if (set_device(w)) return GBP_ERR_CUDA;
  Store &s = w->s;
  const int V = s.V;
  if (!w->timesteps.empty() && !std::equal(w->timesteps.begin(), w->timesteps.end(), timesteps))
    return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_add_robots: variable_timesteps differ from the world's");
  w->timesteps.assign(timesteps, timesteps + V);
  cudaStream_t st = w->stream;
  if (w->sh.ws > 1 && w->grp->committed) {
    // Spawned later (FormationSpawner repeats, spawner.rs:186-323): new entities get ids above every existing one, so
    // they can only join the LAST shard without renumbering anybody (ids are contiguous per shard); where they stand
    // does not matter for correctness — neighbours are found by position.  The group must publish the new total:
    // gbp_world_commit_shards again, on every rank, before the next collective call.
    if (w->sh.rank != w->sh.ws - 1)
      return fail(GBP_ERR_STATE, "gbp_world_add_robots: after gbp_world_commit_shards robots can join the last shard only "
                                 "(global ids are contiguous per shard and follow spawn order)");
    w->grp->committed = false;
  }
  const int64_t N0 = s.Nloc, N1 = int64_t(s.Nloc) + n;
  if (int rc = topology_inputs_change(w)) return rc;
  if (int rc = reserve_robots(w, N1)) return rc;
  const int64_t newNV = s.NV, used = N0 * V;
  {  // eoff: new robots start without edges
    int64_t *eoff = nullptr;
    CK(dalloc(eoff, N1 + 1));
    std::vector<int64_t> h(size_t(N1) + 1, s.E);
    if (s.eoff && N0 > 0) CK(cudaMemcpy(h.data(), s.eoff, size_t(N0 + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    else h[0] = 0;
    for (int64_t k = N0 + 1; k <= N1; ++k) h[k] = s.E;
    CK(cudaMemcpy(eoff, h.data(), h.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    cudaFree(s.eoff);
    s.eoff = eoff;
  }
  CK(cudaMemsetAsync(s.idle + N0, 0, size_t(n), st));
  CK(cudaMemsetAsync(s.finished + N0, 0, size_t(n), st));
  CK(cudaMemsetAsync(s.gone + N0, 0, size_t(n) * sizeof(float), st));
  CK(cudaMemsetAsync(s.latest + N0, 0, size_t(n), st));
  CK(cudaMemsetAsync(s.iter_factor + N0, 0, size_t(n) * sizeof(uint32_t), st));
  CK(cudaMemsetAsync(s.mode + N0, 1, size_t(n), st));  // k_iterate hands a robot to k_iterate_axis once it qualifies
  CK(cudaMemsetAsync(s.nlow + N0, 0, size_t(n) * sizeof(int32_t), st));
  CK(cudaMemsetAsync(s.coll_hits + N0, 0, size_t(n) * sizeof(uint32_t), st));
  s.N = int32_t(N1);
  s.Nloc = int32_t(N1);
  const int64_t N1cap = s.cap;
  mark_halo_stale(w);

  // ---- host staging of the new robots (RobotBundle::new, robot.rs:1134-1355)
  const int64_t nv = int64_t(n) * V;
  std::vector<double> mu(size_t(4) * nv), pl(nv), dt(nv, 0.0);
  std::vector<float> rad(radii, radii + n), t0(n), pos(size_t(2) * n);
  std::vector<uint8_t> ones(n, 1);
  std::vector<int32_t> gid(n), nwp(n, 1);
  for (int r = 0; r < n; ++r) {
    w->max_radius = std::max(w->max_radius, radii[r]);
    t0[r] = radii[r] / 2.0f / w->cfg.target_speed;  // :1225 (f32)
    pos[r] = positions[2 * r];
    pos[size_t(n) + r] = positions[2 * r + 1];
    gid[r] = int32_t(N0 + r);
    for (int i = 0; i < V; ++i) {
      const int64_t t = int64_t(r) * V + i;
      for (int k = 0; k < 4; ++k) mu[size_t(k) * nv + t] = init_means[4 * t + k];
      // sigma 1e30 on the first/last variable, INFINITY -> all-zero elsewhere (:1198-1210, variable.rs:146-148)
      pl[t] = (i == 0 || i == V - 1) ? 1e30 : 0.0;
      if (i < V - 1) dt[t] = double(t0[r] * float(timesteps[i + 1] - timesteps[i]));  // :1232
    }
  }
  double *d_mu = nullptr;  // [5][nv] staging: the four mean planes and delta_t; k_init_vars reads them
  CK(dalloc(d_mu, size_t(5) * size_t(nv)));
  CK(cudaMemcpyAsync(d_mu, mu.data(), size_t(4) * size_t(nv) * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_mu + size_t(4) * size_t(nv), dt.data(), size_t(nv) * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(upload_planes(s.prior_lam, newNV, used, pl.data(), 1, nv, st));
  CK(upload_planes(s.radius, N1cap, N0, rad.data(), 1, n, st));
  CK(upload_planes(s.t0, N1cap, N0, t0.data(), 1, n, st));
  CK(upload_planes(s.pos, N1cap, N0, pos.data(), 2, n, st));
  CK(upload_planes(s.antenna, N1cap, N0, ones.data(), 1, n, st));
  CK(upload_planes(s.gid, N1cap, N0, gid.data(), 1, n, st));
  CK(upload_planes(s.next_wp, N1cap, N0, nwp.data(), 1, n, st));
  // waypoint polylines (CSR, rebuilt whole)
  const int32_t base = w->wp_off.back();
  for (int r = 0; r < n; ++r) w->wp_off.push_back(base + (wp_offsets[r + 1] - wp_offsets[0]));
  w->wp_xy.insert(w->wp_xy.end(), wp_xy + 2 * size_t(wp_offsets[0]), wp_xy + 2 * size_t(wp_offsets[n]));
  CK(cudaStreamSynchronize(st));
  cudaFree(s.wp_off);
  cudaFree(s.wp_xy);
  CK(dalloc(s.wp_off, w->wp_off.size()));
  CK(dalloc(s.wp_xy, w->wp_xy.size()));
  CK(cudaMemcpyAsync(s.wp_off, w->wp_off.data(), w->wp_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.wp_xy, w->wp_xy.data(), w->wp_xy.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  k_init_vars<<<blocks_for(nv, 256), 256, 0, st>>>(s, used, nv, d_mu);
  CK(cudaGetLastError());
  w->launches += 1;
  if (w->count_messages) {
    k_msg_init_robots<<<blocks_for(n, 256), 256, 0, st>>>(s, N0, n, w->msg_cap, w->msg_cnt);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  // one delta_t per Dynamic factor for the whole world? (t0 = radius / 2 / target_speed, robot.rs:1225)
  for (int r = 0; r < n && w->t0_uniform; ++r) {
    if (!w->t0_seen) {
      w->t0_first = t0[r];
      w->t0_seen = true;
    }
    w->t0_uniform = std::memcmp(&t0[r], &w->t0_first, sizeof(float)) == 0;
  }
  if (!w->t0_uniform) {
    s.dyn_tab = nullptr;
  } else if (!s.dyn_tab) {
    if (!w->dyn_tab_dev) CK(dalloc(w->dyn_tab_dev, size_t(4) * size_t(V)));
    k_dyn_table<<<blocks_for(V, 64), 64, 0, st>>>(s, N0, w->dyn_tab_dev);
    CK(cudaGetLastError());
    w->launches += 1;
    s.dyn_tab = w->dyn_tab_dev;
  }
  if (int rc = refresh_ell(w)) return rc;  // the new robots have no edges yet
  CK(cudaStreamSynchronize(st));  // staging vectors go out of scope
  cudaFree(d_mu);
  return 0;
}

int32_t gbp_world_num_robots(const gbp_world_t *w) { return w ? w->s.Nloc : 0; }

int gbp_world_update_topology(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_update_topology")) return rc;
  if (set_device(w)) return GBP_ERR_CUDA;
  return group_update_topology(w->grp);
}

namespace {
// Copies `bytes` from a borrowed host buffer into the device array `dst`, ordered on the world's stream
// like any other operation, but WITHOUT waiting for the work already queued there: host -> staging slot on
// the upload stream (the call returns once that copy is done, so the caller may reuse its buffer), then
// staging -> dst on the world's stream.  Two slots per kind: a slot is rewritten only after the
// staging -> dst copy of the call before last, which ran a tick ago.
int staged_upload(gbp_world *w, int kind, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return 0;
  if (!w->up_stream) {
    CK(cudaStreamCreateWithFlags(&w->up_stream, cudaStreamNonBlocking));
    for (int q = 0; q < gbp_world::kUpKinds; ++q)
      for (int g = 0; g < 2; ++g) {
        CK(cudaEventCreateWithFlags(&w->ev_up_done[q][g], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&w->ev_up_used[q][g], cudaEventDisableTiming));
      }
  }
  if (bytes > w->up_bytes[kind]) {  // grow-only; rare (robots were added)
    CK(cudaStreamSynchronize(w->stream));
    CK(cudaStreamSynchronize(w->up_stream));
    for (int g = 0; g < 2; ++g) {
      cudaFree(w->up_stage[kind][g]);
      w->up_stage[kind][g] = nullptr;
      CK(cudaMalloc(&w->up_stage[kind][g], bytes));
    }
    w->up_bytes[kind] = bytes;
  }
  const int g = (w->up_gen[kind] ^= 1);
  CK(cudaStreamWaitEvent(w->up_stream, w->ev_up_used[kind][g], 0));
  CK(cudaMemcpyAsync(w->up_stage[kind][g], src, bytes, cudaMemcpyHostToDevice, w->up_stream));
  CK(cudaEventRecord(w->ev_up_done[kind][g], w->up_stream));
  CK(cudaStreamWaitEvent(w->stream, w->ev_up_done[kind][g], 0));
  CK(cudaMemcpyAsync(dst, w->up_stage[kind][g], bytes, cudaMemcpyDeviceToDevice, w->stream));
  CK(cudaEventRecord(w->ev_up_used[kind][g], w->stream));
  CK(cudaStreamSynchronize(w->up_stream));  // the host buffer is free again
  return 0;
}
}  // namespace

int gbp_world_set_comms(gbp_world_t *w, const uint8_t *antenna_active, const uint8_t *idle) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_set_comms")) return rc;
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  if (n == 0) return 0;
  if (antenna_active) {
    if (int rc = staged_upload(w, 0, w->s.antenna, antenna_active, size_t(n))) return rc;
  } else {
    CK(cudaMemsetAsync(w->s.antenna, 1, n, w->stream));
  }
  if (idle) {
    if (int rc = staged_upload(w, 1, w->s.idle, idle, size_t(n))) return rc;
  } else {
    CK(cudaMemsetAsync(w->s.idle, 0, n, w->stream));
  }
  if (w->any_gone) {
    k_keep_gone_idle<<<blocks_for(n, 256), 256, 0, w->stream>>>(w->s);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  mark_halo_stale(w);  // antenna / idle bits travel with the halo
  return 0;
}

int gbp_world_remove_robots(gbp_world_t *w, int32_t m, const int32_t *robots) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_remove_robots")) return rc;
  if (m < 0 || (m > 0 && !robots)) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_remove_robots: null list");
  if (set_device(w)) return GBP_ERR_CUDA;
  for (int k = 0; k < m; ++k)
    if (robots[k] < 0 || robots[k] >= w->s.Nloc)
      return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_remove_robots: robot index out of range");
  if (m == 0) return 0;
  if (int rc = topology_inputs_change(w)) return rc;
  w->gone_host.resize(size_t(w->s.Nloc), 0);
  for (int k = 0; k < m; ++k)
    if (!w->gone_host[robots[k]]) {
      w->gone_host[robots[k]] = 1;
      w->n_gone += 1;
    }
  if (int rc = ensure_scratch(w, size_t(m) * sizeof(int32_t))) return rc;
  CK(cudaMemcpyAsync(w->rb_dev, robots, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, w->stream));
  k_remove_robots<<<blocks_for(m, 256), 256, 0, w->stream>>>(w->s, m, static_cast<const int32_t *>(w->rb_dev));
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(w->stream));  // the caller's list may go away
  w->any_gone = true;
  mark_halo_stale(w);
  return 0;
}

int gbp_world_read_removed(gbp_world_t *w, uint8_t *removed) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!removed) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  std::vector<float> h(size_t(n), 0.0f);
  if (n > 0) {
    CK(cudaMemcpyAsync(h.data(), w->s.gone, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  for (int r = 0; r < n; ++r) removed[r] = h[r] != 0.0f ? 1 : 0;
  return 0;
}

int gbp_world_set_waypoint_index(gbp_world_t *w, const int32_t *next_index) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!next_index) return fail(GBP_ERR_BAD_ARGUMENT, "null index array");
  if (set_device(w)) return GBP_ERR_CUDA;
  return staged_upload(w, 2, w->s.next_wp, next_index, size_t(w->s.Nloc) * sizeof(int32_t));
}

int gbp_world_reached_waypoint(gbp_world_t *w, const gbp_reached_when_t *taskpoint, const gbp_reached_when_t *finished,
                               uint8_t *out_reached) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!taskpoint || !finished) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_reached_waypoint: null criterion");
  for (const gbp_reached_when_t *c : {taskpoint, finished})
    if (c->intersects_with < 0 || c->intersects_with > 2 || c->distance < 0 || c->distance > 1)
      return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_reached_waypoint: bad criterion");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  if (n == 0) return 0;
  uint8_t *d_out = nullptr;
  if (out_reached) {
    if (int rc = ensure_scratch(w, size_t(n))) return rc;
    d_out = static_cast<uint8_t *>(w->rb_dev);
  }
  k_reached_waypoint<<<blocks_for(n, 128), 128, 0, w->stream>>>(w->s, w->p, *taskpoint, *finished, d_out);
  CK(cudaGetLastError());
  w->launches += 1;
  if (out_reached) {
    CK(cudaMemcpyAsync(out_reached, d_out, size_t(n), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  return 0;
}

int gbp_world_update_robot_collisions(gbp_world_t *w0, int64_t *num_collisions, int64_t *colliding_now) {
  if (!w0) return fail(GBP_ERR_BAD_HANDLE, "null world");
  gbp_group *g = w0->grp;
  // ghosts must sit at their current Transform: the halo carries it
  // The monitor walks the InterRobot edges instead of all pairs (planner/collisions.rs:72-143 tests every pair):
  // exact as long as two touching robots are always connected, i.e. r_a + r_b <= comms radius.  Every shard checks
  // its own robots against half the radius, which bounds every pair.
  for (gbp_world *w : g->members)
    if (2.0f * w->max_radius > w->cfg.comms_radius)
      return fail(GBP_ERR_STATE, "gbp_world_update_robot_collisions: a robot diameter exceeds the communication radius; "
                                 "colliding pairs could be out of comms range and would be missed");
  if (int rc = group_halo_join(g)) return rc;
  if (g->ws > 1 && g->halo_stale) {
    if (int rc = group_halo(g)) return rc;
  }
  for (gbp_world *w : g->members) {
    if (set_device(w)) return GBP_ERR_CUDA;
    if (!w->coll_totals) {
      CK(dalloc(w->coll_totals, 2));
      CK(cudaMemsetAsync(w->coll_totals, 0, 2 * sizeof(unsigned long long), w->stream));
    }
    CK(cudaMemsetAsync(w->coll_totals + 1, 0, sizeof(unsigned long long), w->stream));
    if (w->s.Nloc == 0 || w->s.E == 0) continue;
    gbp::k_robot_collisions<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(
        w->s, w->edges[w->cur].egid, w->sh.gfirst[w->sh.rank], w->coll_totals);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  if (num_collisions || colliding_now) return gbp_world_read_collision_totals(w0, num_collisions, colliding_now);
  return 0;
}

int gbp_world_read_collision_totals(gbp_world_t *w, int64_t *num_collisions, int64_t *colliding_now) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  unsigned long long h[2] = {0, 0};
  if (w->coll_totals) {
    CK(cudaMemcpyAsync(h, w->coll_totals, sizeof(h), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  if (num_collisions) *num_collisions = int64_t(h[0]);
  if (colliding_now) *colliding_now = int64_t(h[1]);
  return 0;
}

int gbp_world_read_robot_collisions(gbp_world_t *w, uint32_t *per_robot) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!per_robot) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (w->s.Nloc == 0) return 0;
  CK(cudaMemcpyAsync(per_robot, w->s.coll_hits, size_t(w->s.Nloc) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

int gbp_world_read_waypoint_index(gbp_world_t *w, int32_t *next_index) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!next_index) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (w->s.Nloc == 0) return 0;
  CK(cudaMemcpyAsync(next_index, w->s.next_wp, size_t(w->s.Nloc) * sizeof(int32_t), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

// Both prior updates run for every shard of the group living in this process.
// ---- robot-environment collisions (planner/collisions.rs:368-455) --------------------------------------------
int gbp_world_set_environment_colliders(gbp_world_t *w, int32_t n, const gbp_collider_t *colliders, int32_t num_vertices,
                                        const float *vertices_xy) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (n < 0 || num_vertices < 0 || (n > 0 && !colliders) || (num_vertices > 0 && !vertices_xy))
    return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_set_environment_colliders: bad argument");
  std::vector<gbp::ColliderDev> h(size_t(std::max(n, 1)));
  for (int k = 0; k < n; ++k) {
    const gbp_collider_t &c = colliders[k];
    if (c.kind < 0 || c.kind > 3) return fail(GBP_ERR_BAD_ARGUMENT, "collider kind out of range");
    if (c.kind >= 2) {
      if (c.num_vertices < 3 || c.first_vertex < 0 || int64_t(c.first_vertex) + c.num_vertices > num_vertices ||
          (c.kind == 2 && c.num_vertices != 3))
        return fail(GBP_ERR_BAD_ARGUMENT, "collider vertex range out of bounds");
    }
    // Isometry2::new(translation, angle): UnitComplex::new(angle) = (cos, sin) in f32
    h[k] = {c.kind, c.translation[0], c.translation[1], std::cos(c.angle), std::sin(c.angle), c.radius,
            c.half_extents[0], c.half_extents[1], c.first_vertex, c.num_vertices};
  }
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(w->env_cols);
  cudaFree(w->env_verts);
  cudaFree(w->env_state);
  cudaFree(w->env_hits);
  w->env_cols = nullptr;
  w->env_verts = nullptr;
  w->env_state = w->env_hits = nullptr;  // RobotEnvironmentCollisions::clear (a new environment was loaded)
  w->env_robots = 0;
  w->env_ncol = n;
  w->env_words = (n + 31) / 32;
  if (!w->env_totals) CK(dalloc(w->env_totals, 2));
  CK(cudaMemsetAsync(w->env_totals, 0, 2 * sizeof(unsigned long long), w->stream));
  if (n > 0) {
    CK(dalloc(w->env_cols, size_t(n)));
    CK(cudaMemcpy(w->env_cols, h.data(), size_t(n) * sizeof(gbp::ColliderDev), cudaMemcpyHostToDevice));
  }
  CK(dalloc(w->env_verts, size_t(2) * size_t(std::max(num_vertices, 1))));
  if (num_vertices > 0)
    CK(cudaMemcpy(w->env_verts, vertices_xy, size_t(2) * size_t(num_vertices) * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int gbp_world_update_environment_collisions(gbp_world_t *w, int64_t *num_collisions, int64_t *colliding_now) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!w->env_totals) return fail(GBP_ERR_STATE, "gbp_world_update_environment_collisions: no colliders set");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int32_t n = w->s.Nloc;
  if (n > w->env_robots) {  // robots were added: their histories start Free
    const int64_t cap = std::max<int64_t>(w->s.cap, n);
    CK(regrow(w->env_state, 1, w->env_robots * w->env_words, cap * std::max(w->env_words, 1),
              w->env_robots * w->env_words, w->stream));
    CK(regrow(w->env_hits, 1, w->env_robots, cap, w->env_robots, w->stream));
    w->env_robots = cap;
  }
  CK(cudaMemsetAsync(w->env_totals + 1, 0, sizeof(unsigned long long), w->stream));
  if (n > 0 && w->env_ncol > 0) {
    gbp::k_env_collisions<<<blocks_for(n, 128), 128, 0, w->stream>>>(n, w->s.pos, w->s.cap, w->s.radius, w->s.gone,
                                                                     w->env_ncol, w->env_cols, w->env_verts, w->env_words,
                                                                     w->env_state, w->env_hits, w->env_totals);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  unsigned long long h[2] = {0, 0};
  if (num_collisions || colliding_now) {
    CK(cudaMemcpyAsync(h, w->env_totals, sizeof(h), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  if (num_collisions) *num_collisions = int64_t(h[0]);
  if (colliding_now) *colliding_now = int64_t(h[1]);
  return 0;
}

int gbp_world_read_environment_collisions(gbp_world_t *w, uint32_t *per_robot) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!per_robot) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int64_t n = w->s.Nloc, have = std::min<int64_t>(n, w->env_robots);
  std::fill(per_robot, per_robot + n, 0u);
  if (have > 0 && w->env_hits) {
    CK(cudaMemcpyAsync(per_robot, w->env_hits, size_t(have) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  return 0;
}

// ---- PositionTracker / VelocityTracker (planner/tracking.rs:117-260) ------------------------------------------
namespace {
void free_track_rings(gbp_world *w) {
  gbp::TrackRings &t = w->trk;
  void *ptrs[] = {t.elapsed_ns, t.npos, t.nvel, t.has_prev, t.prev_xy, t.prev_t, t.pos_xy, t.vel_xy, t.vel_t, t.vel_over};
  for (void *q : ptrs) cudaFree(q);
  t = gbp::TrackRings{};
}
}  // namespace

int gbp_world_set_tracking_buffers(gbp_world_t *w, int32_t capacity, uint64_t sample_ns) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (capacity < 1) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_set_tracking_buffers: capacity must be >= 1");
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaStreamSynchronize(w->stream));
  free_track_rings(w);
  w->trk.capacity = capacity;
  w->trk_duration_ns = sample_ns;
  w->trk_seen = 0;
  return 0;
}

int gbp_world_track(gbp_world_t *w, uint64_t delta_ns, double elapsed_seconds) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (w->trk.capacity < 1) return fail(GBP_ERR_STATE, "gbp_world_track: gbp_world_set_tracking_buffers first");
  if (set_device(w)) return GBP_ERR_CUDA;
  gbp::TrackRings &t = w->trk;
  const int32_t n = w->s.Nloc;
  if (n > t.stride) {
    const int64_t cap = std::max<int64_t>(w->s.cap, n), old = t.stride, keep = std::min<int64_t>(w->trk_seen, old);
    cudaStream_t st = w->stream;
    CK(regrow(t.elapsed_ns, 1, old, cap, keep, st));
    CK(regrow(t.npos, 1, old, cap, keep, st));
    CK(regrow(t.nvel, 1, old, cap, keep, st));
    CK(regrow(t.has_prev, 1, old, cap, keep, st));
    CK(regrow(t.prev_xy, 2, old, cap, keep, st));
    CK(regrow(t.prev_t, 1, old, cap, keep, st));
    CK(regrow(t.pos_xy, 2 * t.capacity, old, cap, keep, st));
    CK(regrow(t.vel_xy, 2 * t.capacity, old, cap, keep, st));
    CK(regrow(t.vel_t, t.capacity, old, cap, keep, st));
    CK(regrow(t.vel_over, t.capacity, old, cap, keep, st));
    t.stride = cap;
  }
  if (n > 0) {
    gbp::k_track<<<blocks_for(n, 128), 128, 0, w->stream>>>(n, w->trk_seen, w->s.pos, w->s.cap, w->s.idle, w->s.gone, t,
                                                            w->trk_duration_ns, delta_ns, elapsed_seconds);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  w->trk_seen = n;
  return 0;
}

int gbp_world_read_tracks(gbp_world_t *w, uint32_t *num_positions, float *positions_xy, uint32_t *num_velocities,
                          float *velocities_xy, double *velocity_timestamp, double *velocity_measured_over) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (w->trk.capacity < 1) return fail(GBP_ERR_STATE, "gbp_world_read_tracks: gbp_world_set_tracking_buffers first");
  if (set_device(w)) return GBP_ERR_CUDA;
  const gbp::TrackRings &t = w->trk;
  const int64_t n = std::min<int64_t>(w->s.Nloc, t.stride), C = t.capacity;
  const int64_t nall = w->s.Nloc;
  if (num_positions) std::fill(num_positions, num_positions + nall, 0u);
  if (num_velocities) std::fill(num_velocities, num_velocities + nall, 0u);
  if (n == 0) return 0;
  cudaStream_t st = w->stream;
  // device planes [k][r] -> caller rows [k][n] (slot-major, n robots per plane)
  auto rows = [&](void *dst, const void *src, size_t elem, int64_t planes) {
    return cudaMemcpy2DAsync(dst, size_t(nall) * elem, src, size_t(t.stride) * elem, size_t(n) * elem, size_t(planes),
                             cudaMemcpyDeviceToHost, st);
  };
  if (num_positions) CK(rows(num_positions, t.npos, 4, 1));
  if (num_velocities) CK(rows(num_velocities, t.nvel, 4, 1));
  if (positions_xy) CK(rows(positions_xy, t.pos_xy, 4, 2 * C));
  if (velocities_xy) CK(rows(velocities_xy, t.vel_xy, 4, 2 * C));
  if (velocity_timestamp) CK(rows(velocity_timestamp, t.vel_t, 8, C));
  if (velocity_measured_over) CK(rows(velocity_measured_over, t.vel_over, 8, C));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int gbp_world_update_prior_of_horizon_state(gbp_world_t *w0) {
  if (!w0) return fail(GBP_ERR_BAD_HANDLE, "null world");
  for (gbp_world *w : w0->grp->members) {
    if (set_device(w)) return GBP_ERR_CUDA;
    w->epoch += 1;
    if (w->s.Nloc == 0) continue;
    if (w->count_messages) {
      k_msg_prior_horizon<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(
          w->s, w->cfg.iterations_internal, w->msg_cap, w->msg_cnt, w->edges[w->cur].cap, w->edges[w->cur].e_msg);
      w->launches += 1;
    }
    ProfileScope ps(w, GBP_PROFILE_PRIORS);
    k_prior_horizon<<<blocks_for(int64_t(w->s.Nloc) * 32, 128), 128, 0, w->stream>>>(
        w->s, w->p, w->epoch, double(w->cfg.delta_t), double(w->cfg.target_speed), w->cfg.iterations_internal);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  w0->grp->halo_stale = true;
  return 0;
}

int gbp_world_update_prior_of_current_state(gbp_world_t *w0) {
  if (!w0) return fail(GBP_ERR_BAD_HANDLE, "null world");
  for (gbp_world *w : w0->grp->members) {
    if (set_device(w)) return GBP_ERR_CUDA;
    w->epoch += 1;
    if (w->s.Nloc == 0) continue;
    if (w->count_messages) {
      k_msg_prior_current<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(w->s, w->msg_cap, w->msg_cnt);
      w->launches += 1;
    }
    ProfileScope ps(w, GBP_PROFILE_PRIORS);
    k_prior_current<<<blocks_for(int64_t(w->s.Nloc) * 32, 128), 128, 0, w->stream>>>(w->s, w->p, w->epoch, w->cfg.delta_t);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  w0->grp->halo_stale = true;
  w0->grp->topo_version += 1;  // Transform.translation moved
  return group_topology_early(w0->grp);
}

namespace {
// Both prior updates of a tick in one launch (k_prior_both); false if the world has to take them one by one.
bool priors_fusable(const gbp_world *w0) {
  static const bool off = [] { const char *e = std::getenv("GBP_PRIORS_FUSED"); return e && e[0] == '0'; }();
  if (off) return false;
  for (const gbp_world *w : w0->grp->members)
    if (w->s.V < 3 || w->count_messages) return false;
  return true;
}
int update_priors_fused(gbp_world *w0) {
  for (gbp_world *w : w0->grp->members) {
    if (set_device(w)) return GBP_ERR_CUDA;
    w->epoch += 2;
    if (w->s.Nloc == 0) continue;
    ProfileScope ps(w, GBP_PROFILE_PRIORS);
    k_prior_both<<<blocks_for(int64_t(w->s.Nloc) * 32, 128), 128, 0, w->stream>>>(
        w->s, w->p, w->epoch - 1, w->epoch, double(w->cfg.delta_t), double(w->cfg.target_speed),
        w->cfg.iterations_internal, w->cfg.delta_t);
    CK(cudaGetLastError());
    w->launches += 1;
  }
  w0->grp->halo_stale = true;
  w0->grp->topo_version += 1;  // Transform.translation moved
  return group_topology_early(w0->grp);
}
}  // namespace

int gbp_world_change_prior_of_variable(gbp_world_t *w, int32_t var, int32_t m, const int32_t *robots,
                                       const double *new_means) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_change_prior_of_variable")) return rc;
  if (var < 0 || var >= w->s.V || m < 0 || (m > 0 && (!robots || !new_means)))
    return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_change_prior_of_variable: bad argument");
  if (m == 0) return 0;
  for (int k = 0; k < m; ++k)
    if (robots[k] < 0 || robots[k] >= w->s.Nloc) return fail(GBP_ERR_BAD_ARGUMENT, "robot index out of range");
  if (set_device(w)) return GBP_ERR_CUDA;
  int32_t *dr = nullptr;
  double *dm = nullptr;
  CK(dalloc(dr, m));
  CK(dalloc(dm, size_t(4) * m));
  CK(cudaMemcpyAsync(dr, robots, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, w->stream));
  CK(cudaMemcpyAsync(dm, new_means, size_t(4) * m * sizeof(double), cudaMemcpyHostToDevice, w->stream));
  w->epoch += 1;
  if (w->count_messages) {
    k_msg_change_prior<<<blocks_for(m, 128), 128, 0, w->stream>>>(w->s, var, m, dr, w->msg_cap, w->msg_cnt,
                                                                 w->edges[w->cur].cap, w->edges[w->cur].e_msg);
    w->launches += 1;
  }
  k_change_prior_list<<<blocks_for(m, 128), 128, 0, w->stream>>>(w->s, w->p, w->epoch, var, m, dr, dm);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(dr);
  cudaFree(dm);
  mark_halo_stale(w);
  return 0;
}

namespace {
int check_robot_list(const gbp_world *w, int32_t m, const int32_t *robots, const char *what) {
  if (m < 0 || (m > 0 && !robots)) return fail(GBP_ERR_BAD_ARGUMENT, std::string(what) + ": bad robot list");
  for (int k = 0; k < m; ++k)
    if (robots[k] < 0 || robots[k] >= w->s.Nloc) return fail(GBP_ERR_BAD_ARGUMENT, std::string(what) + ": robot index out of range");
  return 0;
}
}  // namespace

int gbp_world_set_tracking_path(gbp_world_t *w, int32_t m, const int32_t *robots, const int32_t *wp_offsets,
                                const float *wp_xy) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = check_robot_list(w, m, robots, "gbp_world_set_tracking_path")) return rc;
  if (m == 0) return 0;
  if (!wp_offsets || !wp_xy) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_set_tracking_path: null path");
  for (int k = 0; k < m; ++k)
    if (wp_offsets[k + 1] - wp_offsets[k] < 2)
      return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_set_tracking_path: a path needs two or more points");
  if (set_device(w)) return GBP_ERR_CUDA;
  // the waypoint CSR is rebuilt whole on the host (a path change is a per-mission event, not per tick)
  const int n = w->s.Nloc;
  std::vector<std::vector<float>> poly(static_cast<size_t>(n));
  for (int r = 0; r < n; ++r)
    poly[r].assign(w->wp_xy.begin() + 2 * size_t(w->wp_off[r]), w->wp_xy.begin() + 2 * size_t(w->wp_off[r + 1]));
  for (int k = 0; k < m; ++k)
    poly[robots[k]].assign(wp_xy + 2 * size_t(wp_offsets[k]), wp_xy + 2 * size_t(wp_offsets[k + 1]));
  w->wp_off.assign(1, 0);
  w->wp_xy.clear();
  for (int r = 0; r < n; ++r) {
    w->wp_xy.insert(w->wp_xy.end(), poly[r].begin(), poly[r].end());
    w->wp_off.push_back(int32_t(w->wp_xy.size() / 2));
  }
  cudaStream_t st = w->stream;
  CK(cudaStreamSynchronize(st));
  cudaFree(w->s.wp_off);
  cudaFree(w->s.wp_xy);
  CK(dalloc(w->s.wp_off, w->wp_off.size()));
  CK(dalloc(w->s.wp_xy, w->wp_xy.size()));
  CK(cudaMemcpyAsync(w->s.wp_off, w->wp_off.data(), w->wp_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(w->s.wp_xy, w->wp_xy.data(), w->wp_xy.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  int32_t *dr = nullptr;
  CK(dalloc(dr, m));
  CK(cudaMemcpyAsync(dr, robots, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  k_set_next_wp<<<blocks_for(m, 128), 128, 0, st>>>(w->s, m, dr, 1);  // Route::update_waypoints: target_index = 1
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(st));
  cudaFree(dr);
  return 0;
}

namespace {
// The shards that hold InterRobot factors toward the robots `gids` (ascending global ids) freeze them.
int freeze_on_shard(gbp_world *v, const int32_t *d_ids, int n_ids) {
  if (n_ids == 0 || v->s.E == 0) return 0;
  CK(cudaSetDevice(v->device));
  k_freeze_edges_toward<<<blocks_for(v->s.E, 256), 256, 0, v->stream>>>(v->s, v->edges[v->cur].egid, n_ids, d_ids);
  CK(cudaGetLastError());
  v->launches += 1;
  return 0;
}
}  // namespace

int gbp_world_reset_variables(gbp_world_t *w, int32_t m, const int32_t *robots, const double *means,
                              double first_last_sigma, double inbetween_sigma) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_reset_variables")) return rc;
  if (int rc = check_robot_list(w, m, robots, "gbp_world_reset_variables")) return rc;
  gbp_group *g = w->grp;
  const bool collective = g->ws > 1 && g->nccl;  // every rank calls, possibly with m == 0
  if (m == 0 && !collective) return 0;
  if (m > 0 && !means) return fail(GBP_ERR_BAD_ARGUMENT, "gbp_world_reset_variables: null means");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (int rc = group_halo_join(g)) return rc;
  const int V = w->s.V;
  cudaStream_t st = w->stream;
  int32_t *dr = nullptr, *dg = nullptr;
  double *dm = nullptr;
  // the robots' global ids, ascending: what the other shards look for among their neighbours
  std::vector<int32_t> gids(robots, robots + m);
  for (int32_t &x : gids) x += w->sh.gfirst[w->sh.rank];
  std::sort(gids.begin(), gids.end());
  if (m > 0) {
    CK(dalloc(dr, m));
    CK(dalloc(dg, m));
    CK(dalloc(dm, size_t(4) * V * m));
    CK(cudaMemcpyAsync(dr, robots, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg, gids.data(), size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dm, means, size_t(4) * V * m * sizeof(double), cudaMemcpyHostToDevice, st));
    k_reset_variables<<<blocks_for(int64_t(m) * V, 128), 128, 0, st>>>(w->s, w->p, m, dr, dm, first_last_sigma,
                                                                       inbetween_sigma);
    if (w->s.eoff && w->s.E > 0) k_reset_reverse_edges<<<m, 64, 0, st>>>(w->s, m, dr);
    CK(cudaGetLastError());
    w->launches += 2;
  }
  int rc = 0;
  int32_t *d_all = nullptr;
  if (g->ws > 1 && !g->nccl) {
    // in-process shards: one device, one stream — the other members' edges are frozen directly
    for (gbp_world *v : g->members)
      if (v != w && (rc = freeze_on_shard(v, dg, m))) break;
  } else if (collective) {
    // 1. how many robots every shard resets; 2. their global ids; 3. freeze the own edges toward them
    const int ws = g->ws, rank = w->sh.rank;
    std::vector<gbp::XferPlan> plans(1);
    int64_t *h = w->hdr_host + 4 * gbp::kMaxShards;
    h[0] = m;
    h[1] = h[2] = h[3] = 0;
    CK(cudaMemcpyAsync(w->hdr_send, h, 4 * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    for (int q = 0; q < ws; ++q) {
      if (q == rank) continue;
      plans[0].sends.push_back({q, w->hdr_send, 4 * sizeof(int64_t)});
      plans[0].recvs.push_back({q, w->hdr_recv + 4 * q, 4 * sizeof(int64_t)});
    }
    if ((rc = exchange(g, plans)) == 0) {
      CK(cudaMemcpyAsync(w->hdr_host, w->hdr_recv, 4 * size_t(gbp::kMaxShards) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<int64_t> off(size_t(ws) + 1, 0);
      for (int q = 0; q < ws; ++q) off[q + 1] = off[q] + (q == rank ? 0 : w->hdr_host[4 * q]);
      const int64_t total = off[ws];
      if (total > 0) CK(dalloc(d_all, size_t(total)));
      plans[0].clear();
      for (int q = 0; q < ws; ++q) {
        if (q == rank) continue;
        if (m > 0) plans[0].sends.push_back({q, dg, size_t(m) * sizeof(int32_t)});
        if (off[q + 1] > off[q])
          plans[0].recvs.push_back({q, d_all + off[q], size_t(off[q + 1] - off[q]) * sizeof(int32_t)});
      }
      // peers' id ranges ascend with the rank and every list is sorted: the concatenation is sorted
      if ((rc = exchange(g, plans)) == 0) rc = freeze_on_shard(w, d_all, int(total));
    }
  }
  CK(cudaStreamSynchronize(st));
  cudaFree(dr);
  cudaFree(dg);
  cudaFree(dm);
  cudaFree(d_all);
  if (rc) return rc;
  mark_halo_stale(w);
  return 0;
}

int gbp_world_reset_tracking_factors(gbp_world_t *w, int32_t m, const int32_t *robots) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = check_robot_list(w, m, robots, "gbp_world_reset_tracking_factors")) return rc;
  if (m == 0) return 0;
  if (set_device(w)) return GBP_ERR_CUDA;
  int32_t *dr = nullptr;
  CK(dalloc(dr, m));
  CK(cudaMemcpyAsync(dr, robots, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, w->stream));
  k_reset_tracking<<<blocks_for(int64_t(m) * w->s.V, 128), 128, 0, w->stream>>>(w->s, m, dr);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(dr);
  return 0;
}

int gbp_world_iterate_schedule(gbp_world_t *w, int32_t n, const uint8_t *internal, const uint8_t *external) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (n < 0 || (n > 0 && (!internal || !external))) return fail(GBP_ERR_BAD_ARGUMENT, "bad schedule");
  if (w->grp->pending_internal_factor || w->grp->pending_external_factor)
    return fail(GBP_ERR_STATE, "a half-iteration pair is open");
  if (set_device(w)) return GBP_ERR_CUDA;
  return run_schedule(w->grp, n, internal, external);
}

int gbp_world_iterate(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  uint8_t oi[256], oe[256];
  // `config.gbp.iteration_schedule.internal as u8` (robot.rs:1782-1783): wraps above 255
  const int n = gbp_schedule(w->cfg.schedule_kind, uint8_t(w->cfg.iterations_internal),
                             uint8_t(w->cfg.iterations_external), oi, oe);
  if (n < 0) return n;
  return gbp_world_iterate_schedule(w, n, oi, oe);
}

int gbp_world_internal_factor_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (w->grp->pending_internal_factor || w->grp->pending_external_factor) return fail(GBP_ERR_STATE, "half-iteration order");
  w->grp->pending_internal_factor = true;
  return 0;
}
int gbp_world_internal_variable_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!w->grp->pending_internal_factor) return fail(GBP_ERR_STATE, "internal_variable_iteration without internal_factor_iteration");
  w->grp->pending_internal_factor = false;
  if (set_device(w)) return GBP_ERR_CUDA;
  return group_launch<false, true>(w->grp);
}
int gbp_world_external_factor_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (w->grp->pending_internal_factor || w->grp->pending_external_factor) return fail(GBP_ERR_STATE, "half-iteration order");
  w->grp->pending_external_factor = true;
  return 0;
}
int gbp_world_external_variable_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!w->grp->pending_external_factor) return fail(GBP_ERR_STATE, "external_variable_iteration without external_factor_iteration");
  w->grp->pending_external_factor = false;
  if (set_device(w)) return GBP_ERR_CUDA;
  return group_launch<true, false>(w->grp);
}

int gbp_world_step(gbp_world_t *w) {
  int rc;
  if ((rc = gbp_world_update_topology(w))) return rc;
  if (priors_fusable(w)) {
    if ((rc = update_priors_fused(w))) return rc;
  } else {
    if ((rc = gbp_world_update_prior_of_horizon_state(w))) return rc;
    if ((rc = gbp_world_update_prior_of_current_state(w))) return rc;
  }
  return gbp_world_iterate(w);
}

int gbp_world_change_factor_enabled(gbp_world_t *w, int32_t kind, uint8_t enabled) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  uint8_t *flags[4] = {&w->cfg.enable_dynamic, &w->cfg.enable_interrobot, &w->cfg.enable_obstacle,
                       &w->cfg.enable_tracking};
  if (kind < 0 || kind > 3) return fail(GBP_ERR_BAD_ARGUMENT, "factor kind out of range");
  *flags[kind] = enabled ? 1 : 0;
  refresh_scalars(w);
  return 0;
}

int gbp_world_set_safety_distance_multiplier(gbp_world_t *w, float multiplier) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!(multiplier > 0.0f)) return fail(GBP_ERR_BAD_ARGUMENT, "multiplier must be > 0");
  if (set_device(w)) return GBP_ERR_CUDA;
  w->cfg.safety_distance_multiplier = multiplier;
  if (w->s.Nloc && w->s.E) {
    k_set_dsafe<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(w->s, double(multiplier));
    CK(cudaGetLastError());
    w->launches += 1;
    if (int rc = refresh_ell(w)) return rc;
  }
  return 0;
}

int gbp_world_set_schedule(gbp_world_t *w, int32_t kind, int32_t internal, int32_t external) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (kind < 0 || kind > 4 || internal < 0 || external < 0) return fail(GBP_ERR_BAD_ARGUMENT, "bad schedule");
  w->cfg.schedule_kind = kind;
  w->cfg.iterations_internal = internal;
  w->cfg.iterations_external = external;
  return 0;
}

namespace {
int wait_readback(gbp_world *w) {
  if (w->copy_pending) {
    CK(cudaEventSynchronize(w->ev_copied));
    w->copy_pending = false;
  }
  return 0;
}
int read_beliefs_impl(gbp_world *w, double *eta, double *lam, double *mean, double *cov, uint8_t *valid, bool async) {
  if (set_device(w)) return GBP_ERR_CUDA;
  if (int rc = wait_readback(w)) return rc;  // the scratch of a pending copy is about to be reused
  const int64_t nv = int64_t(w->s.Nloc) * w->s.V;
  if (nv == 0) return 0;
  // one grow-only device scratch, carved into the requested AoS arrays
  const size_t b_eta = eta ? size_t(4) * nv * 8 : 0, b_lam = lam ? size_t(16) * nv * 8 : 0,
               b_mean = mean ? size_t(4) * nv * 8 : 0, b_cov = cov ? size_t(16) * nv * 8 : 0,
               b_valid = valid ? size_t(nv) : 0;
  if (int rc = ensure_scratch(w, b_eta + b_lam + b_mean + b_cov + b_valid)) return rc;
  char *base = static_cast<char *>(w->rb_dev);
  double *d_eta = eta ? reinterpret_cast<double *>(base) : nullptr;
  double *d_lam = lam ? reinterpret_cast<double *>(base + b_eta) : nullptr;
  double *d_mean = mean ? reinterpret_cast<double *>(base + b_eta + b_lam) : nullptr;
  double *d_cov = cov ? reinterpret_cast<double *>(base + b_eta + b_lam + b_mean) : nullptr;
  uint8_t *d_valid = valid ? reinterpret_cast<uint8_t *>(base + b_eta + b_lam + b_mean + b_cov) : nullptr;
  k_gather_beliefs<<<blocks_for(nv, 256), 256, 0, w->stream>>>(w->s, w->p, d_eta, d_lam, d_mean, d_cov, d_valid);
  CK(cudaGetLastError());
  w->launches += 1;
  cudaStream_t cs = w->stream;
  if (async) {
    if (!w->copy_stream) {
      CK(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&w->ev_gathered, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&w->ev_copied, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(w->ev_gathered, w->stream));
    CK(cudaStreamWaitEvent(w->copy_stream, w->ev_gathered, 0));
    cs = w->copy_stream;
  }
  if (eta) CK(cudaMemcpyAsync(eta, d_eta, b_eta, cudaMemcpyDeviceToHost, cs));
  if (lam) CK(cudaMemcpyAsync(lam, d_lam, b_lam, cudaMemcpyDeviceToHost, cs));
  if (mean) CK(cudaMemcpyAsync(mean, d_mean, b_mean, cudaMemcpyDeviceToHost, cs));
  if (cov) CK(cudaMemcpyAsync(cov, d_cov, b_cov, cudaMemcpyDeviceToHost, cs));
  if (valid) CK(cudaMemcpyAsync(valid, d_valid, b_valid, cudaMemcpyDeviceToHost, cs));
  if (async) {
    CK(cudaEventRecord(w->ev_copied, cs));
    w->copy_pending = true;
  } else {
    CK(cudaStreamSynchronize(w->stream));
  }
  return 0;
}
}  // namespace

int gbp_world_read_beliefs(gbp_world_t *w, double *eta, double *lam, double *mean, double *cov, uint8_t *valid) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (int rc = pair_open(w, "gbp_world_read_beliefs")) return rc;
  return read_beliefs_impl(w, eta, lam, mean, cov, valid, false);
}

int gbp_world_read_beliefs_async(gbp_world_t *w, double *eta, double *lam, double *mean, double *cov, uint8_t *valid) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  return read_beliefs_impl(w, eta, lam, mean, cov, valid, true);
}

int gbp_world_readback_wait(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  return wait_readback(w);
}

int gbp_world_read_positions(gbp_world_t *w, float *xy) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!xy) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  std::vector<float> h(size_t(2) * n);
  CK(cudaMemcpy2DAsync(h.data(), size_t(n) * 4, w->s.pos, size_t(w->s.cap) * 4, size_t(n) * 4, 2,
                       cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  for (int r = 0; r < n; ++r) {
    xy[2 * r] = h[r];
    xy[2 * r + 1] = h[size_t(n) + r];
  }
  return 0;
}

int64_t gbp_world_read_connections(gbp_world_t *w, int64_t *offsets, int32_t *neighbours, int64_t *robot_number,
                                   int64_t capacity) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!offsets) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  const Store &s = w->s;
  if (s.E > capacity) return fail(GBP_ERR_BAD_ARGUMENT, "capacity too small");
  CK(cudaStreamSynchronize(w->stream));
  if (w->t_err) {
    int32_t err = 0;
    CK(cudaMemcpy(&err, w->t_err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (err) return fail(GBP_ERR_STATE, "robot_number lookup failed during the last topology update");
  }
  if (s.Nloc == 0) {
    offsets[0] = 0;
    return 0;
  }
  CK(cudaMemcpy(offsets, s.eoff, size_t(s.Nloc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (s.E > 0) {
    const gbp_world::EdgeSet &e = w->edges[w->cur];
    if (neighbours) CK(cudaMemcpy(neighbours, e.egid, size_t(s.E) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (robot_number) {
      static_assert(sizeof(int64_t) == sizeof(uint64_t), "");
      CK(cudaMemcpy(robot_number, e.e_own, size_t(s.E) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    if (w->cfg.strict_reference_quirks) {
      // robots_connected_with lists live connections only: the factor sets the lossy deletion left behind
      // (e_frozen bit 2) are dropped from the answer, the lists closed up
      std::vector<uint8_t> fl(size_t(s.E));
      CK(cudaMemcpy(fl.data(), e.e_frozen, size_t(s.E), cudaMemcpyDeviceToHost));
      const std::vector<int64_t> orig(offsets, offsets + s.Nloc + 1);
      int64_t out = 0;
      for (int32_t r = 0; r < s.Nloc; ++r) {
        offsets[r] = out;
        for (int64_t k = orig[size_t(r)]; k < orig[size_t(r) + 1]; ++k) {
          if (fl[size_t(k)] & gbp::kEdgeZombie) continue;
          if (neighbours) neighbours[out] = neighbours[k];
          if (robot_number) robot_number[out] = robot_number[k];
          ++out;
        }
      }
      offsets[s.Nloc] = out;
      return out;
    }
  }
  return s.E;
}

int gbp_world_set_message_counting(gbp_world_t *w, int32_t on) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (on && w->sh.ws > 1)
    return fail(GBP_ERR_STATE, "gbp_world_set_message_counting: single-GPU worlds only (what a ghost robot's prior update "
                               "delivers to the own factors is not visible to the shard)");
  if (on && !w->count_messages && w->s.Nloc > 0)
    return fail(GBP_ERR_STATE, "gbp_world_set_message_counting: turn the counters on before the first robot is added "
                               "(the reference counts from the creation of the graph on)");
  w->count_messages = on != 0;
  return 0;
}

int gbp_world_read_message_counts(gbp_world_t *w, int64_t *counts) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!counts) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (int rc = pair_open(w, "gbp_world_read_message_counts")) return rc;
  if (!w->count_messages) return fail(GBP_ERR_STATE, "message counting is off (gbp_world_set_message_counting)");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int64_t n = w->s.Nloc;
  if (n == 0) return 0;
  if (int rc = ensure_scratch(w, size_t(4) * size_t(n) * sizeof(int64_t))) return rc;
  int64_t *d = static_cast<int64_t *>(w->rb_dev);
  k_msg_gather<<<blocks_for(n, 128), 128, 0, w->stream>>>(w->s, w->msg_cap, w->msg_cnt, w->edges[w->cur].cap,
                                                         w->edges[w->cur].e_msg, d);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaMemcpyAsync(counts, d, size_t(4) * size_t(n) * sizeof(int64_t), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

int gbp_world_read_tracking(gbp_world_t *w, int64_t *record, float *last_pos, double *last_value) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!record || !last_pos || !last_value) return fail(GBP_ERR_BAD_ARGUMENT, "null output");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int64_t nv = int64_t(w->s.Nloc) * w->s.V;
  if (nv == 0) return 0;
  const size_t bytes = size_t(nv) * (sizeof(int64_t) + sizeof(double) + 2 * sizeof(float));
  if (int rc = ensure_scratch(w, bytes)) return rc;
  int64_t *d_rec = static_cast<int64_t *>(w->rb_dev);
  double *d_val = reinterpret_cast<double *>(d_rec + nv);
  float *d_pos = reinterpret_cast<float *>(d_val + nv);
  k_gather_tracking<<<blocks_for(nv, 256), 256, 0, w->stream>>>(w->s, nv, d_rec, d_pos, d_val);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaMemcpyAsync(record, d_rec, size_t(nv) * sizeof(int64_t), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaMemcpyAsync(last_value, d_val, size_t(nv) * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaMemcpyAsync(last_pos, d_pos, size_t(nv) * 2 * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

int gbp_world_sdf_lookup(gbp_world_t *w, int32_t m, const double *xy, uint32_t *px, uint32_t *py, double *value) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (m < 0 || (m > 0 && (!xy || !px || !py || !value))) return fail(GBP_ERR_BAD_ARGUMENT, "null argument");
  if (m == 0) return 0;
  if (set_device(w)) return GBP_ERR_CUDA;
  double *dxy = nullptr, *dv = nullptr;
  uint32_t *dpx = nullptr, *dpy = nullptr;
  CK(dalloc(dxy, size_t(2) * m)); CK(dalloc(dv, m)); CK(dalloc(dpx, m)); CK(dalloc(dpy, m));
  CK(cudaMemcpyAsync(dxy, xy, size_t(2) * m * 8, cudaMemcpyHostToDevice, w->stream));
  k_sdf_lookup<<<blocks_for(m, 128), 128, 0, w->stream>>>(w->s, m, dxy, dpx, dpy, dv);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaMemcpyAsync(px, dpx, size_t(m) * 4, cudaMemcpyDeviceToHost, w->stream));
  CK(cudaMemcpyAsync(py, dpy, size_t(m) * 4, cudaMemcpyDeviceToHost, w->stream));
  CK(cudaMemcpyAsync(value, dv, size_t(m) * 8, cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(dxy); cudaFree(dv); cudaFree(dpx); cudaFree(dpy);
  return 0;
}

int gbp_world_set_iterate_path(gbp_world_t *w0, int32_t general_only) {
  if (!w0) return fail(GBP_ERR_BAD_HANDLE, "null world");
  for (gbp_world *w : w0->grp->members) {
    w->general_only = general_only != 0;
    w->fused_tick = general_only == 2;
    // k_iterate does not maintain Store::mode while it runs alone: every robot re-qualifies from scratch
    if (w->s.Nloc > 0) {
      CK(cudaSetDevice(w->device));
      CK(cudaMemsetAsync(w->s.mode, 1, size_t(w->s.Nloc), w->stream));
    }
  }
  return 0;
}

int gbp_world_read_iterate_path(gbp_world_t *w, int64_t *robots_axis, int64_t *robots_general) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  std::vector<uint8_t> h(size_t(n), 1);
  if (n > 0) {
    CK(cudaMemcpyAsync(h.data(), w->s.mode, size_t(n), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
  }
  int64_t gen = 0;
  for (uint8_t m : h) gen += m != 0;
  if (robots_axis) *robots_axis = w->general_only ? 0 : n - gen;
  if (robots_general) *robots_general = w->general_only ? n : gen;
  return 0;
}

int gbp_world_node_counts(gbp_world_t *w, int64_t out[5]) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  const int64_t n = int64_t(w->s.Nloc) - w->n_gone, V = w->s.V;  // despawned robots took their graphs with them
  out[0] = n * V;
  out[1] = n * (V - 1);
  out[2] = n * (V - 2);
  out[3] = n * (V - 2);
  out[4] = w->s.E * (V - 1);
  return 0;
}

int64_t gbp_world_kernel_launches(const gbp_world_t *w) { return w ? w->launches : 0; }

int gbp_world_set_profiling(gbp_world_t *w, int32_t on) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (int rc = drain_profile(w)) return rc;
  w->profiling = on != 0;
  if (on) {
    for (int k = 0; k < GBP_PROFILE_KINDS; ++k) {
      w->prof_ms[k] = 0;
      w->prof_count[k] = 0;
    }
  }
  return 0;
}
int gbp_world_read_profile(gbp_world_t *w, int32_t kind, int64_t *count, double *total_ms) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (kind < 0 || kind >= GBP_PROFILE_KINDS || !count || !total_ms) return fail(GBP_ERR_BAD_ARGUMENT, "bad profile kind");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (int rc = drain_profile(w)) return rc;
  *count = w->prof_count[kind];
  *total_ms = w->prof_ms[kind];
  return 0;
}

void *gbp_host_alloc_pinned(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(GBP_ERR_CUDA, "gbp_host_alloc_pinned: cudaMallocHost failed");
    return nullptr;
  }
  return p;
}
void gbp_host_free_pinned(void *p) {
  if (p) cudaFreeHost(p);
}

int gbp_world_sync(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}
int gbp_world_timer_start(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaEventRecord(w->ev0, w->stream));
  return 0;
}
int gbp_world_timer_stop_ms(gbp_world_t *w, float *ms) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaEventRecord(w->ev1, w->stream));
  CK(cudaEventSynchronize(w->ev1));
  CK(cudaEventElapsedTime(ms, w->ev0, w->ev1));
  return 0;
}

}  // extern "C"
