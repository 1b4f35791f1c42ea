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
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "../../include/gbp_b200.h"
#include "gbp_iterate.cuh"
#include "gbp_math.cuh"
#include "gbp_store.cuh"
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
// first `used` entries of each plane.
template <class T>
cudaError_t regrow(T *&p, int planes, int64_t old_stride, int64_t new_stride, int64_t used,
                   cudaStream_t st) {
  T *q = nullptr;
  cudaError_t e = dalloc(q, size_t(planes) * size_t(new_stride));
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(q, 0, size_t(planes) * size_t(new_stride) * sizeof(T), st);
  if (e != cudaSuccess) return e;
  if (p && used > 0) {
    e = cudaMemcpy2DAsync(q, size_t(new_stride) * sizeof(T), p, size_t(old_stride) * sizeof(T),
                          size_t(used) * sizeof(T), size_t(planes), cudaMemcpyDeviceToDevice, st);
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
// expects mu (rows 20..23 of pub[p]) and prior_lam already uploaded.
__global__ void k_init_vars(Store s, int p, int64_t first, int64_t count) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int64_t vi = first + t, NV = s.NV;
  double mu[4], lam[16], cov[16];
  const double pl = s.prior_lam[vi];
#pragma unroll
  for (int k = 0; k < 4; ++k) mu[k] = s.pub[p][(20 + k) * NV + vi];
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
    for (int k = 0; k < 4; ++k) rec[k * NV + vi] = pl * mu[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) rec[(4 + k) * NV + vi] = lam[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) rec[(20 + k) * NV + vi] = mu[k];
    s.pub_epoch[b][vi] = 0u;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s.prior_eta[k * NV + vi] = pl * mu[k];
    s.bel_ext[k * NV + vi] = pl * mu[k];
    s.bel_ext[(20 + k) * NV + vi] = mu[k];
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    s.bel_ext[(4 + k) * NV + vi] = lam[k];
    s.cov[k * NV + vi] = cov[k];
  }
  s.valid[vi] = fin ? 1 : 0;
  s.mu_ext[vi] = mu[0];
  s.mu_ext[NV + vi] = mu[1];
  s.m_dynL[vi] = gbp::empty_marker();
  s.m_dynR[vi] = gbp::empty_marker();
  s.m_obs[vi] = gbp::empty_marker();
  s.m_trk[vi] = gbp::empty_marker();
  s.trk_record[vi] = 0u;
  s.trk_timeout[vi] = -1;
  s.trk_last[vi] = float(mu[0]);  // with_last_measurement (factor/mod.rs:279-283)
  s.trk_last[NV + vi] = float(mu[1]);
  s.trk_value[vi] = 0.0;
}

// VariableNode::change_prior + FactorGraph::change_prior_of_variable for variable
// `var` of robot r with new mean nm (variable.rs:203-230, factorgraph.rs:494-528):
// every factor that holds a message from this variable now holds
// (eta_belief, Lambda_belief, new mean); the variable's inbox is emptied.
__device__ void change_prior_dev(const Store &s, int p, uint32_t epoch, int64_t r, int var,
                                 const double (&nm)[4]) {
  const int64_t NV = s.NV, vi = r * s.V + var;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const double pl = s.prior_lam[vi];
  double rec[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) rec[k] = src[k * NV + vi];
#pragma unroll
  for (int k = 0; k < 4; ++k) s.prior_eta[k * NV + vi] = pl * nm[k];
  double *dst[2] = {s.pub[p], s.bel_ext};
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int k = 0; k < 20; ++k) dst[b][k * NV + vi] = rec[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[b][(20 + k) * NV + vi] = nm[k];
  }
  s.pub_epoch[p][vi] = epoch;
  s.mu_ext[vi] = nm[0];
  s.mu_ext[NV + vi] = nm[1];
  s.m_dynL[vi] = gbp::empty_marker();
  s.m_dynR[vi] = gbp::empty_marker();
  s.m_obs[vi] = gbp::empty_marker();
  s.m_trk[vi] = gbp::empty_marker();
  if (var >= 1 && s.eoff)
    for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
      const int64_t m = e * (s.V - 1) + (var - 1);
      s.mir[m] = gbp::empty_marker();
      // external factors receive the new mean whatever the antenna state (robot.rs:2272-2282)
      s.mu_frozen[m] = nm[0];
      s.mu_frozen[s.EV + m] = nm[1];
    }
}

// update_prior_of_horizon_state (planner/robot.rs:2182-2283), one thread per robot.
__global__ void k_prior_horizon(Store s, int p, uint32_t epoch, double delta_t, double max_speed,
                                int iterations_internal) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  if (s.finished[r] || s.idle[r]) return;
  const int32_t nwp = s.wp_off[r + 1] - s.wp_off[r], k = s.next_wp[r];
  if (k < 0 || k >= nwp) {
    s.finished[r] = 1;
    return;
  }
  if (iterations_internal == 0) return;
  const int64_t NV = s.NV, vi = r * s.V + (s.V - 1);
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const double ex = src[20 * NV + vi], ey = src[21 * NV + vi];
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
  change_prior_dev(s, p, epoch, r, s.V - 1, nm);
}

// update_prior_of_current_state_v3 (planner/robot.rs:2286-2338).
__global__ void k_prior_current(Store s, int p, uint32_t epoch, float delta_t) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  if (s.idle[r]) return;
  const int64_t NV = s.NV, v0 = r * s.V, v1 = v0 + 1;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  const float time_scale = __fdiv_rn(delta_t, s.t0[r]);
  double ch[4], nm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double c = src[(20 + k) * NV + v0];
    ch[k] = double(time_scale) * (src[(20 + k) * NV + v1] - c);
    nm[k] = c + ch[k];
  }
  change_prior_dev(s, p, epoch, r, 0, nm);
  s.pos[r] = __fadd_rn(s.pos[r], float(ch[0]));
  s.pos[s.cap + r] = __fadd_rn(s.pos[s.cap + r], float(ch[1]));
}

__global__ void k_change_prior_list(Store s, int p, uint32_t epoch, int var, int m,
                                    const int32_t *robots, const double *means) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const double nm[4] = {means[4 * t], means[4 * t + 1], means[4 * t + 2], means[4 * t + 3]};
  change_prior_dev(s, p, epoch, robots[t], var, nm);
}

// Gather VariableBelief of every variable into the ABI's array-of-structs layout.
__global__ void k_gather_beliefs(Store s, int p, double *eta, double *lam, double *mean, double *cov,
                                 uint8_t *valid) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= int64_t(s.Nloc) * s.V) return;
  const int64_t r = t / s.V, NV = s.NV;
  const double *src = s.latest[r] ? s.bel_ext : s.pub[p];
  if (eta)
    for (int k = 0; k < 4; ++k) eta[4 * t + k] = src[k * NV + t];
  if (lam)
    for (int k = 0; k < 16; ++k) lam[16 * t + k] = src[(4 + k) * NV + t];
  if (mean)
    for (int k = 0; k < 4; ++k) mean[4 * t + k] = src[(20 + k) * NV + t];
  if (cov)
    for (int k = 0; k < 16; ++k) cov[16 * t + k] = s.cov[k * NV + t];
  if (valid) valid[t] = s.valid[t];
}

__global__ void k_sdf_lookup(Store s, int m, const double *xy, uint32_t *px, uint32_t *py, double *val) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  val[t] = gbp::sdf_measure(s, xy[2 * t], xy[2 * t + 1], px + t, py + t);
}

__global__ void k_set_dsafe(Store s, double mult) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) s.e_dsafe[e] = mult * double(s.radius[s.enbr[e]]);
}

}  // namespace

// ---------------------------------------------------------------------------
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
    int32_t *enbr = nullptr;
    double *e_dsafe = nullptr;
    uint64_t *e_rnum = nullptr;
    uint32_t *e_birth = nullptr;
    uint8_t *e_frozen = nullptr;
    double *mir = nullptr;
    double *mu_frozen = nullptr;
    int64_t *map = nullptr;
    int64_t cap = 0;
  } edges[2];
  int cur = 0;
  int32_t *t_nlow = nullptr;
  int64_t *t_result_dev = nullptr, *t_result_host = nullptr;
  bool pending_internal_factor = false, pending_external_factor = false;
  // topology scratch
  int32_t *t_cx = nullptr, *t_cz = nullptr, *t_idx = nullptr, *t_idx_sorted = nullptr;
  uint32_t *t_keys = nullptr, *t_keys_sorted = nullptr;
  int64_t *t_cnt = nullptr, *t_off = nullptr, *t_newcnt = nullptr, *t_newoff = nullptr;
  void *t_cub = nullptr;
  size_t t_cub_bytes = 0;
  int64_t t_cap = 0;
  // grow-only device scratch for read-backs / small uploads (no cudaMalloc per call)
  void *rb_dev = nullptr;
  size_t rb_bytes = 0;
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
};

namespace {

int set_device(gbp_world *w) {
  CK(cudaSetDevice(w->device));
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
  gbp_world::Span sp{};
  ProfileScope(gbp_world *w_, int kind) : w(w_) {
    if (!w->profiling) return;
    sp.kind = kind;
    sp.a = take_event(w);
    sp.b = take_event(w);
    cudaEventRecord(sp.a, w->stream);
  }
  ~ProfileScope() {
    if (!w->profiling) return;
    cudaEventRecord(sp.b, w->stream);
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

template <bool EXT, bool INT>
int launch_iterate(gbp_world *w) {
  Store &s = w->s;
  if (s.Nloc == 0) return 0;
  const int rpw = 32 / s.V;
  const int64_t warps = (int64_t(s.Nloc) + rpw - 1) / rpw;
  const int wpb = gbp::kIterBlock / 32;
  const unsigned grid = unsigned((warps + wpb - 1) / wpb);
  w->epoch += 1;
  {
    ProfileScope ps(w, EXT ? (INT ? GBP_PROFILE_ITERATE_EXT_INT : GBP_PROFILE_ITERATE_EXT) : GBP_PROFILE_ITERATE_INT);
    gbp::k_iterate<EXT, INT><<<grid, gbp::kIterBlock, 0, w->stream>>>(s, w->p, w->epoch);
  }
  CK(cudaGetLastError());
  if (w->spans.size() > 4096) {
    if (int rc = drain_profile(w)) return rc;
  }
  w->launches += 1;
  if (INT) w->p ^= 1;
  return 0;
}

// iterate_gbp_v2 (robot.rs:1769-1861): flatten the schedule into half-steps
// I (internal factor+variable) and E (external factor+variable); an E directly
// followed by an I runs as one fused launch.
int run_schedule(gbp_world *w, int n, const uint8_t *internal, const uint8_t *external) {
  std::vector<char> ph;
  ph.reserve(size_t(n) * 2);
  for (int i = 0; i < n; ++i) {
    if (internal[i]) ph.push_back('I');
    if (external[i]) ph.push_back('E');
  }
  for (size_t k = 0; k < ph.size();) {
    int rc;
    if (ph[k] == 'E' && k + 1 < ph.size() && ph[k + 1] == 'I') {
      rc = launch_iterate<true, true>(w);
      k += 2;
    } else if (ph[k] == 'E') {
      rc = launch_iterate<true, false>(w);
      k += 1;
    } else {
      rc = launch_iterate<false, true>(w);
      k += 1;
    }
    if (rc) return rc;
  }
  return 0;
}

using EdgeSet = gbp_world::EdgeSet;

int ensure_scratch(gbp_world *w, size_t bytes) {
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
  s.mir = e.mir;
  s.mu_frozen = e.mu_frozen;
  s.EV = e.cap * (s.V - 1);
}

int grow_edge_set(gbp_world *w, EdgeSet *e, int64_t cap) {
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(e->enbr); cudaFree(e->e_dsafe); cudaFree(e->e_rnum); cudaFree(e->e_birth); cudaFree(e->e_frozen);
  cudaFree(e->mir); cudaFree(e->map); cudaFree(e->mu_frozen);
  const int Vm1 = w->s.V - 1;
  CK(dalloc(e->enbr, size_t(cap)));
  CK(dalloc(e->e_dsafe, size_t(cap)));
  CK(dalloc(e->e_rnum, size_t(cap)));
  CK(dalloc(e->e_birth, size_t(cap)));
  CK(dalloc(e->e_frozen, size_t(cap)));
  CK(dalloc(e->mu_frozen, size_t(2) * size_t(cap) * Vm1));
  CK(dalloc(e->map, size_t(cap)));
  CK(dalloc(e->mir, size_t(6) * size_t(cap) * Vm1));
  e->cap = cap;
  return 0;
}

int ensure_topology_scratch(gbp_world *w, int64_t n) {
  if (!w->t_result_dev) {
    CK(dalloc(w->t_result_dev, 2));
    CK(cudaMallocHost(reinterpret_cast<void **>(&w->t_result_host), 2 * sizeof(int64_t)));
  }
  if (n <= w->t_cap) return 0;
  cudaFree(w->t_nlow);
  CK(dalloc(w->t_nlow, n));
  cudaFree(w->t_cx); cudaFree(w->t_cz); cudaFree(w->t_idx); cudaFree(w->t_idx_sorted);
  cudaFree(w->t_keys); cudaFree(w->t_keys_sorted); cudaFree(w->t_cnt); cudaFree(w->t_off);
  cudaFree(w->t_newcnt); cudaFree(w->t_newoff); cudaFree(w->t_cub);
  CK(dalloc(w->t_cx, n)); CK(dalloc(w->t_cz, n)); CK(dalloc(w->t_idx, n)); CK(dalloc(w->t_idx_sorted, n));
  CK(dalloc(w->t_keys, n)); CK(dalloc(w->t_keys_sorted, n));
  CK(dalloc(w->t_cnt, n + 1)); CK(dalloc(w->t_off, n + 1));
  CK(dalloc(w->t_newcnt, n + 1)); CK(dalloc(w->t_newoff, n + 1));
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, w->t_keys, w->t_keys_sorted, w->t_idx, w->t_idx_sorted, int(n));
  cub::DeviceScan::ExclusiveSum(nullptr, b2, w->t_cnt, w->t_off, int(n + 1));
  w->t_cub_bytes = std::max(b1, b2);
  CK(cudaMalloc(&w->t_cub, w->t_cub_bytes));
  w->t_cap = n;
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

gbp_world_t *gbp_world_create(const gbp_config_t *cfg, int32_t device) {
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
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&w->ev0) != cudaSuccess || cudaEventCreate(&w->ev1) != cudaSuccess) {
    fail(GBP_ERR_CUDA, "gbp_world_create: stream/event creation failed");
    delete w;
    return nullptr;
  }
  w->s.V = cfg->num_variables;
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
  refresh_scalars(w);
  return w;
}

void gbp_world_destroy(gbp_world_t *w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->stream);
  Store &s = w->s;
  void *ptrs[] = {s.prior_eta, s.prior_lam, s.pub[0], s.pub[1], s.pub_epoch[0], s.pub_epoch[1], s.bel_ext,
                  s.mu_ext, s.cov, s.valid, s.m_dynL, s.m_dynR, s.m_obs, s.m_trk, s.dyn_dt,
                  s.trk_record, s.trk_timeout, s.trk_last, s.trk_value, s.radius, s.t0, s.pos, s.antenna,
                  s.idle, s.finished, s.latest, s.iter_factor, s.gid, s.next_wp, s.wp_off, s.wp_xy, s.eoff,
                  s.nlow, w->edges[0].enbr, w->edges[0].e_dsafe, w->edges[0].e_rnum, w->edges[0].e_birth,
                  w->edges[0].e_frozen, w->edges[0].mir, w->edges[0].map, w->edges[0].mu_frozen, w->edges[1].enbr,
                  w->edges[1].e_dsafe, w->edges[1].e_rnum, w->edges[1].e_birth, w->edges[1].e_frozen,
                  w->edges[1].mir, w->edges[1].map, w->edges[1].mu_frozen,
                  w->t_nlow, w->t_result_dev, w->sdf_dev, w->t_cx,
                  w->t_cz, w->t_idx, w->t_idx_sorted, w->t_keys, w->t_keys_sorted, w->t_cnt, w->t_off,
                  w->t_newcnt, w->t_newoff, w->t_cub, w->rb_dev};
  for (void *q : ptrs) cudaFree(q);
  for (auto &sp : w->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (cudaEvent_t e : w->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(w->ev0);
  cudaEventDestroy(w->ev1);
  cudaStreamDestroy(w->stream);
  if (w->t_result_host) cudaFreeHost(w->t_result_host);
  delete w;
}

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
  const int64_t N0 = s.N, N1 = int64_t(s.N) + n, oldNV = s.NV, newNV = N1 * V, used = N0 * V;
  const int64_t oldcap = s.cap;
  // ---- re-stride every plane to the new capacity
  CK(regrow(s.prior_eta, 4, oldNV, newNV, used, st));
  CK(regrow(s.prior_lam, 1, oldNV, newNV, used, st));
  CK(regrow(s.pub[0], gbp::kRec, oldNV, newNV, used, st));
  CK(regrow(s.pub[1], gbp::kRec, oldNV, newNV, used, st));
  CK(regrow(s.pub_epoch[0], 1, oldNV, newNV, used, st));
  CK(regrow(s.pub_epoch[1], 1, oldNV, newNV, used, st));
  CK(regrow(s.bel_ext, gbp::kRec, oldNV, newNV, used, st));
  CK(regrow(s.mu_ext, 2, oldNV, newNV, used, st));
  CK(regrow(s.cov, 16, oldNV, newNV, used, st));
  CK(regrow(s.valid, 1, oldNV, newNV, used, st));
  CK(regrow(s.m_dynL, 20, oldNV, newNV, used, st));
  CK(regrow(s.m_dynR, 20, oldNV, newNV, used, st));
  CK(regrow(s.m_obs, 4, oldNV, newNV, used, st));
  CK(regrow(s.m_trk, 3, oldNV, newNV, used, st));
  CK(regrow(s.dyn_dt, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_record, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_timeout, 1, oldNV, newNV, used, st));
  CK(regrow(s.trk_last, 2, oldNV, newNV, used, st));
  CK(regrow(s.trk_value, 1, oldNV, newNV, used, st));
  CK(regrow(s.radius, 1, oldcap, N1, N0, st));
  CK(regrow(s.t0, 1, oldcap, N1, N0, st));
  CK(regrow(s.pos, 2, oldcap, N1, N0, st));
  CK(regrow(s.antenna, 1, oldcap, N1, N0, st));
  CK(regrow(s.idle, 1, oldcap, N1, N0, st));
  CK(regrow(s.finished, 1, oldcap, N1, N0, st));
  CK(regrow(s.latest, 1, oldcap, N1, N0, st));
  CK(regrow(s.iter_factor, 1, oldcap, N1, N0, st));
  CK(regrow(s.gid, 1, oldcap, N1, N0, st));
  CK(regrow(s.next_wp, 1, oldcap, N1, N0, st));
  {  // eoff / nlow: new robots start without edges
    int64_t *eoff = nullptr;
    CK(dalloc(eoff, N1 + 1));
    std::vector<int64_t> h(size_t(N1) + 1, s.E);
    if (s.eoff && N0 > 0) CK(cudaMemcpy(h.data(), s.eoff, size_t(N0 + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    else h[0] = 0;
    for (int64_t k = N0 + 1; k <= N1; ++k) h[k] = s.E;
    CK(cudaMemcpy(eoff, h.data(), h.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    cudaFree(s.eoff);
    s.eoff = eoff;
    CK(regrow(s.nlow, 1, oldcap, N1, N0, st));
  }
  s.NV = newNV;
  s.cap = N1;
  s.N = int32_t(N1);
  s.Nloc = int32_t(N1);

  // ---- host staging of the new robots (RobotBundle::new, robot.rs:1134-1355)
  const int64_t nv = int64_t(n) * V;
  std::vector<double> mu(size_t(4) * nv), pl(nv), dt(nv, 0.0);
  std::vector<float> rad(radii, radii + n), t0(n), pos(size_t(2) * n);
  std::vector<uint8_t> ones(n, 1);
  std::vector<int32_t> gid(n), nwp(n, 1);
  for (int r = 0; r < n; ++r) {
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
  CK(upload_planes(s.pub[w->p] + 20 * newNV, newNV, used, mu.data(), 4, nv, st));
  CK(upload_planes(s.prior_lam, newNV, used, pl.data(), 1, nv, st));
  CK(upload_planes(s.dyn_dt, newNV, used, dt.data(), 1, nv, st));
  CK(upload_planes(s.radius, N1, N0, rad.data(), 1, n, st));
  CK(upload_planes(s.t0, N1, N0, t0.data(), 1, n, st));
  CK(upload_planes(s.pos, N1, N0, pos.data(), 2, n, st));
  CK(upload_planes(s.antenna, N1, N0, ones.data(), 1, n, st));
  CK(upload_planes(s.gid, N1, N0, gid.data(), 1, n, st));
  CK(upload_planes(s.next_wp, N1, N0, nwp.data(), 1, n, st));
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
  k_init_vars<<<blocks_for(nv, 256), 256, 0, st>>>(s, w->p, used, nv);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(st));  // staging vectors go out of scope
  return 0;
}

int32_t gbp_world_num_robots(const gbp_world_t *w) { return w ? w->s.Nloc : 0; }

int gbp_world_update_topology(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  Store &s = w->s;
  const int32_t n = s.Nloc;
  if (n == 0) return 0;
  cudaStream_t st = w->stream;
  if (int rc = ensure_topology_scratch(w, n)) return rc;
  ProfileScope ps(w, GBP_PROFILE_TOPOLOGY);
  const int T = 128;
  const float R = w->cfg.comms_radius;
  const double cell = double(R) * 1.001;
  gbp::k_cell_keys<<<blocks_for(n, T), T, 0, st>>>(n, s.pos, s.pos + s.cap, cell, w->t_cx, w->t_cz, w->t_keys, w->t_idx);
  size_t cb = w->t_cub_bytes;
  CK(cub::DeviceRadixSort::SortPairs(w->t_cub, cb, w->t_keys, w->t_keys_sorted, w->t_idx, w->t_idx_sorted, n, 0, 32, st));
  gbp::k_neighbours<false><<<blocks_for(n, T), T, 0, st>>>(n, 0, n, s.pos, s.pos + s.cap, w->t_cx, w->t_cz,
                                                           w->t_keys_sorted, w->t_idx_sorted, R, w->t_cnt, nullptr, 0);
  CK(cudaMemsetAsync(w->t_cnt + n, 0, sizeof(int64_t), st));
  cb = w->t_cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_cnt, w->t_off, n + 1, st));
  w->launches += 4;
  // The new CSR is written straight into the spare edge set; one host sync per
  // tick reads (edge count, new edge count).  If the spare set is too small the
  // guarded kernels did nothing: grow it and run them again.
  EdgeSet *spare = &w->edges[1 - w->cur];
  int64_t E1 = 0, total_new = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    gbp::k_neighbours<true><<<blocks_for(n, T), T, 0, st>>>(n, 0, n, s.pos, s.pos + s.cap, w->t_cx, w->t_cz,
                                                            w->t_keys_sorted, w->t_idx_sorted, R, w->t_off,
                                                            spare->enbr, spare->cap);
    gbp::k_edge_diff<<<blocks_for(n, T), T, 0, st>>>(n, w->t_off, spare->enbr, s.eoff, s.enbr, n, spare->map,
                                                     w->t_newcnt, w->t_nlow, spare->cap);
    CK(cudaMemsetAsync(w->t_newcnt + n, 0, sizeof(int64_t), st));
    cb = w->t_cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(w->t_cub, cb, w->t_newcnt, w->t_newoff, n + 1, st));
    gbp::k_topology_result<<<1, 1, 0, st>>>(w->t_off, w->t_newoff, n, w->t_result_dev);
    CK(cudaMemcpyAsync(w->t_result_host, w->t_result_dev, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    w->launches += 4;
    E1 = w->t_result_host[0];
    total_new = w->t_result_host[1];
    if (E1 <= spare->cap) break;
    if (int rc = grow_edge_set(w, spare, E1 + E1 / 4 + 1024)) return rc;
  }
  if (total_new == 0 && E1 == s.E) return 0;  // connectivity unchanged: keep the store as is
  w->epoch += 1;
  const int Vm1 = s.V - 1;
  gbp::k_edge_assign<<<blocks_for(n, T), T, 0, st>>>(n, s.V, w->t_off, spare->enbr, spare->map, w->t_newoff, s.radius,
                                                     double(w->cfg.safety_distance_multiplier), w->robot_number,
                                                     w->epoch, s.e_dsafe, s.e_rnum, s.e_birth, s.e_frozen,
                                                     spare->e_dsafe, spare->e_rnum, spare->e_birth, spare->e_frozen);
  if (E1 > 0)
    gbp::k_mirror_move<<<blocks_for(E1 * Vm1, 256), 256, 0, st>>>(s, w->p, E1 * Vm1, Vm1, w->t_off, spare->map, s.mir,
                                                                  s.mu_frozen, s.EV, spare->mir, spare->mu_frozen,
                                                                  spare->cap * Vm1);
  CK(cudaGetLastError());
  w->launches += 3;
  CK(cudaMemcpyAsync(s.eoff, w->t_off, size_t(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(s.nlow, w->t_nlow, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  w->cur = 1 - w->cur;
  bind_edge_set(w);
  s.E = E1;
  w->robot_number += uint64_t(Vm1) * uint64_t(total_new);
  return 0;
}

int gbp_world_set_comms(gbp_world_t *w, const uint8_t *antenna_active, const uint8_t *idle) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  const int n = w->s.Nloc;
  if (n == 0) return 0;
  if (antenna_active) CK(cudaMemcpyAsync(w->s.antenna, antenna_active, n, cudaMemcpyHostToDevice, w->stream));
  else CK(cudaMemsetAsync(w->s.antenna, 1, n, w->stream));
  if (idle) CK(cudaMemcpyAsync(w->s.idle, idle, n, cudaMemcpyHostToDevice, w->stream));
  else CK(cudaMemsetAsync(w->s.idle, 0, n, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

int gbp_world_set_waypoint_index(gbp_world_t *w, const int32_t *next_index) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!next_index) return fail(GBP_ERR_BAD_ARGUMENT, "null index array");
  if (set_device(w)) return GBP_ERR_CUDA;
  CK(cudaMemcpyAsync(w->s.next_wp, next_index, size_t(w->s.Nloc) * sizeof(int32_t), cudaMemcpyHostToDevice, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
}

int gbp_world_update_prior_of_horizon_state(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (w->s.Nloc == 0) return 0;
  w->epoch += 1;
  ProfileScope ps(w, GBP_PROFILE_PRIORS);
  k_prior_horizon<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(
      w->s, w->p, w->epoch, double(w->cfg.delta_t), double(w->cfg.target_speed), w->cfg.iterations_internal);
  CK(cudaGetLastError());
  w->launches += 1;
  return 0;
}

int gbp_world_update_prior_of_current_state(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
  if (w->s.Nloc == 0) return 0;
  w->epoch += 1;
  ProfileScope ps(w, GBP_PROFILE_PRIORS);
  k_prior_current<<<blocks_for(w->s.Nloc, 128), 128, 0, w->stream>>>(w->s, w->p, w->epoch, w->cfg.delta_t);
  CK(cudaGetLastError());
  w->launches += 1;
  return 0;
}

int gbp_world_change_prior_of_variable(gbp_world_t *w, int32_t var, int32_t m, const int32_t *robots,
                                       const double *new_means) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
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
  k_change_prior_list<<<blocks_for(m, 128), 128, 0, w->stream>>>(w->s, w->p, w->epoch, var, m, dr, dm);
  CK(cudaGetLastError());
  w->launches += 1;
  CK(cudaStreamSynchronize(w->stream));
  cudaFree(dr);
  cudaFree(dm);
  return 0;
}

int gbp_world_iterate_schedule(gbp_world_t *w, int32_t n, const uint8_t *internal, const uint8_t *external) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (n < 0 || (n > 0 && (!internal || !external))) return fail(GBP_ERR_BAD_ARGUMENT, "bad schedule");
  if (w->pending_internal_factor || w->pending_external_factor)
    return fail(GBP_ERR_STATE, "a half-iteration pair is open");
  if (set_device(w)) return GBP_ERR_CUDA;
  return run_schedule(w, n, internal, external);
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
  if (w->pending_internal_factor || w->pending_external_factor) return fail(GBP_ERR_STATE, "half-iteration order");
  w->pending_internal_factor = true;
  return 0;
}
int gbp_world_internal_variable_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!w->pending_internal_factor) return fail(GBP_ERR_STATE, "internal_variable_iteration without internal_factor_iteration");
  w->pending_internal_factor = false;
  if (set_device(w)) return GBP_ERR_CUDA;
  return launch_iterate<false, true>(w);
}
int gbp_world_external_factor_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (w->pending_internal_factor || w->pending_external_factor) return fail(GBP_ERR_STATE, "half-iteration order");
  w->pending_external_factor = true;
  return 0;
}
int gbp_world_external_variable_iteration(gbp_world_t *w) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (!w->pending_external_factor) return fail(GBP_ERR_STATE, "external_variable_iteration without external_factor_iteration");
  w->pending_external_factor = false;
  if (set_device(w)) return GBP_ERR_CUDA;
  return launch_iterate<true, false>(w);
}

int gbp_world_step(gbp_world_t *w) {
  int rc;
  if ((rc = gbp_world_update_topology(w))) return rc;
  if ((rc = gbp_world_update_prior_of_horizon_state(w))) return rc;
  if ((rc = gbp_world_update_prior_of_current_state(w))) return rc;
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

int gbp_world_read_beliefs(gbp_world_t *w, double *eta, double *lam, double *mean, double *cov, uint8_t *valid) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  if (set_device(w)) return GBP_ERR_CUDA;
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
  if (eta) CK(cudaMemcpyAsync(eta, d_eta, b_eta, cudaMemcpyDeviceToHost, w->stream));
  if (lam) CK(cudaMemcpyAsync(lam, d_lam, b_lam, cudaMemcpyDeviceToHost, w->stream));
  if (mean) CK(cudaMemcpyAsync(mean, d_mean, b_mean, cudaMemcpyDeviceToHost, w->stream));
  if (cov) CK(cudaMemcpyAsync(cov, d_cov, b_cov, cudaMemcpyDeviceToHost, w->stream));
  if (valid) CK(cudaMemcpyAsync(valid, d_valid, b_valid, cudaMemcpyDeviceToHost, w->stream));
  CK(cudaStreamSynchronize(w->stream));
  return 0;
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
  if (s.Nloc == 0) {
    offsets[0] = 0;
    return 0;
  }
  CK(cudaMemcpy(offsets, s.eoff, size_t(s.Nloc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (s.E > 0) {
    if (neighbours) CK(cudaMemcpy(neighbours, s.enbr, size_t(s.E) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (robot_number) {
      // the store keeps, per (receiver r <- a), the number of a's factor toward r;
      // the ABI reports per (owner r -> a): swap through the symmetric edge
      std::vector<int32_t> nb(s.E);
      std::vector<uint64_t> rn(s.E);
      std::vector<int64_t> off(size_t(s.Nloc) + 1);
      CK(cudaMemcpy(nb.data(), s.enbr, size_t(s.E) * sizeof(int32_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(rn.data(), s.e_rnum, size_t(s.E) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(off.data(), s.eoff, off.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
      for (int32_t r = 0; r < s.Nloc; ++r)
        for (int64_t e = off[r]; e < off[r + 1]; ++e) {
          const int32_t a = nb[e];
          const int32_t *lo = nb.data() + off[a], *hi = nb.data() + off[a + 1];
          const int32_t *it = std::lower_bound(lo, hi, r);
          robot_number[e] = (it != hi && *it == r) ? int64_t(rn[it - nb.data()]) : -1;
        }
    }
  }
  return s.E;
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

int gbp_world_node_counts(gbp_world_t *w, int64_t out[5]) {
  if (!w) return fail(GBP_ERR_BAD_HANDLE, "null world");
  const int64_t n = w->s.Nloc, V = w->s.V;
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
