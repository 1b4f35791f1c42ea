// gbp_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A deliberately literal, single-source CPU restatement of the GBP hot path of
// AU-Master-Thesis/magics (Rust), used only as the parity checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Nothing under magics_b200/ may include, link or call this file.
//
// The reference cannot be compiled here (no Rust toolchain, ~600 crates, nightly),
// so this follows the Rust source function by function; every function cites the
// file:line it restates (paths relative to /root/reference/crates/magics/src
// unless another crate is named).  It keeps the reference's data-structure
// choices on purpose: one graph object per robot, heap-allocated dynamically
// sized vectors/matrices (ndarray stand-ins), ordered-map inboxes keyed by
// (graph id, node index) (BTreeMap stand-ins), boxed optional message payloads.
//
// PARITY PINNING.  Pinned by the reference's own tests: marginalise block
// extraction / single-neighbour pass-through (marginalise_factor_distance.rs:182-233),
// get_variable_timesteps (utils.rs:95-133), the gbp_schedule sequence tests.
// PARITY UNPINNED for FactorNode::update, the factors' measure/jacobian, the
// belief update and a whole GBP iterate: the reference ships no test or golden
// vector for them, and the 4x4 `.inv()` comes from the un-vendored third-party
// crate ndarray-inverse 0.1.9 (Cargo.lock:4870-4873).  Its published algorithm
// (determinant by explicit expansion; inverse = adjugate / det; None iff det == 0)
// is restated in `inv()` below.  ndarray 0.15.6 `.dot` is restated as: mat*mat =
// k-ascending accumulation from 0 (matrixmultiply), mat*vec = per-row
// `unrolled_dot` (ndarray numeric_util.rs: eight partial sums combined
// (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail sequentially).
//
// SDF generation (crates/env_to_png): image_to_tile_coords and one is_tile_obstacle case
// are pinned by the crate's #[test]s (the other three assertions there are stale against the crate's own
// code, tests/golden/make_golden.py records which); the placeable shapes and `image::imageops::blur`
// (image 0.25.1, third party) are PARITY UNPINNED and checked against independent geometry / a scipy
// Gaussian instead (tests/test_oracle_env.py).  The RRT* hand-off functions (next-4) restate
// factorgraph.rs:1467-1590 literally; the reference has no test for them.
//
// Evaluation outputs: update_robot_environment_collisions (planner/collisions.rs:368-455) calls
// parry2d 0.13.7 (a git fork, Cargo.lock:5372-5374, absent) and the trackers (planner/tracking.rs:117-260) bevy_time's
// Timer (absent): PARITY UNPINNED — intersection_test / project_local_point and Timer::tick are restated from the
// crates' published algorithms and checked against float64 geometry and hand-computed timer sequences
// (tests/test_oracle_evaluation.py).
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off -pthread -shared -fPIC (see Makefile).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <atomic>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr int DOFS = 4;  // factorgraph/mod.rs:21

// Stand-in for Bevy's ComputeTaskPool behind Query::par_iter_mut (robot.rs:1789):
// dynamic chunks of `grain` items over `threads` std::threads.
template <class F>
void parallel_for(int n, int threads, int grain, F &&body) {
  if (threads <= 1 || n <= grain) {
    for (int i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (;;) {
      int b = next.fetch_add(grain);
      if (b >= n) break;
      int e = std::min(n, b + grain);
      for (int i = b; i < e; ++i) body(i);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto &t : pool) t.join();
}

// ---------------------------------------------------------------------------
// gbp_linalg stand-ins (crates/gbp_linalg/src/lib.rs:31-128): heap vectors.
// ---------------------------------------------------------------------------
using Vec = std::vector<double>;
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, 0.0) {}
  double &operator()(int i, int j) { return a[size_t(i) * c + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * c + j]; }
  static Mat eye(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
};

// ndarray numeric_util::unrolled_dot (contiguous slices).
double unrolled_dot(const double *xs, const double *ys, int len) {
  double sum = 0.0, p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  while (len >= 8) {
    for (int k = 0; k < 8; ++k) p[k] = p[k] + xs[k] * ys[k];
    xs += 8;
    ys += 8;
    len -= 8;
  }
  sum = sum + (p[0] + p[4]);
  sum = sum + (p[1] + p[5]);
  sum = sum + (p[2] + p[6]);
  sum = sum + (p[3] + p[7]);
  for (int i = 0; i < len && i < 7; ++i) sum = sum + xs[i] * ys[i];
  return sum;
}
// Array2.dot(Array2): matrixmultiply gemm, k ascending, accumulator starts at 0.
Mat matmul(const Mat &A, const Mat &B) {
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < B.c; ++j) {
      double acc = 0.0;
#ifdef GBPO_FMA_MATMUL
      // sensitivity build only (Makefile target `fma`): matrixmultiply's x86-64 dgemm kernels with the `fma` feature
      // detected at run time accumulate with fused multiply-adds; everything else stays rounded per operation.
      for (int k = 0; k < A.c; ++k) acc = std::fma(A(i, k), B(k, j), acc);
#else
      for (int k = 0; k < A.c; ++k) acc = acc + A(i, k) * B(k, j);
#endif
      C(i, j) = acc;
    }
  return C;
}
// Array2.dot(Array1): row.dot(x) per row.
Vec matvec(const Mat &A, const Vec &x) {
  Vec y(A.r);
  for (int i = 0; i < A.r; ++i) y[i] = unrolled_dot(&A.a[size_t(i) * A.c], x.data(), A.c);
  return y;
}
Mat transpose(const Mat &A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
  return T;
}
// VectorNorm::euclidean_norm (gbp_linalg/src/lib.rs:68-70): sqrt(fold(0, acc + x*x)).
double euclidean_norm(const Vec &v) {
  double acc = 0.0;
  for (double x : v) acc = acc + x * x;
  return std::sqrt(acc);
}
// NdarrayVectorExt::normalized (gbp_linalg/src/lib.rs:113-124).
Vec normalized(Vec v) {
  double mag = euclidean_norm(v);
  if (mag == 0.0 || std::isinf(mag)) return v;
  for (double &x : v) x /= mag;
  return v;
}

// ndarray-inverse 0.1.9 `Inverse::inv` (third party, not under /root/reference):
// determinant by explicit expansion, inverse = transposed cofactors / det,
// `None` iff det == 0.  Call sites: marginalise_factor_distance.rs:79,
// variable.rs:153, variable.rs:278.
double det3(double a, double b, double c, double d, double e, double f, double g, double h,
            double i) {
  return a * e * i + b * f * g + c * d * h - c * e * g - b * d * i - a * f * h;
}
double minor4(const Mat &m, int sr, int sc) {
  double s[9];
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    if (i == sr) continue;
    for (int j = 0; j < 4; ++j) {
      if (j == sc) continue;
      s[n++] = m(i, j);
    }
  }
  return det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
}
double det4(const Mat &m) {
  // Laplace expansion along row 0.
  return m(0, 0) * minor4(m, 0, 0) - m(0, 1) * minor4(m, 0, 1) + m(0, 2) * minor4(m, 0, 2) -
         m(0, 3) * minor4(m, 0, 3);
}
std::optional<Mat> inv(const Mat &m) {
  // only 4x4 occurs on the hot path (DOFS = 4)
  double det = det4(m);
  if (det == 0.0) return std::nullopt;  // NaN det is "not zero", like Float::is_zero()
  Mat out(4, 4);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double sign = ((i + j) % 2 == 0) ? 1.0 : -1.0;
      out(j, i) = (sign * minor4(m, i, j)) / det;
    }
  return out;
}

// ---------------------------------------------------------------------------
// message.rs:18-53,121-173 — Message{payload: Option<Box<Payload>>}
// ---------------------------------------------------------------------------
struct Payload {
  Vec eta;
  Mat lam;
  Vec mu;
};
struct Message {
  std::shared_ptr<Payload> payload;  // None == Message::empty()
  static Message empty() { return Message{}; }
  static Message make(Vec eta, Mat lam, Vec mu) {
    Message m;
    m.payload = std::make_shared<Payload>(Payload{std::move(eta), std::move(lam), std::move(mu)});
    return m;
  }
  bool is_empty() const { return !payload; }
};

// id.rs:25-61,83-118 — (graph id, node index), ordered by graph id then index.
using NodeId = std::pair<int, int>;
using Inbox = std::map<NodeId, Message>;  // BTreeMap (message.rs:210-217)

struct Cfg {  // mirror of gbp_config_t (include/gbp_b200.h), same field order
  int32_t num_variables;
  float sigma_factor_dynamics, sigma_factor_interrobot, sigma_factor_obstacle,
      sigma_factor_tracking;
  float safety_distance_multiplier, comms_radius, target_speed, delta_t;
  float tracking_switch_padding, tracking_attraction_distance;
  uint8_t enable_dynamic, enable_interrobot, enable_obstacle, enable_tracking;
  int32_t schedule_kind, iterations_internal, iterations_external;
  double world_width, world_height;
  int32_t strict_reference_quirks;  // 1: delete_interrobot_factors through the lossy per-robot map (robot.rs:1391-1404)
};

struct Sdf {
  int w = 0, h = 0;
  std::vector<uint8_t> rgb;
};

enum Kind { DYNAMIC = 0, INTERROBOT = 1, OBSTACLE = 2, TRACKING = 3 };

// factor/mod.rs:597-650 FactorState + the kind-specific fields of the four
// live factor kinds (dynamic.rs, interrobot.rs, obstacle.rs, tracking.rs).
// MessageCount (factorgraph/mod.rs:103-137): per node, never reset (reset_message_count has no caller).
struct MsgCount {
  uint64_t sent_int = 0, sent_ext = 0, recv_int = 0, recv_ext = 0;
};

struct Factor {
  MsgCount cnt;
  Kind kind;
  int graph;       // factorgraph_id
  int index;       // node index
  bool enabled;    // factor/mod.rs:147
  Inbox inbox;     // keyed by VariableId
  Vec z;           // initial_measurement
  Mat meas_prec;   // measurement_precision
  Vec lin;         // linearisation_point
  // dynamic.rs
  Mat cached_jacobian;
  // interrobot.rs:40-47
  double safety_distance = 0, robot_radius = 0, tiny_offset = 0;
  int ext_robot = -1, ext_var = -1;
  int own_var = -1;  // index of the (first) own-graph variable the factor is attached to
  uint64_t robot_number = 0;
  // obstacle.rs
  const Sdf *sdf = nullptr;
  double world_w = 0, world_h = 0, jac_delta = 0;
  // tracking.rs:15-28,72-79
  std::vector<std::pair<float, float>> path;
  size_t trk_index = 1, trk_record = 0;
  float last_pos[2] = {0, 0};
  double last_value = 0;
  std::optional<size_t> timeout;
  float switch_padding = 1.0f, attraction_distance = 2.0f;
  // last obstacle pixel (for bit-exact indexing checks)
  uint32_t last_px = 0, last_py = 0;
};

// FactorState::new (factor/mod.rs:626-643): eye(len(z)) / strength^2.
void init_state(Factor &f, Vec z, double strength, int neighbours) {
  int m = int(z.size());
  f.z = std::move(z);
  f.meas_prec = Mat::eye(m);
  double s2 = strength * strength;  // Float::powi(strength, 2)
  for (double &x : f.meas_prec.a) x = x / s2;
  f.lin.assign(size_t(DOFS) * neighbours, 0.0);
}

// DynamicFactor::new (factor/dynamic.rs:22-52).
void init_dynamic(Factor &f, double strength, double delta_t) {
  const int h = DOFS / 2;
  Mat eye = Mat::eye(h);
  double qs = 1.0 / (strength * strength);         // powi(strength, -2)
  double p3 = 1.0 / ((delta_t * delta_t) * delta_t);  // powi(delta_t, -3)
  double p2 = 1.0 / (delta_t * delta_t);            // powi(delta_t, -2)
  Mat qc(h, h);
  for (int i = 0; i < h * h; ++i) qc.a[i] = qs * eye.a[i];
  Mat qi(DOFS, DOFS);
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < h; ++j) {
      qi(i, j) = (12.0 * p3) * qc(i, j);
      qi(i, j + h) = (-6.0 * p2) * qc(i, j);
      qi(i + h, j) = (-6.0 * p2) * qc(i, j);
      qi(i + h, j + h) = (4.0 / delta_t) * qc(i, j);
    }
  f.meas_prec = qi;
  Mat J(DOFS, DOFS * 2);
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < h; ++j) {
      double e = eye(i, j);
      J(i, j) = e;
      J(i, j + h) = delta_t * e;
      J(i, j + 2 * h) = -1.0 * e;
      J(i, j + 3 * h) = 0.0;
      J(i + h, j) = 0.0;
      J(i + h, j + h) = e;
      J(i + h, j + 2 * h) = 0.0;
      J(i + h, j + 3 * h) = -1.0 * e;
    }
  f.cached_jacobian = J;
}

// ObstacleFactor::measure (factor/obstacle.rs:141-188).
// Rust `as u32` saturates: NaN -> 0, negative -> 0, overflow -> u32::MAX.
uint32_t sat_u32(double v) {
  if (std::isnan(v)) return 0;
  if (v <= 0.0) return 0;
  if (v >= 4294967295.0) return 4294967295u;
  return uint32_t(v);  // truncation toward zero
}
double sdf_measure(const Sdf &sdf, double world_w, double world_h, double x_pos, double y_pos,
                   uint32_t *opx, uint32_t *opy) {
  double x_offset = world_w / 2.0;
  double y_offset = world_h / 2.0;
  double x_scale = double(uint32_t(sdf.w)) / world_w;
  double y_scale = double(uint32_t(sdf.h)) / world_h;
  uint32_t x_pixel = sat_u32((x_pos + x_offset) * x_scale);
  uint32_t y_pixel = sat_u32((-y_pos + y_offset) * y_scale);
  if (opx) *opx = x_pixel;
  if (opy) *opy = y_pixel;
  // image::ImageBuffer::get_pixel_checked: x < width && y < height
  if (!(x_pixel < uint32_t(sdf.w) && y_pixel < uint32_t(sdf.h))) return 0.0;
  uint8_t red = sdf.rgb[(size_t(y_pixel) * sdf.w + x_pixel) * 3];
  return 1.0 - double(red) / 255.0;
}

// Factor::measure dispatch (factor/mod.rs:556-563).
Vec measure(Factor &f, const Vec &x) {
  switch (f.kind) {
    case DYNAMIC:  // dynamic.rs:74-76
      return matvec(f.cached_jacobian, x);
    case INTERROBOT: {  // interrobot.rs:165-204 (+ :91-107)
      Vec m(f.z.size(), 0.0);
      Vec d(2);
      for (int i = 0; i < 2; ++i) d[i] = (x[i] - x[DOFS + i]);
      for (int i = 0; i < 2; ++i) d[i] += f.tiny_offset;
      double radius = euclidean_norm(d);
      if (radius <= f.safety_distance) m[0] = 1.0 * (1.0 - radius / f.safety_distance);
      return m;
    }
    case OBSTACLE: {
      double v = sdf_measure(*f.sdf, f.world_w, f.world_h, x[0], x[1], &f.last_px, &f.last_py);
      return Vec{v};
    }
    case TRACKING: {  // tracking.rs:197-346
      size_t rec = f.trk_record;
      Vec x_pos{x[0], x[1]}, x_vel{x[2], x[3]};
      auto seg = [&](size_t k, Vec &s, Vec &e) {
        s = {double(f.path[k].first), double(f.path[k].second)};
        e = {double(f.path[k + 1].first), double(f.path[k + 1].second)};
      };
      auto dot2 = [](const Vec &a, const Vec &b) { return unrolled_dot(a.data(), b.data(), 2); };
      Vec cs, ce;
      seg(rec, cs, ce);
      Vec line{ce[0] - cs[0], ce[1] - cs[1]};
      Vec rel{x_pos[0] - cs[0], x_pos[1] - cs[1]};
      double t = dot2(rel, line) / dot2(line, line);
      Vec cur{cs[0] + t * line[0], cs[1] + t * line[1]};
      double d0 = double(f.switch_padding), d1 = d0 * 0.01;
      double cur_to_end = euclidean_norm(Vec{ce[0] - cur[0], ce[1] - cur[1]});
      bool have_prev = false;
      Vec prevp;
      if (rec > 0) {
        Vec ps, pe;
        seg(rec - 1, ps, pe);
        Vec pl{pe[0] - ps[0], pe[1] - ps[1]};
        Vec prel{x_pos[0] - ps[0], x_pos[1] - ps[1]};
        double tp = dot2(prel, pl) / dot2(pl, pl);
        Vec pp{ps[0] + tp * pl[0], ps[1] + tp * pl[1]};
        double cur_to_prev_end = euclidean_norm(Vec{pe[0] - cur[0], pe[1] - cur[1]});
        double prev_to_prev_end = euclidean_norm(Vec{cs[0] - pp[0], cs[1] - pp[1]});
        if (cur_to_prev_end < d0 && cur_to_prev_end > d1 && prev_to_prev_end < d0) {
          have_prev = true;
          prevp = pp;
        }
      }
      if (cur_to_end < d0) {  // Tracking::increment_record (tracking.rs:55-65)
        f.trk_record = std::min(f.trk_record + 1, f.path.size() - 2);
      }
      Vec mp(2);
      if (have_prev) {
        for (int i = 0; i < 2; ++i) {
          double x_to_cur = cur[i] - x_pos[i];
          double x_to_prev = prevp[i] - x_pos[i];
          mp[i] = x_pos[i] + (x_to_cur + x_to_prev);
        }
      } else {
        Vec ln = normalized(line);
        double vn = euclidean_norm(x_vel);
        for (int i = 0; i < 2; ++i) mp[i] = cur[i] + ln[i] * vn / 5.0;
      }
      double dist = euclidean_norm(Vec{mp[0] - x_pos[0], mp[1] - x_pos[1]});
      double ad = double(f.attraction_distance);
      double meas = dist < ad ? dist / ad : 1.0;
      f.last_pos[0] = float(mp[0]);
      f.last_pos[1] = float(mp[1]);
      f.last_value = meas;
      return Vec{meas};
    }
  }
  return {};
}

double jacobian_delta(const Factor &f) {
  switch (f.kind) {
    case DYNAMIC: return 1e-8;
    case INTERROBOT: return 1e-2;
    case OBSTACLE: return f.jac_delta;  // obstacle.rs:98-102
    case TRACKING: return 1e-8;
  }
  return 0;
}

// Factor::first_order_jacobian (factor/mod.rs:102-128).
Mat first_order_jacobian(Factor &f, Vec x) {
  Vec h0 = measure(f, x);
  Mat J(int(h0.size()), int(x.size()));
  double delta = jacobian_delta(f);
  for (size_t i = 0; i < x.size(); ++i) {
    x[i] += delta;
    Vec h1 = measure(f, x);
    for (size_t r = 0; r < h0.size(); ++r) J(int(r), int(i)) = (h1[r] - h0[r]) / delta;
    x[i] -= delta;
  }
  return J;
}

// Factor::jacobian dispatch (factor/mod.rs:541-552).
Mat jacobian(Factor &f, const Vec &x) {
  switch (f.kind) {
    case DYNAMIC: return f.cached_jacobian;
    case INTERROBOT: {  // interrobot.rs:121-161
      Mat J(int(f.z.size()), DOFS * 2);
      Vec d(2);
      for (int i = 0; i < 2; ++i) d[i] = (x[i] - x[DOFS + i]);
      for (int i = 0; i < 2; ++i) d[i] += f.tiny_offset;
      double radius = euclidean_norm(d);
      if (radius <= f.safety_distance) {
        double a = -1.0 / f.safety_distance / radius;
        double b = 1.0 / f.safety_distance / radius;
        for (int i = 0; i < 2; ++i) {
          J(0, i) = a * d[i];
          J(0, DOFS + i) = b * d[i];
        }
      }
      return J;
    }
    case OBSTACLE: return first_order_jacobian(f, x);  // obstacle.rs:129-137
    case TRACKING: {  // tracking.rs:171-194
      Mat J(1, DOFS);
      double h0 = f.last_value;
      for (int i = 0; i < 2; ++i) {
        double xd = x[i] - double(f.last_pos[i]);
        J(0, i) = (1.0 / h0) * xd;
      }
      return J;
    }
  }
  return {};
}

// Factor::skip dispatch (factor/mod.rs:565-572).
bool skip(Factor &f) {
  switch (f.kind) {
    case DYNAMIC:
    case OBSTACLE: return false;
    case INTERROBOT: {  // interrobot.rs:213-226
      double sq = 0.0;
      for (int i = 0; i < 2; ++i) {
        double d = f.lin[i] - f.lin[DOFS + i];
        sq = sq + d * d;  // mapv(powi 2).sum()
      }
      return sq >= f.safety_distance * f.safety_distance;
    }
    case TRACKING: {  // tracking.rs:362-381
      if (f.timeout) {
        if (*f.timeout == 0) f.timeout.reset();
        else {
          f.timeout = *f.timeout - 1;
          return true;
        }
      }
      // the reference computes `len - 1` on usize; an empty path never occurs
      // once a route exists (robot.rs:1316-1322 builds it from >= 2 waypoints)
      if (f.path.size() < 2 || f.trk_record >= f.path.size() - 1) return true;
      return false;
    }
  }
  return false;
}

// extract_submatrices_from_precision_matrix / the inline slicing of
// marginalise_factor_distance (factor/marginalise_factor_distance.rs:22-52,74-102):
// "a" = rows/cols [marg_idx, marg_idx+4); "b" = everything after it when
// marg_idx == 0, everything before it otherwise.
struct Blocks {
  Mat aa, ab, ba, bb;
};
Blocks extract_blocks(const Mat &lam, int marg_idx) {
  int n = lam.r, nb = n - DOFS;
  auto bi = [&](int k) { return marg_idx == 0 ? DOFS + k : k; };
  Blocks b{Mat(DOFS, DOFS), Mat(DOFS, nb), Mat(nb, DOFS), Mat(nb, nb)};
  for (int i = 0; i < DOFS; ++i)
    for (int j = 0; j < DOFS; ++j) b.aa(i, j) = lam(marg_idx + i, marg_idx + j);
  for (int i = 0; i < DOFS; ++i)
    for (int j = 0; j < nb; ++j) {
      b.ab(i, j) = lam(marg_idx + i, bi(j));
      b.ba(j, i) = lam(bi(j), marg_idx + i);
    }
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j) b.bb(i, j) = lam(bi(i), bi(j));
  return b;
}

// marginalise_factor_distance (factor/marginalise_factor_distance.rs:55-127).
Message marginalise_factor_distance(const Vec &eta, const Mat &lam, int marg_idx) {
  int n = int(eta.size());
  if (n == DOFS) return Message::make(eta, lam, Vec(n, 0.0));
  int nb = n - DOFS;
  Blocks b = extract_blocks(lam, marg_idx);
  auto lam_bb_inv = inv(b.bb);
  if (!lam_bb_inv) return Message::empty();
  Vec eta_a(DOFS), eta_b(nb);
  for (int i = 0; i < DOFS; ++i) eta_a[i] = eta[marg_idx + i];
  for (int i = 0; i < nb; ++i) eta_b[i] = eta[marg_idx == 0 ? DOFS + i : i];
  Mat t = matmul(b.ab, *lam_bb_inv);
  Vec te = matvec(t, eta_b);
  Mat t2 = matmul(b.ab, *lam_bb_inv);  // the reference computes it twice (:114-115)
  Mat tl = matmul(t2, b.ba);
  Vec eta_o(DOFS);
  Mat lam_o(DOFS, DOFS);
  for (int i = 0; i < DOFS; ++i) eta_o[i] = eta_a[i] - te[i];
  for (int i = 0; i < DOFS * DOFS; ++i) lam_o.a[i] = b.aa.a[i] - tl.a[i];
  for (double v : lam_o.a)
    if (std::isinf(v)) return Message::empty();
  return Message::make(eta_o, lam_o, Vec(DOFS, 0.0));
}

// FactorNode::update (factor/mod.rs:334-454).
std::vector<std::pair<NodeId, Message>> factor_update(Factor &f) {
  int i = 0;
  for (auto &kv : f.inbox) {
    if (size_t(i + 1) * DOFS > f.lin.size()) break;  // cannot happen: |inbox| <= neighbours
    if (!kv.second.is_empty())
      for (int k = 0; k < DOFS; ++k) f.lin[i * DOFS + k] = kv.second.payload->mu[k];
    else
      for (int k = 0; k < DOFS; ++k) f.lin[i * DOFS + k] = 0.0;
    ++i;
  }
  std::vector<std::pair<NodeId, Message>> out;
  // messages_sent, per inbox key, in both the skip branch and the regular one (factor/mod.rs:353-367, 410-452)
  for (auto &kv : f.inbox) (kv.first.first == f.graph ? f.cnt.sent_int : f.cnt.sent_ext) += 1;
  if (skip(f)) {
    for (auto &kv : f.inbox) out.emplace_back(kv.first, Message::empty());
    return out;
  }
  Vec h = measure(f, f.lin);
  Mat J = jacobian(f, f.lin);
  Mat Jt = transpose(J);
  Mat JtL = matmul(Jt, f.meas_prec);
  Mat lam_p = matmul(JtL, J);
  Vec residual(f.z.size());
  for (size_t k = 0; k < f.z.size(); ++k) residual[k] = f.z[k] - h[k];
  Vec jx = matvec(J, f.lin);
  for (size_t k = 0; k < jx.size(); ++k) jx[k] = jx[k] + residual[k];
  Mat JtL2 = matmul(Jt, f.meas_prec);
  Vec eta_p = matvec(JtL2, jx);

  int marg = 0;
  for (auto &to : f.inbox) {
    Vec eta = eta_p;
    Mat lam = lam_p;
    int j = 0;
    for (auto &other : f.inbox) {
      if (other.first != to.first && !other.second.is_empty()) {
        const Payload &p = *other.second.payload;
        for (int a = 0; a < DOFS; ++a) eta[j * DOFS + a] += p.eta[a];
        for (int a = 0; a < DOFS; ++a)
          for (int b = 0; b < DOFS; ++b) lam(j * DOFS + a, j * DOFS + b) += p.lam(a, b);
      }
      ++j;
    }
    out.emplace_back(to.first, marginalise_factor_distance(eta, lam, marg));
    marg += DOFS;
  }
  return out;
}

// variable.rs:15-54,86-106
struct Variable {
  MsgCount cnt;
  int graph, index;
  Vec prior_eta;
  Mat prior_lam;
  Vec eta, mu;
  Mat lam, cov;
  bool valid;
  Inbox inbox;  // keyed by FactorId
};

// VariableNode::new (variable.rs:140-166).
Variable make_variable(int graph, int index, const Vec &prior_mean, Mat prior_lam) {
  bool finite = true;
  for (double v : prior_lam.a) finite = finite && std::isfinite(v);
  if (!finite) std::fill(prior_lam.a.begin(), prior_lam.a.end(), 0.0);
  Variable v;
  v.graph = graph;
  v.index = index;
  v.prior_eta = matvec(prior_lam, prior_mean);
  auto sigma = inv(prior_lam);
  v.cov = sigma ? *sigma : Mat(DOFS, DOFS);
  v.prior_lam = prior_lam;
  v.eta = v.prior_eta;
  v.lam = prior_lam;
  v.mu = prior_mean;
  v.valid = true;
  for (double c : v.cov.a) v.valid = v.valid && std::isfinite(c);
  return v;
}
// VariableNode::prepare_message (variable.rs:234-240).
Message prepare_message(const Variable &v) { return Message::make(v.eta, v.lam, v.mu); }

// VariableNode::change_prior (variable.rs:203-230).
std::vector<std::pair<NodeId, Message>> change_prior(Variable &v, const Vec &mean) {
  v.prior_eta = matvec(v.prior_lam, mean);
  v.mu = mean;
  std::vector<std::pair<NodeId, Message>> out;
  for (auto &kv : v.inbox) out.emplace_back(kv.first, prepare_message(v));
  for (auto &kv : v.inbox) kv.second = Message::empty();
  return out;
}

// VariableNode::update_belief_and_create_factor_responses (variable.rs:251-342).
std::vector<std::pair<NodeId, Message>> variable_update(Variable &v) {
  v.eta = v.prior_eta;
  v.lam = v.prior_lam;
  for (auto &kv : v.inbox) {
    if (kv.second.is_empty()) continue;
    const Payload &p = *kv.second.payload;
    for (int a = 0; a < DOFS; ++a) v.eta[a] = v.eta[a] + p.eta[a];
    for (int a = 0; a < DOFS * DOFS; ++a) v.lam.a[a] = v.lam.a[a] + p.lam.a[a];
  }
  bool precision_not_zero = false;
  for (double x : v.lam.a) precision_not_zero = precision_not_zero || (x - 1e-6 > 0.0);
  if (precision_not_zero) {
    if (auto cov = inv(v.lam)) {
      v.cov = *cov;
      v.valid = true;
      for (double c : v.cov.a) v.valid = v.valid && std::isfinite(c);
      if (v.valid) v.mu = matvec(v.cov, v.eta);
    }
  }
  std::vector<std::pair<NodeId, Message>> out;
  // one response per inbox key, counted whether or not the caller delivers it (variable.rs:299-332);
  // change_prior's messages are NOT counted as sent (variable.rs:208-229 drops its local counter)
  for (auto &kv : v.inbox) (kv.first.first == v.graph ? v.cnt.sent_int : v.cnt.sent_ext) += 1;
  for (auto &kv : v.inbox) {
    if (kv.second.is_empty()) {
      out.emplace_back(kv.first, prepare_message(v));
    } else {
      const Payload &p = *kv.second.payload;
      Vec e(DOFS), m(DOFS);
      Mat l(DOFS, DOFS);
      for (int a = 0; a < DOFS; ++a) e[a] = v.eta[a] - p.eta[a];
      for (int a = 0; a < DOFS * DOFS; ++a) l.a[a] = v.lam.a[a] - p.lam.a[a];
      for (int a = 0; a < DOFS; ++a) m[a] = v.mu[a] - p.mu[a];
      out.emplace_back(kv.first, Message::make(e, l, m));
    }
  }
  return out;
}

// factorgraph.rs:74-120 — one graph per robot.  Node indices: variables 0..V-1,
// then factors in creation order (petgraph StableGraph; free-slot reuse after
// deletion only permutes indices of a robot's OWN InterRobot factors, whose
// entries in own-variable inboxes are permanently Empty, so it has no effect).
struct Graph {
  int id;
  std::vector<Variable> vars;
  std::map<int, Factor> factors;  // node index -> factor (factor_indices order)
  int next_index = 0;
  uint64_t iter_factor = 0, iter_variable = 0;
};

// FactorNode::receive_message_from (factor/mod.rs:307-318).
void factor_receive(Factor &f, NodeId from, const Message &m) {
  if (!f.enabled) return;
  f.inbox[from] = m;
  (from.first == f.graph ? f.cnt.recv_int : f.cnt.recv_ext) += 1;
}
// VariableNode::receive_message_from (variable.rs:176-190).
void variable_receive(Variable &v, NodeId from, const Message &m) {
  v.inbox[from] = m;
  (from.first == v.graph ? v.cnt.recv_int : v.cnt.recv_ext) += 1;
}

struct Robot {
  Graph g;
  float radius, t0;
  float pos[2];  // Transform.translation.x / .z
  bool antenna = true, idle = false, finished = false;
  bool gone = false;  // the entity has been despawned (RobotDespawned, robot.rs:2171-2172): no query yields it any more
  std::set<int> within, connected;  // RobotConnections
  std::vector<std::pair<float, float>> waypoints;
  int next_wp = 1;
};

// ---- parry2d 0.13.7 (un-vendored git dependency, Cargo.lock:5372-5374), restated for the one call the path makes:
// query::intersection_test(&collider.isometry, collider.shape, &robot_pos, ball) (planner/collisions.rs:402-408).
// f32 throughout; this file is built with -ffp-contract=off like rustc (no fused multiply-add).
struct V2 {
  float x, y;
};
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, float t) { return {a.x * t, a.y * t}; }
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline float perp(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }  // nalgebra Vector2::perp
inline float norm_squared(V2 a) { return dot(a, a); }
struct PointProjection {
  bool is_inside;
  V2 point;
};
struct EnvCollider {
  int kind;  // 0 Ball, 1 Cuboid, 2 Triangle, 3 ConvexPolygon
  V2 translation;
  float re, im;  // UnitComplex::new(angle)
  float radius;
  V2 half_extents;
  std::vector<V2> points;
};
// Aabb::project_local_point (bounding_volume/aabb_utils / query/point/point_aabb.rs), solid = true
inline PointProjection project_cuboid(V2 he, V2 pt) {
  const V2 mins{-he.x, -he.y}, maxs = he;
  const V2 mins_pt = mins - pt, pt_maxs = pt - maxs;
  const V2 shift{std::max(mins_pt.x, 0.0f) - std::max(pt_maxs.x, 0.0f), std::max(mins_pt.y, 0.0f) - std::max(pt_maxs.y, 0.0f)};
  const bool inside = shift.x == 0.0f && shift.y == 0.0f;
  if (!inside) return {false, pt + shift};
  return {true, pt};
}
// Triangle::project_local_point_and_get_location (query/point/point_triangle.rs), 2-D
inline PointProjection project_triangle(V2 a, V2 b, V2 c, V2 pt) {
  const V2 ab = b - a, ac = c - a, ap = pt - a;
  const float ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
  if (ab_ap <= 0.0f && ac_ap <= 0.0f) return {false, a};
  const V2 bp = pt - b;
  const float ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
  if (ab_bp >= 0.0f && ac_bp <= ab_bp) return {false, b};
  const V2 cp = pt - c;
  const float ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
  if (ac_cp >= 0.0f && ab_cp <= ac_cp) return {false, c};
  // stable_check_edges_voronoi, DIM == 2
  const float n = perp(ab, ac);
  const float vc = n * perp(ab, ap);
  if (vc < 0.0f && ab_ap >= 0.0f && ab_bp <= 0.0f) {
    const float v = ab_ap / norm_squared(ab);
    return {false, a + ab * v};
  }
  const float vb = -n * perp(ac, cp);
  if (vb < 0.0f && ac_ap >= 0.0f && ac_cp <= 0.0f) {
    const float w = ac_ap / norm_squared(ac);
    return {false, a + ac * w};
  }
  const V2 bc = c - b;
  const float va = n * perp(bc, bp);
  if (va < 0.0f && ac_bp - ab_bp >= 0.0f && ab_cp - ac_cp >= 0.0f) {
    const float w = dot(bc, bp) / norm_squared(bc);
    return {false, b + bc * w};
  }
  return {true, pt};  // on the face: inside in two dimensions
}
// ConvexPolygon: parry projects through GJK on the support map (query/point/point_support_map.rs); restated as the
// predicate GJK converges to — inside every edge of the counter-clockwise hull, else the nearest point of an edge.
inline PointProjection project_convex_polygon(const std::vector<V2> &pts, V2 pt) {
  bool inside = pts.size() >= 3;
  float best = 3.4e38f;
  V2 bestp = pt;
  for (size_t k = 0; k < pts.size(); ++k) {
    const V2 a = pts[k], b = pts[k + 1 == pts.size() ? 0 : k + 1];
    const V2 e = b - a, w = pt - a;
    inside = inside && perp(e, w) >= 0.0f;
    const float ee = dot(e, e);
    float t = ee > 0.0f ? dot(w, e) / ee : 0.0f;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    const V2 q = a + e * t;
    const float d2 = norm_squared(pt - q);
    if (d2 < best) {
      best = d2;
      bestp = q;
    }
  }
  if (inside) return {true, pt};
  return {false, bestp};
}
// query::intersection_test -> DefaultQueryDispatcher::intersection_test (query/default_query_dispatcher.rs):
// Ball / Ball -> intersection_test_ball_ball, otherwise shape2 is the Ball -> intersection_test_point_query_ball.
inline bool intersection_test(const EnvCollider &c, V2 robot, float ball_radius) {
  // pos12 = pos1.inv_mul(&pos2): translation = rotation1.inverse() * (t2 - t1)  (nalgebra Isometry::inv_mul)
  const V2 tr = robot - c.translation;
  const float r = c.re, i = -c.im;  // UnitComplex::inverse = conjugate
  const V2 local{r * tr.x - i * tr.y, i * tr.x + r * tr.y};
  if (c.kind == 0) {
    const float sum_radius = c.radius + ball_radius;
    return norm_squared(local) <= sum_radius * sum_radius;
  }
  PointProjection proj;
  if (c.kind == 1) proj = project_cuboid(c.half_extents, local);
  else if (c.kind == 2) proj = project_triangle(c.points[0], c.points[1], c.points[2], local);
  else proj = project_convex_polygon(c.points, local);
  return proj.is_inside || norm_squared(local - proj.point) <= ball_radius * ball_radius;
}

// PositionTracker + VelocityTracker of one robot (planner/tracking.rs:36-200); both Timers are created together with
// the same duration and tick on the same frames, so one stopwatch serves both.
struct Tracker {
  uint64_t elapsed_ns = 0;
  std::vector<std::array<float, 2>> positions;  // HeapRb contents, oldest first
  struct Vel {
    float v[2];
    double timestamp, measured_over;
  };
  std::vector<Vel> velocities;
  uint64_t npos = 0, nvel = 0;
  bool has_prev = false;
  float prev[2] = {0, 0};
  double prev_t = 0;
};

struct World {
  // RobotRobotCollisions (planner/collisions.rs:146-200): state per unordered pair, total Hit count
  std::map<std::pair<int, int>, bool> coll_state;
  std::vector<uint32_t> coll_hits;
  int64_t collisions = 0;
  Cfg cfg;
  Sdf sdf;
  std::vector<uint32_t> timesteps;
  std::vector<Robot> robots;
  uint64_t robot_number = 1;  // RobotNumberGenerator (robot.rs:121-144)
  int threads = 1;
  // Colliders resource + RobotEnvironmentCollisions (planner/collisions.rs:202-330)
  std::vector<EnvCollider> colliders;
  std::map<std::pair<int, int>, bool> env_state;  // (robot, collider) -> CollisionState::Colliding
  std::vector<uint32_t> env_hits;
  int64_t env_collisions = 0;
  // CollisionHistory::aabbs (collisions.rs:463-470): one Aabb per Hit, pushed by record_aabb_when_two_robots_collide /
  // record_aabb_when_robot_collides_with_environment (:700-716) from the events the update systems send
  struct CollEvent {
    int a, b;       // robot-robot: the map key (r, c), r < c; robot-environment: (robot, collider index)
    float aabb[4];  // mins x, mins z, maxs x, maxs z of the intersection
  };
  std::vector<CollEvent> robot_events, env_events;
  // PositionTracker / VelocityTracker per robot (planner/tracking.rs:36-200)
  int track_capacity = 0;
  uint64_t track_duration_ns = 0;
  std::vector<Tracker> trackers;
  size_t track_seen = 0;
};

// FactorGraph::add_internal_edge (factorgraph.rs:304-330).
void add_internal_edge(Graph &g, int var, int fidx) {
  Variable &v = g.vars[var];
  variable_receive(v, {g.id, fidx}, Message::empty());
  Message vm = prepare_message(v);
  Factor &f = g.factors.at(fidx);
  if (f.kind == TRACKING) factor_receive(f, {g.id, var}, vm);
  else factor_receive(f, {g.id, var}, Message::empty());
}

// RobotBundle::new (planner/robot.rs:1134-1355).
void add_robot(World &w, float radius, const double *means, const float *pos,
               const float *wp_xy, int nwp) {
  const Cfg &c = w.cfg;
  int V = c.num_variables;
  w.robots.emplace_back();
  Robot &r = w.robots.back();
  int id = int(w.robots.size()) - 1;
  r.g.id = id;
  r.radius = radius;
  r.pos[0] = pos[0];
  r.pos[1] = pos[1];
  for (int k = 0; k < nwp; ++k) r.waypoints.emplace_back(wp_xy[2 * k], wp_xy[2 * k + 1]);
  for (int i = 0; i < V; ++i) {  // :1179-1223
    double sigma = (i == 0 || i == V - 1) ? 1e30 : std::numeric_limits<double>::infinity();
    Mat P(DOFS, DOFS);
    for (int a = 0; a < DOFS; ++a) P(a, a) = sigma;  // from_diag_elem
    Vec m(means + size_t(i) * DOFS, means + size_t(i + 1) * DOFS);
    r.g.vars.push_back(make_variable(id, i, m, P));
  }
  r.g.next_index = V;
  r.t0 = radius / 2.0f / c.target_speed;  // :1225 (f32)
  for (int i = 0; i < V - 1; ++i) {        // :1228-1255
    float delta_t = r.t0 * float(w.timesteps[i + 1] - w.timesteps[i]);
    Factor f;
    f.kind = DYNAMIC;
    f.graph = id;
    f.index = r.g.next_index++;
    f.enabled = c.enable_dynamic;
    init_state(f, Vec(DOFS, 0.0), double(c.sigma_factor_dynamics), 2);
    init_dynamic(f, double(c.sigma_factor_dynamics), double(delta_t));
    int fi = f.index;
    f.own_var = i;
    r.g.factors.emplace(fi, std::move(f));
    add_internal_edge(r.g, i + 1, fi);
    add_internal_edge(r.g, i, fi);
  }
  for (int i = 1; i < V - 1; ++i) {  // :1269-1285
    Factor f;
    f.kind = OBSTACLE;
    f.graph = id;
    f.index = r.g.next_index++;
    f.enabled = c.enable_obstacle;
    init_state(f, Vec{0.0}, double(c.sigma_factor_obstacle), 1);
    f.sdf = &w.sdf;
    f.world_w = c.world_width;
    f.world_h = c.world_height;
    f.jac_delta = (c.world_width / double(uint32_t(w.sdf.w)) +
                   c.world_height / double(uint32_t(w.sdf.h))) / 2.0;  // obstacle.rs:98-102
    int fi = f.index;
    f.own_var = i;
    r.g.factors.emplace(fi, std::move(f));
    add_internal_edge(r.g, i, fi);
  }
  for (int i = 1; i < V - 1; ++i) {  // :1305-1334
    Factor f;
    f.kind = TRACKING;
    f.graph = id;
    f.index = r.g.next_index++;
    f.enabled = c.enable_tracking;
    init_state(f, Vec{0.0}, double(c.sigma_factor_tracking), 1);
    f.lin = Vec{means[size_t(i) * DOFS], means[size_t(i) * DOFS + 1], 0.0, 0.0};
    f.path = r.waypoints;
    f.last_pos[0] = float(f.lin[0]);
    f.last_pos[1] = float(f.lin[1]);
    f.last_value = 0.0;
    f.switch_padding = c.tracking_switch_padding;
    f.attraction_distance = c.tracking_attraction_distance;
    int fi = f.index;
    f.own_var = i;
    r.g.factors.emplace(fi, std::move(f));
    add_internal_edge(r.g, i, fi);
  }
}

// update_robot_neighbours (planner/robot.rs:1362-1384): glam Vec3::distance in f32.
void update_neighbours(World &w) {
  int n = int(w.robots.size());
  float R = w.cfg.comms_radius;
  parallel_for(n, w.threads, 64, [&](int a) {
    Robot &ra = w.robots[a];
    ra.within.clear();
    if (ra.gone) return;  // not in the query
    for (int b = 0; b < n; ++b) {
      if (b == a || w.robots[b].gone) continue;
      float dx = ra.pos[0] - w.robots[b].pos[0];
      float dy = 0.0f;
      float dz = ra.pos[1] - w.robots[b].pos[1];
      float d = std::sqrt((dx * dx + dy * dy) + dz * dz);
      if (R < d) continue;
      ra.within.insert(b);
    }
  });
}

// FactorGraph::delete_interrobot_factors_connected_to (factorgraph.rs:380-436).
void delete_ir_connected_to(Graph &g, int other) {
  std::vector<int> removed;
  for (auto &v : g.vars)
    for (auto it = v.inbox.begin(); it != v.inbox.end();)
      it = (it->first.first == other) ? v.inbox.erase(it) : std::next(it);
  for (auto it = g.factors.begin(); it != g.factors.end();) {
    if (it->second.kind == INTERROBOT && it->second.ext_robot == other) {
      removed.push_back(it->first);
      it = g.factors.erase(it);
    } else ++it;
  }
  for (auto &v : g.vars)
    for (int fi : removed) v.inbox.erase({g.id, fi});
}

// delete_interrobot_factors (planner/robot.rs:1386-1439).  The reference funnels
// the (robot, lost neighbour) pairs through a HashMap<RobotId, RobotId>, which
// drops pairs when one robot loses several neighbours in a tick.
// cfg.strict_reference_quirks = 1 reproduces that; 0 deletes every lost pair (the
// symmetric pair covers it in the reference whenever the quirk does not trigger).
// B.2 (stale interrobot_factor_indices, factorgraph.rs:409-415) needs nothing here:
// a stale entry can only make an InterRobot factor update twice in one half, which
// is idempotent (same inbox, same messages); factor indices are never reused here.
void delete_interrobot_factors(World &w) {
  std::vector<std::pair<int, int>> pairs;
  for (auto &r : w.robots) {
    if (r.gone) {  // despawned with its RobotConnections; its graph is never looked at again
      r.connected.clear();
      continue;
    }
    std::vector<int> lost;
    for (int c : r.connected)
      if (!r.within.count(c)) lost.push_back(c);
    for (int c : lost) {
      pairs.emplace_back(r.g.id, c);
      r.connected.erase(c);
    }
  }
  if (w.cfg.strict_reference_quirks) {
    // `HashMap<RobotId, RobotId>::extend`: one entry per robot, the last inserted pair wins — the lost neighbours are
    // inserted in BTreeSet (ascending) order, so it is the largest id.  Which pairs are processed does not depend on
    // the map's iteration order, and deleting is idempotent.
    std::map<int, int> last;
    for (auto &p : pairs) last[p.first] = p.second;
    pairs.assign(last.begin(), last.end());
  }
  for (auto &p : pairs) {
    delete_ir_connected_to(w.robots[p.first].g, p.second);
    // query.get_mut(robot2) fails for a despawned robot (robot.rs:1428-1437: error!, nothing deleted)
    if (!w.robots[p.second].gone) delete_ir_connected_to(w.robots[p.second].g, p.first);
  }
}

// create_interrobot_factors (planner/robot.rs:1441-1586).
void create_interrobot_factors(World &w) {
  const Cfg &c = w.cfg;
  int V = c.num_variables;
  struct Ext { int robot, fidx, other, i; };
  std::vector<Ext> ext;
  for (auto &r : w.robots) {
    std::vector<int> fresh;
    for (int o : r.within)
      if (!r.connected.count(o)) fresh.push_back(o);  // BTreeSet::difference, ascending
    for (int other : fresh) {
      for (int i = 1; i < V; ++i) {
        Factor f;
        f.kind = INTERROBOT;
        f.graph = r.g.id;
        f.index = r.g.next_index++;
        f.enabled = c.enable_interrobot;
        init_state(f, Vec(DOFS, 0.0), double(c.sigma_factor_interrobot), 2);
        f.robot_radius = double(r.radius);
        f.safety_distance = double(c.safety_distance_multiplier) * f.robot_radius;
        f.ext_robot = other;
        f.ext_var = i;
        f.robot_number = w.robot_number++;
        f.tiny_offset = double(1e-6f) * double(f.robot_number);  // interrobot.rs:52,75
        int fi = f.index;
        f.own_var = i;
        r.g.factors.emplace(fi, std::move(f));
        add_internal_edge(r.g, i, fi);
        ext.push_back({r.g.id, fi, other, i});
      }
      r.connected.insert(other);
    }
  }
  struct Tmp { int robot, fidx; Message m; NodeId from; };
  std::vector<Tmp> tmp;
  for (auto &e : ext) {  // add_external_edge (factorgraph.rs:340-353)
    Graph &og = w.robots[e.other].g;
    variable_receive(og.vars[e.i], {e.robot, e.fidx}, Message::empty());
    tmp.push_back({e.robot, e.fidx, prepare_message(og.vars[e.i]), {e.other, e.i}});
  }
  for (auto &t : tmp) {
    auto it = w.robots[t.robot].g.factors.find(t.fidx);
    if (it != w.robots[t.robot].g.factors.end()) factor_receive(it->second, t.from, t.m);
  }
}

// FactorGraph::internal_factor_iteration (factorgraph.rs:688-714).
void internal_factor_iteration(Graph &g) {
  for (auto &kv : g.factors) {
    Factor &f = kv.second;
    if (!f.enabled) continue;
    if (f.kind == INTERROBOT) continue;
    if (f.kind == TRACKING && g.iter_factor < 10) continue;
    auto msgs = factor_update(f);
    for (auto &m : msgs) variable_receive(g.vars[m.first.second], {g.id, f.index}, m.second);
  }
  g.iter_factor += 1;
}
// FactorGraph::internal_variable_iteration (factorgraph.rs:762-790).
void internal_variable_iteration(Graph &g) {
  for (auto &v : g.vars) {
    auto msgs = variable_update(v);
    for (auto &m : msgs) {
      if (m.first.first != g.id) continue;
      auto it = g.factors.find(m.first.second);
      if (it == g.factors.end()) continue;
      if (!it->second.enabled) continue;
      factor_receive(it->second, {g.id, v.index}, m.second);
    }
  }
  g.iter_variable += 1;
}
struct Routed { NodeId from, to; Message m; };
// FactorGraph::external_factor_iteration (factorgraph.rs:719-760).
void external_factor_iteration(Graph &g, std::vector<Routed> &out) {
  for (auto &kv : g.factors) {
    Factor &f = kv.second;
    if (f.kind != INTERROBOT || !f.enabled) continue;
    auto msgs = factor_update(f);
    for (auto &m : msgs)
      if (m.first.first != g.id) out.push_back({{g.id, f.index}, m.first, m.second});
  }
  g.iter_factor += 1;
}
// FactorGraph::external_variable_iteration (factorgraph.rs:794-826).
void external_variable_iteration(Graph &g, std::vector<Routed> &out) {
  for (auto &v : g.vars) {
    auto msgs = variable_update(v);
    for (auto &m : msgs)
      if (m.first.first != g.id) out.push_back({{g.id, v.index}, m.first, m.second});
  }
  g.iter_variable += 1;
}

void world_internal(World &w, bool factors, bool variables) {
  int n = int(w.robots.size());
  // query.par_iter_mut() (robot.rs:1789-1800): threads over robots
  parallel_for(n, w.threads, 16, [&](int i) {
    Robot &r = w.robots[i];
    if (r.idle || r.gone) return;
    if (factors) internal_factor_iteration(r.g);
    if (variables) internal_variable_iteration(r.g);
  });
}
void world_external_factor(World &w) {  // robot.rs:1803-1831 (single thread)
  std::vector<Routed> msgs;
  for (auto &r : w.robots) {
    if (!r.antenna || r.idle || r.gone) continue;
    external_factor_iteration(r.g, msgs);
  }
  for (auto &m : msgs) {
    Robot &t = w.robots[m.to.first];
    if (!t.antenna || t.idle || t.gone) continue;  // gone: query.get_mut fails (robot.rs:1815-1819, 1844-1848)
    variable_receive(t.g.vars[m.to.second], m.from, m.m);
  }
}
void world_external_variable(World &w) {  // robot.rs:1833-1858 (single thread)
  std::vector<Routed> msgs;
  for (auto &r : w.robots) {
    if (!r.antenna || r.idle || r.gone) continue;
    external_variable_iteration(r.g, msgs);
  }
  for (auto &m : msgs) {
    Robot &t = w.robots[m.to.first];
    if (!t.antenna || t.idle || t.gone) continue;  // gone: query.get_mut fails (robot.rs:1815-1819, 1844-1848)
    auto it = t.g.factors.find(m.to.second);
    if (it != t.g.factors.end()) factor_receive(it->second, m.from, m.m);
  }
}

// gbp_schedule crate ---------------------------------------------------------
// interleave_evenly.rs:40-110
void ie_recurse(uint8_t *s, int len, int n) {
  int max = len, half = max / 2;
  auto fill_cycle = [&](int times) {
    for (int i = 0; i < len; ++i) s[i] = (i % times) == 0;
  };
  if (n == max) { std::fill(s, s + len, 1); return; }
  if (n == 0) { std::fill(s, s + len, 0); return; }
  bool on = n % 2 == 1, om = max % 2 == 1;
  if (on && om) {
    if (max % n == 0) fill_cycle(max / n);
    else {
      int k = n / 2;
      ie_recurse(s, half, k);
      s[half] = 1;
      ie_recurse(s + half + 1, len - half - 1, k);
      std::reverse(s + half + 1, s + len);
    }
  } else if (!on && om) {
    int k = n / 2;
    ie_recurse(s, half, k);
    std::reverse(s, s + half);
    s[half] = 0;
    ie_recurse(s + half + 1, len - half - 1, k);
  } else if (!on && !om) {
    if (max % n == 0) fill_cycle(max / n);
    else {
      int k = n / 2;
      ie_recurse(s, half, k);
      ie_recurse(s + half, len - half, k);
    }
  } else {
    int k = n / 2;
    ie_recurse(s, half, k + 1);
    std::reverse(s, s + half);
    ie_recurse(s + half, len - half, k);
  }
}
void expand(int kind, int n, int max, uint8_t *out) {
  switch (kind) {
    case 0: {  // centered.rs:12-49
      for (int idx = 0; idx < max; ++idx) {
        if (n == 0 && max == 1) { out[idx] = 0; continue; }
        int mid = max / 2, hn = n / 2;
        int start = mid >= hn ? mid - hn : 0;
        int end = (start + n <= max) ? start + n - 1 : max - 1;
        out[idx] = idx >= start && idx <= end;
      }
      break;
    }
    case 1: ie_recurse(out, max, n); break;
    case 2:  // soon_as_possible.rs:26-49
      for (int i = 0; i < max; ++i) out[i] = i < n;
      break;
    case 3:  // late_as_possible.rs:29-50
      for (int i = 0; i < max; ++i) out[i] = (n == max) ? 1 : (n == 0 ? 0 : i >= max - n);
      break;
    case 4: {  // half_beginning_half_end.rs:19-45
      int hn = n / 2, rem = n % 2, sm = hn, em = max - hn - rem;
      for (int i = 0; i < max; ++i) out[i] = (i < sm || i >= em);
      break;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// C entry points for ctypes (tests / bench only)
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// crates/env_to_png/src/lib.rs — Environment -> SDF image: rasterise the tile
// grid (env_to_image :166-207, is_tile_obstacle :340-479), then `image::imageops::blur`.
// and the placeable obstacles (is_placeable_obstacle :283-336, shapes of gbp_environment/src/lib.rs:118-435;
// glam 0.25 `Quat::from_rotation_z(..).mul_vec3` is third party and restated from its formula).
// Third party, absent from /root/reference: image 0.25.1 `imageops::blur` (Cargo.lock) — restated
// from its published source: sigma <= 0 -> 1; separable Gaussian with support 2*sigma, vertical pass
// into an f32 image, horizontal pass clamped to [0, 255] and rounded half away from zero; per output
// row/column the window is [floor(c - support), ceil(c + support)) around c = index + 0.5, clamped
// to the image, weights gaussian(i - index, sigma) in f32 normalised by their in-order sum, taps
// accumulated in order as t += v * w.  PARITY UNPINNED: no reference test covers the blur.
// ---------------------------------------------------------------------------
namespace envpng {
// image_to_tile_units (:213-223)
inline void image_to_tile_units(uint32_t px, uint32_t py, uint32_t res, float tile_size, float &ox, float &oy) {
  const float x = float(px) + 0.5f, y = float(py) + 0.5f;
  ox = x / float(res) * tile_size;
  oy = y / float(res) * tile_size;
}
// offset_modulus (:241-243)
inline float offset_modulus(float value, float modulus) {
  return -(std::ceil(value / modulus) * modulus - value) / modulus + 1.0f;
}
// image_to_tile_coords (:249-258)
inline void image_to_tile_coords(uint32_t px, uint32_t py, uint32_t res, uint64_t &tx, uint64_t &ty) {
  tx = uint64_t(std::floor(float(px) / float(res)));
  ty = uint64_t(std::floor(float(py) / float(res)));
}
// is_tile_obstacle (:340-479); tile = Unicode code point
inline bool is_tile_obstacle(uint32_t tile, float path_width_in, float px, float py, float expansion) {
  const float path_width = path_width_in - expansion;
  const float almost_full = 1.0f - path_width;
  const float ow = almost_full / 2.0f;       // obstacle_width
  const float opw = ow + path_width;         // obstacle_and_path_width
  const float half_lo = 0.5f - expansion / 2.0f, half_hi = 0.5f + expansion / 2.0f;
  switch (tile) {
    case 0x2500: return py < ow || py > opw;                                        // ─
    case 0x2502: return px < ow || px > opw;                                        // │
    case 0x2574: return py < ow || py > opw || px > half_lo;                        // ╴
    case 0x2576: return py < ow || py > opw || px < half_hi;                        // ╶
    case 0x2577: return px < ow || px > opw || py < half_hi;                        // ╷
    case 0x2575: return px < ow || px > opw || py > half_lo;                        // ╵
    case 0x250C: return px < ow || py < ow || (px > opw && py > opw);               // ┌
    case 0x2510: return px > opw || py < ow || (px < ow && py > opw);               // ┐
    case 0x2514: return px < ow || py > opw || (px > opw && py < ow);               // └
    case 0x2518: return px > opw || py > opw || (px < ow && py < ow);               // ┘
    case 0x252C: return py < ow || (py > opw && (px < ow || px > opw));             // ┬
    case 0x2534: return py > opw || (py < ow && (px < ow || px > opw));             // ┴
    case 0x251C: return px < ow || (px > opw && (py < ow || py > opw));             // ├
    case 0x2524: return px > opw || (px < ow && (py < ow || py > opw));             // ┤
    case 0x253C: return (px < ow || px > opw) && (py < ow || py > opw);             // ┼
    case 0x20: return true;                                                         // ' '
    default: return false;
  }
}
// ---- placeable obstacles (crates/gbp_environment/src/lib.rs:118-435, env_to_png/src/lib.rs:283-336),
// restated literally: every `inside()` recomputes its vertices, as the reference does per pixel.
struct Obstacle {  // same layout as gbp_obstacle_t of the product header (plain data from the caller)
  int32_t kind, tile_row, tile_col;
  double rotation, tx, ty, radius, angle_a, angle_b;
  int32_t sides, n_points;
  int64_t point_offset;
  double width, height;
};
struct V2 {
  float x, y;
};
constexpr float kPiF = 3.14159265358979323846f, kHalfPiF = 1.57079632679489661923f;
constexpr double kPiD = 3.14159265358979323846, kQuarterPiD = 0.78539816339744830962;
inline V2 from_angle(float a) { return V2{std::cos(a), std::sin(a)}; }  // glam Vec2::from_angle
inline float tri_sign(V2 p1, V2 p2, V2 p3) {                             // :229-231
  return (p1.x - p3.x) * (p2.y - p3.y) - (p2.x - p3.x) * (p1.y - p3.y);
}
inline bool circle_inside(double radius, V2 p) {  // :138-141
  const float sq = p.x * p.x + p.y * p.y;
  return sq <= float(radius * radius);
}
inline bool triangle_inside(double angle_a, double angle_b, double radius_d, V2 p) {  // :192-226
  const float a = float(angle_a), b = float(angle_b);
  const float c = kPiF - (a + b);
  const float radius = float(radius_d);
  const float ah = radius / std::sin(a), bh = radius / std::sin(b), ch = radius / std::sin(c);
  const V2 ad = from_angle(kPiF + a / 2.0f), bd = from_angle(-b / 2.0f), cd = from_angle(kPiF - b - c / 2.0f);
  const V2 A{ad.x * ah, ad.y * ah}, B{bd.x * bh, bd.y * bh}, C{cd.x * ch, cd.y * ch};
  const float d1 = tri_sign(p, A, B), d2 = tri_sign(p, B, C), d3 = tri_sign(p, C, A);
  const bool has_neg = d1 < 0.0f || d2 < 0.0f || d3 < 0.0f;
  const bool has_pos = d1 > 0.0f || d2 > 0.0f || d3 > 0.0f;
  return !(has_neg && has_pos);
}
inline void regular_point_at(int sides, double radius, int i, double &x, double &y) {  // :271-287
  const double angle = 2.0 * kPiD / double(sides) * double(i) + kQuarterPiD;
  x = std::cos(angle) * radius;
  y = std::sin(angle) * radius;
}
inline bool regular_polygon_inside(int sides, double radius, V2 p) {  // :298-313
  bool inside = false;
  const double x = double(p.x) * 2.0, y = double(p.y) * 2.0;
  int j = sides - 1;
  for (int i = 0; i < sides; ++i) {
    double xi, yi, xj, yj;
    regular_point_at(sides, radius, i, xi, yi);
    regular_point_at(sides, radius, j, xj, yj);
    if ((yi < y && yj >= y) || (yj < y && yi >= y)) {
      if (xi + (y - yi) / (yj - yi) * (xj - xi) < x) inside = !inside;
    }
    j = i;
  }
  return inside;
}
inline bool rectangle_inside(double width, double height, V2 p) {  // :347-359
  const double x = double(p.x), y = double(p.y);
  const double half_width = width / 4.0, half_height = height / 4.0;
  return x >= -half_height && x <= half_height && y >= -half_width && y <= half_width;
}
inline bool polygon_inside(const std::vector<double> &pts, V2 p) {  // is_point_in_polygon :412-428
  const double px = double(p.x), py = double(p.y);
  bool inside = false;
  const size_t n = pts.size() / 2;
  size_t j = n - 1;
  for (size_t i = 0; i < n; ++i) {
    const double ix = pts[2 * i], iy = pts[2 * i + 1], jx = pts[2 * j], jy = pts[2 * j + 1];
    if ((iy > py) != (jy > py) && px < (jx - ix) * (py - iy) / (jy - iy) + ix) inside = !inside;
    j = i;
  }
  return inside;
}
// glam 0.25 Quat::from_rotation_z(angle).mul_vec3((x, y, 0)).xy() (third party, restated):
// rhs * (w*w - b.b) + b * (rhs.b * 2) + (b x rhs) * (w * 2) with b = (0, 0, sin(angle/2)), w = cos(angle/2)
inline V2 rotate_z(float angle, V2 v) {
  const float s = std::sin(angle * 0.5f), w = std::cos(angle * 0.5f);
  const float b2 = (0.0f * 0.0f + 0.0f * 0.0f) + s * s;
  const float k = w * w - b2;
  const float d2 = ((v.x * 0.0f + v.y * 0.0f) + 0.0f * s) * 2.0f;
  const float cx = 0.0f * 0.0f - s * v.y, cy = s * v.x - 0.0f * 0.0f;
  return V2{(v.x * k + 0.0f * d2) + cx * (w * 2.0f), (v.y * k + 0.0f * d2) + cy * (w * 2.0f)};
}
// is_placeable_obstacle (env_to_png/src/lib.rs:283-336)
inline bool is_placeable_obstacle(const Obstacle *obs, int n_obs, const double *poly_pts, uint64_t tile_x,
                                  uint64_t tile_y, float px, float py, float expansion_f) {
  const double expansion = double(expansion_f);
  for (int k = 0; k < n_obs; ++k) {
    const Obstacle &o = obs[k];
    if (uint64_t(o.tile_col) != tile_x || uint64_t(o.tile_row) != tile_y) continue;
    const V2 translated{px - float(o.tx), py - float(o.ty)};
    float rotation_offset = kHalfPiF;
    if (o.kind == 2) rotation_offset = kHalfPiF + kHalfPiF + ((o.sides % 2 != 0) ? kPiF / float(o.sides) : 0.0f);
    else if (o.kind == 3) rotation_offset = 0.0f;
    const V2 r = rotate_z(float(o.rotation) + rotation_offset, translated);
    bool inside = false;
    switch (o.kind) {
      case 0: inside = circle_inside(o.radius + expansion, r); break;                               // :128-134
      case 1: inside = triangle_inside(o.angle_a, o.angle_b, o.radius + expansion, r); break;       // :171-190
      case 2: inside = regular_polygon_inside(o.sides, o.radius + expansion * 2.0, r); break;       // :250-267
      case 3: {                                                                                     // :374-401
        const double *p = poly_pts + 2 * o.point_offset;
        double ax = 0.0, ay = 0.0;
        for (int i = 0; i < o.n_points; ++i) {
          ax = ax + p[2 * i];
          ay = ay + p[2 * i + 1];
        }
        const double cx = ax / double(o.n_points), cy = ay / double(o.n_points);
        std::vector<double> e;
        for (int i = 0; i < o.n_points; ++i) {
          const double dx = p[2 * i] - cx, dy = p[2 * i + 1] - cy;
          e.push_back(p[2 * i] + dx * 4.0 * expansion);
          e.push_back(p[2 * i + 1] + dy * 4.0 * expansion);
        }
        inside = polygon_inside(e, r);
        break;
      }
      case 4: inside = rectangle_inside(o.width + expansion * 2.0, o.height + expansion * 2.0, r); break;  // :335-344
      default: break;
    }
    if (inside) return true;
  }
  return false;
}

// image 0.25.1 imageops::sample::gaussian
inline float gaussian(float x, float r) {
  return (1.0f / (std::sqrt(2.0f * 3.14159265358979323846f) * r)) * std::exp(-(x * x) / (2.0f * (r * r)));
}
inline int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
}  // namespace envpng

extern "C" {

int gbpo_schedule(int32_t kind, uint8_t internal, uint8_t external, uint8_t *oi, uint8_t *oe) {
  if (kind < 0 || kind > 4) return -2;
  int max = std::max(internal, external);
  expand(kind, internal, max, oi);
  expand(kind, external, max, oe);
  return max;
}

// utils::get_variable_timesteps (utils.rs:35-75), f32 arithmetic with mul_add.
int gbpo_variable_timesteps(uint32_t h, uint32_t m, uint32_t *out, int32_t cap) {
  uint32_t n = 1 + uint32_t(0.5f * (-1.0f + std::sqrt(1.0f + 8.0f * float(h) / float(m))));
  int cnt = 0;
  for (uint32_t i = 0; i < m * (n + 1); ++i) {
    uint32_t section = i / m;
    float f = std::fmaf(float(m) / 2.0f, float(section),
                        std::fmaf(float(section), -float(m), float(i))) *
              (float(section) + 1.0f);
    if (cnt >= cap) return -2;
    if (f >= float(h)) { out[cnt++] = h; break; }
    out[cnt++] = uint32_t(f);
  }
  return cnt;
}

// marginalise_factor_distance on raw arrays (n = 4 or 8), for the reference KATs.
int gbpo_marginalise(int n, const double *eta, const double *lam, int marg_idx, double *oeta,
                     double *olam, double *omu) {
  Vec e(eta, eta + n);
  Mat l(n, n);
  std::copy(lam, lam + n * n, l.a.begin());
  Message m = marginalise_factor_distance(e, l, marg_idx);
  if (m.is_empty()) return 0;
  std::copy(m.payload->eta.begin(), m.payload->eta.end(), oeta);
  std::copy(m.payload->lam.a.begin(), m.payload->lam.a.end(), olam);
  std::copy(m.payload->mu.begin(), m.payload->mu.end(), omu);
  return int(m.payload->eta.size());
}
// block extraction of an 8x8 (reference KATs marginalise_factor_distance.rs:182-210)
int gbpo_extract_blocks(const double *lam, int marg_idx, double *aa, double *ab, double *ba, double *bb) {
  Mat l(8, 8);
  std::copy(lam, lam + 64, l.a.begin());
  Blocks b = extract_blocks(l, marg_idx);
  std::copy(b.aa.a.begin(), b.aa.a.end(), aa);
  std::copy(b.ab.a.begin(), b.ab.a.end(), ab);
  std::copy(b.ba.a.begin(), b.ba.a.end(), ba);
  std::copy(b.bb.a.begin(), b.bb.a.end(), bb);
  return 0;
}
int gbpo_inv4(const double *m, double *out) {
  Mat a(4, 4);
  std::copy(m, m + 16, a.a.begin());
  auto r = inv(a);
  if (!r) return 0;
  std::copy(r->a.begin(), r->a.end(), out);
  return 1;
}


void gbpo_image_to_tile_units(uint32_t px, uint32_t py, uint32_t res, float tile_size, float *out2) {
  envpng::image_to_tile_units(px, py, res, tile_size, out2[0], out2[1]);
}
void gbpo_tile_units_to_percentage(float x, float y, float tile_size, float *out2) {  // :230-238
  out2[0] = envpng::offset_modulus(x, tile_size);
  out2[1] = envpng::offset_modulus(y, tile_size);
}
void gbpo_image_to_tile_coords(uint32_t px, uint32_t py, uint32_t res, uint64_t *out2) {
  envpng::image_to_tile_coords(px, py, res, out2[0], out2[1]);
}
int gbpo_is_tile_obstacle(uint32_t tile, float path_width, float px, float py, float expansion) {
  return envpng::is_tile_obstacle(tile, path_width, px, py, expansion) ? 1 : 0;
}
// env_to_sdf_image (:149-163): out_rgb is [nrows*res][ncols*res][3]
int gbpo_env_to_sdf_image(int nrows, int ncols, const uint32_t *tiles, float tile_size, float path_width,
                          uint32_t res, float expansion, float blur_percent, int n_obstacles, const void *obstacles,
                          const double *polygon_points, uint8_t *out_rgb) {
  using namespace envpng;
  const uint32_t W = uint32_t(ncols) * res, H = uint32_t(nrows) * res;
  std::vector<uint8_t> img(size_t(W) * H);
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      uint64_t tx, ty;
      image_to_tile_coords(x, y, res, tx, ty);
      float ux, uy;
      image_to_tile_units(x, y, res, tile_size, ux, uy);
      const float fx = offset_modulus(ux, tile_size), fy = offset_modulus(uy, tile_size);
      if (ty >= uint64_t(nrows) || tx >= uint64_t(ncols)) return -1;  // "Tile not found"
      const bool obstacle = is_tile_obstacle(tiles[ty * ncols + tx], path_width, fx, fy, expansion) ||
                            is_placeable_obstacle(static_cast<const Obstacle *>(obstacles), n_obstacles, polygon_points,
                                                  tx, ty, fx, fy, expansion);
      img[size_t(y) * W + x] = obstacle ? 0 : 255;
    }
  const float blur_pixels = blur_percent * float(res);
  if (!(blur_pixels < 1.0f)) {
    const float sigma = blur_pixels <= 0.0f ? 1.0f : blur_pixels;
    const float support = 2.0f * sigma;
    // vertical_sample -> f32 image
    std::vector<float> tmp(size_t(W) * H);
    std::vector<float> ws;
    for (uint32_t oy = 0; oy < H; ++oy) {
      float c = (float(oy) + 0.5f) * 1.0f;
      const int64_t left = clampi(int64_t(std::floor(c - support)), 0, int64_t(H) - 1);
      const int64_t right = clampi(int64_t(std::ceil(c + support)), left + 1, int64_t(H));
      c = c - 0.5f;
      ws.clear();
      float sum = 0.0f;
      for (int64_t i = left; i < right; ++i) {
        const float w = gaussian((float(i) - c) / 1.0f, sigma);
        ws.push_back(w);
        sum += w;
      }
      for (float &w : ws) w /= sum;
      for (uint32_t x = 0; x < W; ++x) {
        float t = 0.0f;
        for (size_t k = 0; k < ws.size(); ++k) t += float(img[size_t(left + int64_t(k)) * W + x]) * ws[k];
        tmp[size_t(oy) * W + x] = t;
      }
    }
    // horizontal_sample -> u8, clamp + round half away from zero (FloatNearest)
    for (uint32_t ox = 0; ox < W; ++ox) {
      float c = (float(ox) + 0.5f) * 1.0f;
      const int64_t left = clampi(int64_t(std::floor(c - support)), 0, int64_t(W) - 1);
      const int64_t right = clampi(int64_t(std::ceil(c + support)), left + 1, int64_t(W));
      c = c - 0.5f;
      ws.clear();
      float sum = 0.0f;
      for (int64_t i = left; i < right; ++i) {
        const float w = gaussian((float(i) - c) / 1.0f, sigma);
        ws.push_back(w);
        sum += w;
      }
      for (float &w : ws) w /= sum;
      for (uint32_t y = 0; y < H; ++y) {
        float t = 0.0f;
        for (size_t k = 0; k < ws.size(); ++k) t += tmp[size_t(y) * W + size_t(left + int64_t(k))] * ws[k];
        const float cl = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
        img[size_t(y) * W + ox] = uint8_t(std::round(cl));
      }
    }
  }
  for (size_t k = 0; k < img.size(); ++k) out_rgb[3 * k] = out_rgb[3 * k + 1] = out_rgb[3 * k + 2] = img[k];
  return 0;
}

void *gbpo_create(const void *cfg) {
  World *w = new World();
  std::memcpy(&w->cfg, cfg, sizeof(Cfg));
  w->sdf.w = 1;
  w->sdf.h = 1;
  w->sdf.rgb = {255, 255, 255};
  return w;
}
int gbpo_config_size(void) { return int(sizeof(Cfg)); }
void gbpo_destroy(void *p) { delete static_cast<World *>(p); }
void gbpo_set_threads(void *p, int t) { static_cast<World *>(p)->threads = t < 1 ? 1 : t; }

int gbpo_set_sdf(void *p, const uint8_t *rgb, int w_, int h_) {
  World *w = static_cast<World *>(p);
  if (!w->robots.empty()) return -5;  // obstacle factors capture jac_delta at creation
  w->sdf.w = w_;
  w->sdf.h = h_;
  w->sdf.rgb.assign(rgb, rgb + size_t(w_) * h_ * 3);
  return 0;
}
int gbpo_add_robots(void *p, int n, const float *radii, const uint32_t *timesteps,
                    const double *means, const float *pos, const int32_t *wp_off,
                    const float *wp_xy) {
  World *w = static_cast<World *>(p);
  int V = w->cfg.num_variables;
  w->timesteps.assign(timesteps, timesteps + V);
  w->robots.reserve(w->robots.size() + n);
  for (int i = 0; i < n; ++i)
    add_robot(*w, radii[i], means + size_t(i) * V * DOFS, pos + 2 * i, wp_xy + 2 * wp_off[i],
              wp_off[i + 1] - wp_off[i]);
  return 0;
}
int gbpo_num_robots(void *p) { return int(static_cast<World *>(p)->robots.size()); }

int gbpo_update_topology(void *p) {
  World *w = static_cast<World *>(p);
  update_neighbours(*w);
  delete_interrobot_factors(*w);
  create_interrobot_factors(*w);
  return 0;
}
int gbpo_set_comms(void *p, const uint8_t *antenna, const uint8_t *idle) {
  World *w = static_cast<World *>(p);
  for (size_t i = 0; i < w->robots.size(); ++i) {
    w->robots[i].antenna = antenna ? antenna[i] != 0 : true;
    w->robots[i].idle = idle ? idle[i] != 0 : false;
  }
  return 0;
}
// RobotDespawned: the entity (FactorGraph, RobotConnections, Transform) is gone.
int gbpo_remove_robots(void *p, int m, const int32_t *robots) {
  World *w = static_cast<World *>(p);
  for (int k = 0; k < m; ++k) {
    if (robots[k] < 0 || size_t(robots[k]) >= w->robots.size()) return -1;
    w->robots[robots[k]].gone = true;
  }
  return 0;
}

int gbpo_set_waypoint_index(void *p, const int32_t *idx) {
  World *w = static_cast<World *>(p);
  for (size_t i = 0; i < w->robots.size(); ++i) w->robots[i].next_wp = idx[i];
  return 0;
}

// FactorGraph::change_prior_of_variable (factorgraph.rs:494-528) + the caller's
// deferred delivery to other graphs' factors (robot.rs:2272-2282).
static void change_prior_of_variable(World &w, int robot, int var, const Vec &mean,
                                     std::vector<Routed> &deferred) {
  Graph &g = w.robots[robot].g;
  auto msgs = change_prior(g.vars[var], mean);
  for (auto &m : msgs) {
    if (m.first.first == g.id) {
      auto it = g.factors.find(m.first.second);
      if (it != g.factors.end()) factor_receive(it->second, {g.id, var}, m.second);
    } else {
      deferred.push_back({{g.id, var}, m.first, m.second});
    }
  }
}
static void deliver_to_factors(World &w, std::vector<Routed> &msgs) {
  for (auto &m : msgs) {
    if (w.robots[m.to.first].gone) continue;  // query.get_mut fails for a despawned robot (robot.rs:2273-2277)
    auto &fs = w.robots[m.to.first].g.factors;
    auto it = fs.find(m.to.second);
    if (it != fs.end()) factor_receive(it->second, m.from, m.m);
  }
}

// update_prior_of_horizon_state (planner/robot.rs:2182-2283).
int gbpo_update_prior_of_horizon_state(void *p) {
  World *w = static_cast<World *>(p);
  double delta_t = double(w->cfg.delta_t);
  double max_speed = double(w->cfg.target_speed);
  std::vector<Routed> deferred;
  for (auto &r : w->robots) {
    if (r.finished || r.idle || r.gone) continue;
    if (r.next_wp < 0 || r.next_wp >= int(r.waypoints.size())) {
      r.finished = true;
      continue;
    }
    if (w->cfg.iterations_internal == 0) continue;
    Variable &hv = r.g.vars.back();
    Vec wp{double(r.waypoints[r.next_wp].first), double(r.waypoints[r.next_wp].second)};
    Vec h2w{wp[0] - hv.mu[0], wp[1] - hv.mu[1]};
    double dist = euclidean_norm(h2w);
    Vec dir = normalized(h2w);
    double sp = std::fmin(max_speed, dist);
    Vec nv{sp * dir[0], sp * dir[1]};
    Vec nm{hv.mu[0] + nv[0] * delta_t, hv.mu[1] + nv[1] * delta_t, nv[0], nv[1]};
    hv.mu = nm;
    change_prior_of_variable(*w, r.g.id, int(r.g.vars.size()) - 1, nm, deferred);
  }
  deliver_to_factors(*w, deferred);
  return 0;
}
// reached_waypoint (planner/robot.rs:2080-2176) for single-route missions: the chosen variable's
// estimated position as an f32 Vec2 (variable.rs:127-130) against the next waypoint, glam
// Vec2::distance_squared in f32; Route::advance (robot.rs:431-438) on success.
// crit = {intersects_with (0 Current, 1 Horizon, 2 Variable(ix)), ix, distance kind (0 RobotRadius, 1 Meter)},
// crit[0..2] for ordinary waypoints (taskpoint_reached_when_intersects), crit[3..5] for the last one
// (finished_when_intersects); meters[0..1] likewise.  out_reached[n] (optional) flags the robots that advanced.
int gbpo_reached_waypoint(void *p, const int32_t *crit, const float *meters, uint8_t *out_reached) {
  World *w = static_cast<World *>(p);
  for (size_t k = 0; k < w->robots.size(); ++k) {
    Robot &r = w->robots[k];
    if (out_reached) out_reached[k] = 0;
    if (r.gone) continue;
    const int nwp = int(r.waypoints.size());
    if (r.next_wp < 0 || r.next_wp >= nwp) continue;  // mission.next_waypoint() is None
    const bool last = r.next_wp == nwp - 1;
    const int32_t *c = crit + (last ? 3 : 0);
    const float meter = meters[last ? 1 : 0];
    const int V = int(r.g.vars.size());
    int var = c[0] == 0 ? 0 : (c[0] == 1 ? V - 1 : (c[1] >= 0 && c[1] < V ? c[1] : V - 1));
    const Variable &v = r.g.vars[var];
    const float ex = float(v.mu[0]), ey = float(v.mu[1]);
    const float dsq = c[2] == 0 ? r.radius * r.radius : meter * meter;
    const float dx = ex - r.waypoints[r.next_wp].first, dy = ey - r.waypoints[r.next_wp].second;
    const float d2 = dx * dx + dy * dy;
    if (d2 < dsq) {
      r.next_wp += 1;
      if (out_reached) out_reached[k] = 1;
    }
  }
  return 0;
}
// Ball::aabb(&Isometry2::translation(x, z)) (parry2d shape/ball.rs -> bounding_volume::ball_aabb): centre -+ radius.
static void ball_aabb(const float pos[2], float radius, float out[4]) {
  out[0] = pos[0] - radius;
  out[1] = pos[1] - radius;
  out[2] = pos[0] + radius;
  out[3] = pos[1] + radius;
}
// Aabb::intersection (bounding_volume/aabb.rs): sup of the mins, inf of the maxs (None when they cross; the reference
// unwraps / expects right after a positive intersection test, so the box is kept as computed).
static void aabb_intersection(const float a[4], const float b[4], float out[4]) {
  out[0] = std::max(a[0], b[0]);
  out[1] = std::max(a[1], b[1]);
  out[2] = std::min(a[2], b[2]);
  out[3] = std::min(a[3], b[3]);
}
// Collider::aabb = shape.compute_aabb(&isometry) (gbp_global_planner/src/lib.rs:94-98; parry2d, PARITY UNPINNED):
// Ball: translation -+ radius; Cuboid: translation -+ |R| half_extents; Triangle / ConvexPolygon: min / max over the
// vertices moved by the isometry (rotation (re x - im y, im x + re y), then the translation).
static void collider_aabb(const EnvCollider &c, float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  if (c.kind == 0 || c.kind == 1) {
    float hx = c.radius, hy = c.radius;
    if (c.kind == 1) {
      hx = std::fabs(c.re) * c.half_extents.x + std::fabs(c.im) * c.half_extents.y;
      hy = std::fabs(c.im) * c.half_extents.x + std::fabs(c.re) * c.half_extents.y;
    }
    out[0] = c.translation.x - hx;
    out[1] = c.translation.y - hy;
    out[2] = c.translation.x + hx;
    out[3] = c.translation.y + hy;
    return;
  }
  for (size_t k = 0; k < c.points.size(); ++k) {
    const V2 p = c.points[k];
    const float x = (c.re * p.x - c.im * p.y) + c.translation.x, y = (c.im * p.x + c.re * p.y) + c.translation.y;
    out[0] = k ? std::min(out[0], x) : x;
    out[1] = k ? std::min(out[1], y) : y;
    out[2] = k ? std::max(out[2], x) : x;
    out[3] = k ? std::max(out[3], y) : y;
  }
}
// update_robot_robot_collisions (planner/collisions.rs:72-143): ALL pairs (r < c) in robot order;
// parry2d BoundingSphere::intersects (un-vendored crate, restated: |c_b - c_a|^2 <= (r_a + r_b)^2 in
// f32); CollisionHistory::update (:472-488): Free -> Colliding is a Hit.
int gbpo_update_robot_collisions(void *p, int64_t *num_collisions, int64_t *colliding_now, uint32_t *per_robot) {
  World *w = static_cast<World *>(p);
  const int n = int(w->robots.size());
  w->coll_hits.resize(size_t(n), 0u);
  int64_t now_count = 0;
  for (int r = 0; r < n; ++r)
    for (int c = r + 1; c < n; ++c) {
      const Robot &a = w->robots[r], &b = w->robots[c];
      if (a.gone || b.gone) continue;  // the query iterates living entities only
      const float dx = b.pos[0] - a.pos[0], dz = b.pos[1] - a.pos[1];
      const float d2 = dx * dx + dz * dz;
      const float sr = a.radius + b.radius;
      const bool now = d2 <= sr * sr;
      bool &state = w->coll_state[{r, c}];
      if (now && !state) {
        w->collisions += 1;
        w->coll_hits[r] += 1;
        w->coll_hits[c] += 1;
        // :117-138: r_aabb.intersection(&c_aabb) travels in the event and lands in the pair's history (:190-196)
        World::CollEvent ev{r, c, {0, 0, 0, 0}};
        float ra[4], ca[4];
        ball_aabb(a.pos, a.radius, ra);
        ball_aabb(b.pos, b.radius, ca);
        aabb_intersection(ra, ca, ev.aabb);
        w->robot_events.push_back(ev);
      }
      state = now;
      if (now) ++now_count;
    }
  if (num_collisions) *num_collisions = w->collisions;
  if (colliding_now) *colliding_now = now_count;
  if (per_robot)
    for (int r = 0; r < n; ++r) per_robot[r] = w->coll_hits[r];
  return 0;
}
int gbpo_read_waypoint_index(void *p, int32_t *out) {
  World *w = static_cast<World *>(p);
  for (size_t k = 0; k < w->robots.size(); ++k) out[k] = w->robots[k].next_wp;
  return 0;
}

// update_prior_of_current_state_v3 (planner/robot.rs:2286-2338).
int gbpo_update_prior_of_current_state(void *p) {
  World *w = static_cast<World *>(p);
  std::vector<Routed> deferred;
  for (auto &r : w->robots) {
    if (r.idle || r.gone) continue;
    float time_scale = w->cfg.delta_t / r.t0;
    Variable &c = r.g.vars[0];
    Variable &nx = r.g.vars[1];
    Vec change(DOFS), upd(DOFS);
    for (int k = 0; k < DOFS; ++k) change[k] = double(time_scale) * (nx.mu[k] - c.mu[k]);
    for (int k = 0; k < DOFS; ++k) upd[k] = c.mu[k] + change[k];
    change_prior_of_variable(*w, r.g.id, 0, upd, deferred);
    r.pos[0] += float(change[0]);
    r.pos[1] += float(change[1]);
  }
  return 0;
}
int gbpo_change_prior_of_variable(void *p, int var, int m, const int32_t *robots,
                                  const double *means) {
  World *w = static_cast<World *>(p);
  std::vector<Routed> deferred;
  for (int k = 0; k < m; ++k)
    change_prior_of_variable(*w, robots[k], var, Vec(means + 4 * k, means + 4 * k + 4), deferred);
  deliver_to_factors(*w, deferred);
  return 0;
}

// ---- global-planner hand-off: what `update_robot_mission` does to the factor graph
// when an RRT* path arrives (planner/robot.rs:655-776).
// FactorGraph::modify_tracking_factors(|t| t.set_tracking_path(..)) (factorgraph.rs:1467-1477,
// factor/tracking.rs:134-136) + Route::update_waypoints (robot.rs:389-392): the tracking path of every
// tracking factor and the mission's waypoints become the new polyline, target_index = 1.
int gbpo_set_tracking_path(void *p, int m, const int32_t *robots, const int32_t *wp_offsets, const float *wp_xy) {
  World *w = static_cast<World *>(p);
  for (int k = 0; k < m; ++k) {
    Robot &r = w->robots[robots[k]];
    std::vector<std::pair<float, float>> path;
    for (int q = wp_offsets[k]; q < wp_offsets[k + 1]; ++q) path.emplace_back(wp_xy[2 * q], wp_xy[2 * q + 1]);
    if (path.size() < 2) return -2;  // min_len_vec::TwoOrMore
    for (auto &kv : r.g.factors)
      if (kv.second.kind == TRACKING) kv.second.path = path;
    r.waypoints = path;
    r.next_wp = 1;
  }
  return 0;
}
// FactorGraph::reset_variables (factorgraph.rs:1541-1564): VariableNode::reset (variable.rs:350-360) for
// every variable — mean and belief precision replaced, every inbox entry emptied — then
// FactorNode::empty_inbox (factor/mod.rs:480-483) for every factor of the graph.
int gbpo_reset_variables(void *p, int m, const int32_t *robots, const double *means, double first_last_sigma,
                         double inbetween_sigma) {
  World *w = static_cast<World *>(p);
  for (int k = 0; k < m; ++k) {
    Graph &g = w->robots[robots[k]].g;
    const int V = int(g.vars.size());
    for (int i = 0; i < V; ++i) {
      Variable &v = g.vars[i];
      const double sigma = (i == 0 || i == V - 1) ? first_last_sigma : inbetween_sigma;
      for (int a = 0; a < DOFS; ++a) v.mu[a] = means[(size_t(k) * V + i) * DOFS + a];
      for (int a = 0; a < DOFS; ++a)
        for (int b = 0; b < DOFS; ++b) v.lam(a, b) = (a == b) ? sigma : 0.0;  // Matrix::from_diag_elem
      for (auto &kv : v.inbox) kv.second = Message::empty();
    }
    for (auto &kv : g.factors)
      for (auto &in : kv.second.inbox) in.second = Message::empty();
  }
  return 0;
}
// FactorGraph::reset_tracking_factors (factorgraph.rs:1566-1590): set_timeout(10) on the tracking factor of
// every variable but the first and the last.
int gbpo_reset_tracking_factors(void *p, int m, const int32_t *robots) {
  World *w = static_cast<World *>(p);
  for (int k = 0; k < m; ++k) {
    Graph &g = w->robots[robots[k]].g;
    const int V = int(g.vars.size());
    for (auto &kv : g.factors) {
      Factor &f = kv.second;
      if (f.kind == TRACKING && f.own_var >= 1 && f.own_var <= V - 2) f.timeout = 10;
    }
  }
  return 0;
}

int gbpo_internal_factor_iteration(void *p) { world_internal(*static_cast<World *>(p), true, false); return 0; }
int gbpo_internal_variable_iteration(void *p) { world_internal(*static_cast<World *>(p), false, true); return 0; }
int gbpo_external_factor_iteration(void *p) { world_external_factor(*static_cast<World *>(p)); return 0; }
int gbpo_external_variable_iteration(void *p) { world_external_variable(*static_cast<World *>(p)); return 0; }

// iterate_gbp_v2 (planner/robot.rs:1769-1861).
int gbpo_iterate_schedule(void *p, int n, const uint8_t *internal, const uint8_t *external) {
  World *w = static_cast<World *>(p);
  for (int s = 0; s < n; ++s) {
    if (internal[s]) world_internal(*w, true, true);
    if (external[s]) {
      world_external_factor(*w);
      world_external_variable(*w);
    }
  }
  return 0;
}
int gbpo_iterate(void *p) {
  World *w = static_cast<World *>(p);
  uint8_t oi[256], oe[256];
  int n = gbpo_schedule(w->cfg.schedule_kind, uint8_t(w->cfg.iterations_internal),
                        uint8_t(w->cfg.iterations_external), oi, oe);
  if (n < 0) return n;
  return gbpo_iterate_schedule(p, n, oi, oe);
}
int gbpo_step(void *p) {  // FixedUpdate chain (robot.rs:85-108), comms mask supplied by caller
  gbpo_update_topology(p);
  gbpo_update_prior_of_horizon_state(p);
  gbpo_update_prior_of_current_state(p);
  return gbpo_iterate(p);
}

int gbpo_change_factor_enabled(void *p, int kind, uint8_t en) {  // factorgraph.rs:1529-1539
  World *w = static_cast<World *>(p);
  for (auto &r : w->robots)
    for (auto &kv : r.g.factors)
      if (int(kv.second.kind) == kind) kv.second.enabled = en;
  uint8_t *flags[4] = {&w->cfg.enable_dynamic, &w->cfg.enable_interrobot, &w->cfg.enable_obstacle,
                       &w->cfg.enable_tracking};
  if (kind >= 0 && kind < 4) *flags[kind] = en;
  return 0;
}
int gbpo_set_safety_distance_multiplier(void *p, float mult) {  // factorgraph.rs:892
  World *w = static_cast<World *>(p);
  w->cfg.safety_distance_multiplier = mult;
  for (auto &r : w->robots)
    for (auto &kv : r.g.factors)
      if (kv.second.kind == INTERROBOT)
        kv.second.safety_distance = double(mult) * kv.second.robot_radius;
  return 0;
}
int gbpo_set_schedule(void *p, int kind, int internal, int external) {
  World *w = static_cast<World *>(p);
  w->cfg.schedule_kind = kind;
  w->cfg.iterations_internal = internal;
  w->cfg.iterations_external = external;
  return 0;
}

int gbpo_read_beliefs(void *p, double *eta, double *lam, double *mean, double *cov,
                      uint8_t *valid) {
  World *w = static_cast<World *>(p);
  size_t k = 0;
  for (auto &r : w->robots)
    for (auto &v : r.g.vars) {
      if (eta) std::copy(v.eta.begin(), v.eta.end(), eta + 4 * k);
      if (lam) std::copy(v.lam.a.begin(), v.lam.a.end(), lam + 16 * k);
      if (mean) std::copy(v.mu.begin(), v.mu.end(), mean + 4 * k);
      if (cov) std::copy(v.cov.a.begin(), v.cov.a.end(), cov + 16 * k);
      if (valid) valid[k] = v.valid;
      ++k;
    }
  return 0;
}
int gbpo_read_positions(void *p, float *xy) {
  World *w = static_cast<World *>(p);
  for (size_t i = 0; i < w->robots.size(); ++i) {
    xy[2 * i] = w->robots[i].pos[0];
    xy[2 * i + 1] = w->robots[i].pos[1];
  }
  return 0;
}
int64_t gbpo_read_connections(void *p, int64_t *offsets, int32_t *nbrs, int64_t *robot_number,
                              int64_t cap) {
  World *w = static_cast<World *>(p);
  int64_t e = 0;
  for (size_t i = 0; i < w->robots.size(); ++i) {
    Robot &r = w->robots[i];
    offsets[i] = e;
    for (int o : r.connected) {
      if (e >= cap) return -2;
      nbrs[e] = o;
      uint64_t first = 0;
      for (auto &kv : r.g.factors)
        if (kv.second.kind == INTERROBOT && kv.second.ext_robot == o && kv.second.ext_var == 1)
          first = kv.second.robot_number;
      robot_number[e] = int64_t(first);
      ++e;
    }
  }
  offsets[w->robots.size()] = e;
  return e;
}
int gbpo_sdf_lookup(void *p, int m, const double *xy, uint32_t *px, uint32_t *py, double *val) {
  World *w = static_cast<World *>(p);
  for (int i = 0; i < m; ++i)
    val[i] = sdf_measure(w->sdf, w->cfg.world_width, w->cfg.world_height, xy[2 * i], xy[2 * i + 1],
                         px + i, py + i);
  return 0;
}
int gbpo_node_counts(void *p, int64_t out[5]) {
  World *w = static_cast<World *>(p);
  for (int i = 0; i < 5; ++i) out[i] = 0;
  for (auto &r : w->robots) {
    if (r.gone) continue;  // despawned with its FactorGraph component
    out[0] += int64_t(r.g.vars.size());
    for (auto &kv : r.g.factors) {
      switch (kv.second.kind) {
        case DYNAMIC: out[1]++; break;
        case OBSTACLE: out[2]++; break;
        case TRACKING: out[3]++; break;
        case INTERROBOT: out[4]++; break;
      }
    }
  }
  return 0;
}
// Inbox introspection for white-box tests: the message a variable holds from
// the mirror InterRobot factor owned by `from_robot` (returns 0 if Empty/absent).
int gbpo_read_mirror_message(void *p, int robot, int var, int from_robot, double *eta,
                             double *lam) {
  World *w = static_cast<World *>(p);
  for (auto &kv : w->robots[robot].g.vars[var].inbox) {
    if (kv.first.first != from_robot) continue;
    if (kv.second.is_empty()) return 0;
    std::copy(kv.second.payload->eta.begin(), kv.second.payload->eta.end(), eta);
    std::copy(kv.second.payload->lam.a.begin(), kv.second.payload->lam.a.end(), lam);
    return 1;
  }
  return 0;
}
// FactorGraph::messages_sent / messages_received (factorgraph.rs:876-890): sums over the nodes the graph holds NOW
// (a deleted InterRobot factor takes its counters with it).  out[4 * r + k]: sent.internal, sent.external,
// received.internal, received.external of robot r; zeros for a despawned robot.
int gbpo_read_message_counts(void *p, int64_t *out) {
  World *w = static_cast<World *>(p);
  for (size_t r = 0; r < w->robots.size(); ++r) {
    MsgCount t;
    if (!w->robots[r].gone) {
      for (auto &v : w->robots[r].g.vars) {
        t.sent_int += v.cnt.sent_int; t.sent_ext += v.cnt.sent_ext;
        t.recv_int += v.cnt.recv_int; t.recv_ext += v.cnt.recv_ext;
      }
      for (auto &kv : w->robots[r].g.factors) {
        const MsgCount &c = kv.second.cnt;
        t.sent_int += c.sent_int; t.sent_ext += c.sent_ext; t.recv_int += c.recv_int; t.recv_ext += c.recv_ext;
      }
    }
    out[4 * r + 0] = int64_t(t.sent_int);
    out[4 * r + 1] = int64_t(t.sent_ext);
    out[4 * r + 2] = int64_t(t.recv_int);
    out[4 * r + 3] = int64_t(t.recv_ext);
  }
  return 0;
}
int gbpo_has_mirror_slot(void *p, int robot, int var, int from_robot) {
  World *w = static_cast<World *>(p);
  for (auto &kv : w->robots[robot].g.vars[var].inbox)
    if (kv.first.first == from_robot) return 1;
  return 0;
}
// Tracking factor state of robot r, variable i: record, last_pos (f32), last_value.
int gbpo_read_tracking(void *p, int robot, int var, int64_t *record, float *pos, double *value) {
  World *w = static_cast<World *>(p);
  for (auto &kv : w->robots[robot].g.factors) {
    Factor &f = kv.second;
    if (f.kind != TRACKING || f.own_var != var) continue;
    *record = int64_t(f.trk_record);
    pos[0] = f.last_pos[0];
    pos[1] = f.last_pos[1];
    *value = f.last_value;
    return 1;
  }
  return 0;
}

// Colliders resource; clears RobotEnvironmentCollisions (collisions.rs:40-46).  cols: [n][9] floats
// (kind, tx, ty, angle, radius, hx, hy, first_vertex, num_vertices).
int gbpo_set_environment_colliders(void *p, int n, const float *cols, int nverts, const float *verts) {
  World *w = static_cast<World *>(p);
  w->colliders.clear();
  w->env_state.clear();
  w->env_hits.clear();
  w->env_collisions = 0;
  w->env_events.clear();
  for (int k = 0; k < n; ++k) {
    const float *c = cols + 9 * k;
    EnvCollider e;
    e.kind = int(c[0]);
    e.translation = {c[1], c[2]};
    e.re = std::cos(c[3]);  // UnitComplex::new(angle): f32 sin_cos
    e.im = std::sin(c[3]);
    e.radius = c[4];
    e.half_extents = {c[5], c[6]};
    const int v0 = int(c[7]), nv = int(c[8]);
    if (e.kind >= 2 && (v0 < 0 || nv < 3 || v0 + nv > nverts)) return -2;
    for (int q = 0; q < (e.kind >= 2 ? nv : 0); ++q) e.points.push_back({verts[2 * (v0 + q)], verts[2 * (v0 + q) + 1]});
    w->colliders.push_back(e);
  }
  return 0;
}
// update_robot_environment_collisions (planner/collisions.rs:368-431): every living robot against every collider.
int gbpo_update_environment_collisions(void *p, int64_t *num_collisions, int64_t *colliding_now, uint32_t *per_robot) {
  World *w = static_cast<World *>(p);
  const int n = int(w->robots.size());
  w->env_hits.resize(size_t(n), 0u);
  int64_t now_count = 0;
  for (int r = 0; r < n; ++r) {
    const Robot &rb = w->robots[r];
    if (rb.gone) continue;
    for (int c = 0; c < int(w->colliders.size()); ++c) {
      const bool now = intersection_test(w->colliders[c], {rb.pos[0], rb.pos[1]}, rb.radius);
      bool &state = w->env_state[{r, c}];
      if (now && !state) {  // CollisionStatus::Hit
        w->env_collisions += 1;
        w->env_hits[r] += 1;
        // :417-426: robot_aabb.intersection(&env_aabb)
        World::CollEvent ev{r, c, {0, 0, 0, 0}};
        float ra[4], ca[4];
        ball_aabb(rb.pos, rb.radius, ra);
        collider_aabb(w->colliders[c], ca);
        aabb_intersection(ra, ca, ev.aabb);
        w->env_events.push_back(ev);
      }
      state = now;
      if (now) ++now_count;
    }
  }
  if (num_collisions) *num_collisions = w->env_collisions;
  if (colliding_now) *colliding_now = now_count;
  if (per_robot)
    for (int r = 0; r < n; ++r) per_robot[r] = w->env_hits[r];
  return 0;
}
// Every Hit so far in the order the update systems produced them: kind 0 robot-robot (pairs = (r, c), r < c), kind 1
// robot-environment (pairs = (robot, collider)); aabbs [k][4].  Returns the number of events (fills at most `cap`).
int64_t gbpo_read_collision_events(void *p, int kind, int64_t cap, int32_t *pairs, float *aabbs) {
  World *w = static_cast<World *>(p);
  const std::vector<World::CollEvent> &ev = kind == 0 ? w->robot_events : w->env_events;
  for (int64_t k = 0; k < int64_t(ev.size()) && k < cap; ++k) {
    if (pairs) {
      pairs[2 * k] = ev[size_t(k)].a;
      pairs[2 * k + 1] = ev[size_t(k)].b;
    }
    if (aabbs) std::memcpy(aabbs + 4 * k, ev[size_t(k)].aabb, sizeof(float) * 4);
  }
  return int64_t(ev.size());
}
int gbpo_read_removed(void *p, uint8_t *out) {
  World *w = static_cast<World *>(p);
  for (size_t k = 0; k < w->robots.size(); ++k) out[k] = w->robots[k].gone ? 1 : 0;
  return 0;
}
int gbpo_set_tracking_buffers(void *p, int capacity, uint64_t sample_ns) {
  World *w = static_cast<World *>(p);
  if (capacity < 1) return -2;
  w->track_capacity = capacity;
  w->track_duration_ns = sample_ns;
  w->trackers.clear();
  w->track_seen = 0;
  return 0;
}
// track_positions + track_velocities (planner/tracking.rs:117-137, 226-260) for one FixedUpdate.
int gbpo_track(void *p, uint64_t delta_ns, double now) {
  World *w = static_cast<World *>(p);
  if (w->track_capacity < 1) return -5;
  w->trackers.resize(w->robots.size());
  const size_t cap = size_t(w->track_capacity);
  for (size_t r = 0; r < w->robots.size(); ++r) {
    const Robot &rb = w->robots[r];
    if (rb.gone) continue;
    if (rb.idle && r < w->track_seen) continue;  // Changed<Transform>: update_prior_of_current_state skipped it
    Tracker &t = w->trackers[r];
    // bevy_time Timer::tick, TimerMode::Repeating
    t.elapsed_ns += delta_ns;
    const bool finished = t.elapsed_ns >= w->track_duration_ns;
    if (finished) t.elapsed_ns = w->track_duration_ns ? t.elapsed_ns % w->track_duration_ns : 0;
    if (!finished) continue;
    if (t.positions.size() == cap) t.positions.erase(t.positions.begin());  // push_overwrite
    t.positions.push_back({rb.pos[0], rb.pos[1]});
    t.npos += 1;
    if (t.has_prev) {
      const double dt = now - t.prev_t;
      Tracker::Vel v;
      v.v[0] = (rb.pos[0] - t.prev[0]) / float(dt);
      v.v[1] = (rb.pos[1] - t.prev[1]) / float(dt);
      v.timestamp = now;
      v.measured_over = dt;
      if (t.velocities.size() == cap) t.velocities.erase(t.velocities.begin());
      t.velocities.push_back(v);
      t.nvel += 1;
    }
    t.prev[0] = rb.pos[0];
    t.prev[1] = rb.pos[1];
    t.prev_t = now;
    t.has_prev = true;
  }
  w->track_seen = w->robots.size();
  return 0;
}
// Ring contents of one robot, oldest first; returns the number of position samples, *nvel the velocity samples.
int gbpo_read_track(void *p, int robot, float *pos_xy, int *nvel, float *vel_xy, double *vel_t, double *vel_over) {
  World *w = static_cast<World *>(p);
  *nvel = 0;
  if (size_t(robot) >= w->trackers.size()) return 0;
  const Tracker &t = w->trackers[robot];
  for (size_t k = 0; k < t.positions.size(); ++k) {
    pos_xy[2 * k] = t.positions[k][0];
    pos_xy[2 * k + 1] = t.positions[k][1];
  }
  for (size_t k = 0; k < t.velocities.size(); ++k) {
    vel_xy[2 * k] = t.velocities[k].v[0];
    vel_xy[2 * k + 1] = t.velocities[k].v[1];
    vel_t[k] = t.velocities[k].timestamp;
    vel_over[k] = t.velocities[k].measured_over;
  }
  *nvel = int(t.velocities.size());
  return int(t.positions.size());
}

}  // extern "C"
