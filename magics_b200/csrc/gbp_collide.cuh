// gbp_collide.cuh — the evaluation outputs next to the iteration: robot-environment
// collisions (planner/collisions.rs:368-455) and the position / velocity sample buffers
// (planner/tracking.rs:117-260).
//
// update_robot_environment_collisions tests every robot's Ball against every environment collider with
// parry2d::query::intersection_test(&collider.isometry, collider.shape, &robot_pos, ball) and feeds the
// result into the per-(robot, collider) CollisionHistory (Free / Colliding, `times` += 1 on Free -> Colliding,
// collisions.rs:455-493).  parry2d is a third-party crate (0.13.7, the AU-Master-Thesis fork with Bevy
// conversions, Cargo.lock:5372-5374) that is not in the reference tree; what is restated here is its published
// algorithm for this call, in f32 with the operations in its order and no fused multiply-add:
//   pos12 = pos1.inv_mul(pos2)              translation R1^-1 (t2 - t1)                       (nalgebra Isometry)
//   Ball / Ball                             |c12|^2 <= (r1 + r2)^2                            (intersection_test_ball_ball)
//   any other shape / Ball                  proj = shape.project_local_point(c12, solid = true);
//                                           proj.is_inside || |c12 - proj.point|^2 <= r^2     (..._point_query_ball)
//   Cuboid projection                       Aabb::project_local_point: shift = max(mins - p, 0) - max(p - maxs, 0)
//   Triangle projection                     Ericson's Voronoi-region walk (2-D: a point on the face is inside)
//   ConvexPolygon projection                parry runs GJK on the support map; restated as the exact geometric
//                                           predicate it converges to: inside all edges, else the nearest edge point
// Parity is therefore unpinned for this row: engine == oracle bit for bit, oracle == parry2d up
// to rounding at the boundary of a shape.
#pragma once
#include <cstdint>

#ifndef GBP_DEV  // gbp_collide_host.cpp compiles the predicate below for the host and defines GBP_DEV itself
#include "gbp_math.cuh"
#endif

namespace gbp {

enum ColliderKind : int32_t { kColliderBall = 0, kColliderCuboid = 1, kColliderTriangle = 2, kColliderConvexPolygon = 3 };

// One environment collider as the kernels read it (the host turned the angle into UnitComplex::new(angle)).
struct ColliderDev {
  int32_t kind;
  float tx, ty;    // isometry.translation
  float re, im;    // isometry.rotation = (cos angle, sin angle)
  float radius;    // Ball
  float hx, hy;    // Cuboid half extents
  int32_t v0, nv;  // Triangle (nv = 3) / ConvexPolygon: vertices [v0, v0 + nv) of the vertex array, counter-clockwise
};

GBP_DEV float cl_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
GBP_DEV float cl_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
GBP_DEV float cl_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
GBP_DEV float cl_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
GBP_DEV float cl_dot(float ax, float ay, float bx, float by) { return cl_add(cl_mul(ax, bx), cl_mul(ay, by)); }
GBP_DEV float cl_perp(float ax, float ay, float bx, float by) { return cl_sub(cl_mul(ax, by), cl_mul(ay, bx)); }

// |p - segment(a, b)|^2 and whether p lies on the inner side (left of a -> b).
GBP_DEV float cl_segment_dist2(float px, float py, float ax, float ay, float bx, float by, bool &left) {
  const float ex = cl_sub(bx, ax), ey = cl_sub(by, ay), wx = cl_sub(px, ax), wy = cl_sub(py, ay);
  left = cl_perp(ex, ey, wx, wy) >= 0.0f;
  const float ee = cl_dot(ex, ey, ex, ey);
  float t = ee > 0.0f ? cl_div(cl_dot(wx, wy, ex, ey), ee) : 0.0f;
  t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  const float qx = cl_add(ax, cl_mul(ex, t)), qy = cl_add(ay, cl_mul(ey, t));
  const float dx = cl_sub(px, qx), dy = cl_sub(py, qy);
  return cl_dot(dx, dy, dx, dy);
}

// Triangle::project_local_point (solid): is_inside and |p - proj|^2.
GBP_DEV bool cl_triangle(float px, float py, const float *v, float &d2) {
  const float ax = v[0], ay = v[1], bx = v[2], by = v[3], cx = v[4], cy = v[5];
  const float abx = cl_sub(bx, ax), aby = cl_sub(by, ay), acx = cl_sub(cx, ax), acy = cl_sub(cy, ay);
  const float apx = cl_sub(px, ax), apy = cl_sub(py, ay);
  const float ab_ap = cl_dot(abx, aby, apx, apy), ac_ap = cl_dot(acx, acy, apx, apy);
  auto to = [&](float qx, float qy) {
    const float dx = cl_sub(px, qx), dy = cl_sub(py, qy);
    d2 = cl_dot(dx, dy, dx, dy);
    return false;
  };
  if (ab_ap <= 0.0f && ac_ap <= 0.0f) return to(ax, ay);  // Voronoi region of a
  const float bpx = cl_sub(px, bx), bpy = cl_sub(py, by);
  const float ab_bp = cl_dot(abx, aby, bpx, bpy), ac_bp = cl_dot(acx, acy, bpx, bpy);
  if (ab_bp >= 0.0f && ac_bp <= ab_bp) return to(bx, by);
  const float cpx = cl_sub(px, cx), cpy = cl_sub(py, cy);
  const float ab_cp = cl_dot(abx, aby, cpx, cpy), ac_cp = cl_dot(acx, acy, cpx, cpy);
  if (ac_cp >= 0.0f && ab_cp <= ac_cp) return to(cx, cy);
  const float n = cl_perp(abx, aby, acx, acy);
  const float vc = cl_mul(n, cl_perp(abx, aby, apx, apy));
  if (vc < 0.0f && ab_ap >= 0.0f && ab_bp <= 0.0f) {  // edge ab
    const float t = cl_div(ab_ap, cl_dot(abx, aby, abx, aby));
    return to(cl_add(ax, cl_mul(abx, t)), cl_add(ay, cl_mul(aby, t)));
  }
  const float vb = cl_mul(-n, cl_perp(acx, acy, cpx, cpy));
  if (vb < 0.0f && ac_ap >= 0.0f && ac_cp <= 0.0f) {  // edge ac
    const float t = cl_div(ac_ap, cl_dot(acx, acy, acx, acy));
    return to(cl_add(ax, cl_mul(acx, t)), cl_add(ay, cl_mul(acy, t)));
  }
  const float bcx = cl_sub(cx, bx), bcy = cl_sub(cy, by);
  const float va = cl_mul(n, cl_perp(bcx, bcy, bpx, bpy));
  if (va < 0.0f && cl_sub(ac_bp, ab_bp) >= 0.0f && cl_sub(ab_cp, ac_cp) >= 0.0f) {  // edge bc
    const float t = cl_div(cl_dot(bcx, bcy, bpx, bpy), cl_dot(bcx, bcy, bcx, bcy));
    return to(cl_add(bx, cl_mul(bcx, t)), cl_add(by, cl_mul(bcy, t)));
  }
  d2 = 0.0f;  // on the face: in two dimensions that is inside
  return true;
}

// intersection_test(collider, robot ball at (x, z) with radius r).
GBP_DEV bool collider_hits_ball(const ColliderDev &c, const float *__restrict__ verts, float x, float z, float r) {
  const float dx = cl_sub(x, c.tx), dy = cl_sub(z, c.ty);
  // inverse rotation (re, -im) applied to (dx, dy)
  const float lx = cl_sub(cl_mul(c.re, dx), cl_mul(-c.im, dy)), ly = cl_add(cl_mul(-c.im, dx), cl_mul(c.re, dy));
  const float rr = cl_mul(r, r);
  switch (c.kind) {
    case kColliderBall: {
      const float sum = cl_add(c.radius, r);
      return cl_dot(lx, ly, lx, ly) <= cl_mul(sum, sum);
    }
    case kColliderCuboid: {
      const float m0 = cl_sub(-c.hx, lx), m1 = cl_sub(-c.hy, ly), q0 = cl_sub(lx, c.hx), q1 = cl_sub(ly, c.hy);
      const float s0 = cl_sub(m0 > 0.0f ? m0 : 0.0f, q0 > 0.0f ? q0 : 0.0f);
      const float s1 = cl_sub(m1 > 0.0f ? m1 : 0.0f, q1 > 0.0f ? q1 : 0.0f);
      if (s0 == 0.0f && s1 == 0.0f) return true;  // inside
      const float ex = cl_sub(lx, cl_add(lx, s0)), ey = cl_sub(ly, cl_add(ly, s1));
      return cl_dot(ex, ey, ex, ey) <= rr;
    }
    case kColliderTriangle: {
      float d2;
      if (cl_triangle(lx, ly, verts + 2 * c.v0, d2)) return true;
      return d2 <= rr;
    }
    default: {
      bool inside = c.nv >= 3;
      float best = 3.4e38f;
      for (int k = 0; k < c.nv; ++k) {
        const float *a = verts + 2 * (c.v0 + k), *b = verts + 2 * (c.v0 + (k + 1 == c.nv ? 0 : k + 1));
        bool left;
        const float d2 = cl_segment_dist2(lx, ly, a[0], a[1], b[0], b[1], left);
        inside = inside && left;
        best = d2 < best ? d2 : best;
      }
      return inside || best <= rr;
    }
  }
}

#ifdef __CUDACC__
// One thread per own robot: every collider, CollisionHistory::update per pair (state bit in `state`, robot-major,
// `words` 32-bit words per robot).  hits[r] = RobotEnvironmentCollisions::get(robot); totals[0] = num_collisions(),
// totals[1] = pairs colliding now.  Despawned robots are no longer in the system's query.
__global__ void k_env_collisions(int32_t n, const float *__restrict__ pos, int64_t cap, const float *__restrict__ radius,
                                 const float *__restrict__ gone, int32_t ncol, const ColliderDev *__restrict__ cols,
                                 const float *__restrict__ verts, int32_t words, uint32_t *__restrict__ state,
                                 uint32_t *__restrict__ hits, unsigned long long *totals) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || gone[r] != 0.0f) return;
  const float x = pos[r], z = pos[cap + r], rad = radius[r];
  unsigned nh = 0, now_n = 0;
  for (int w = 0; w < words; ++w) {
    const uint32_t old = state[int64_t(r) * words + w];
    uint32_t cur = 0u;
    const int hi = ncol - 32 * w < 32 ? ncol - 32 * w : 32;
    for (int b = 0; b < hi; ++b)
      if (collider_hits_ball(cols[32 * w + b], verts, x, z, rad)) cur |= 1u << b;
    nh += __popc(cur & ~old);
    now_n += __popc(cur);
    if (cur != old) state[int64_t(r) * words + w] = cur;
  }
  if (nh) {
    hits[r] += nh;
    atomicAdd(&totals[0], (unsigned long long)nh);
  }
  if (now_n) atomicAdd(&totals[1], (unsigned long long)now_n);
}

// track_positions + track_velocities (planner/tracking.rs:117-137, 226-260) for one FixedUpdate: the robots whose
// Transform changed since the systems last ran (not idle, or spawned since) tick their repeating Timer by `delta_ns`
// (bevy_time Timer::tick: elapsed += delta; finished = elapsed >= duration; elapsed %= duration) and, when it
// fires, push (x, z) into the position ring and (x - x_prev, z - z_prev) / dt with `now` and dt into the velocity
// ring (HeapRb::push_overwrite).  Rings are slot-major: entry k of robot r at [k % capacity][r].
struct TrackRings {
  int32_t capacity;
  int64_t stride;            // robots per slot plane
  uint64_t *elapsed_ns;      // [stride] Timer.stopwatch.elapsed (both trackers tick in lock step)
  uint32_t *npos, *nvel;     // [stride] pushes so far
  uint8_t *has_prev;         // [stride] VelocityTracker.previous_position.is_some()
  float *prev_xy;            // [2][stride]
  double *prev_t;            // [stride]
  float *pos_xy;             // [capacity][2][stride]
  float *vel_xy;             // [capacity][2][stride]
  double *vel_t, *vel_over;  // [capacity][stride]
};
__global__ void k_track(int32_t n, int32_t first_fresh, const float *__restrict__ pos, int64_t cap,
                        const uint8_t *__restrict__ idle, const float *__restrict__ gone, TrackRings t,
                        uint64_t duration_ns, uint64_t delta_ns, double now) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || gone[r] != 0.0f) return;
  if (idle[r] != 0 && r < first_fresh) return;  // Changed<Transform>
  uint64_t el = t.elapsed_ns[r] + delta_ns;
  const bool fired = el >= duration_ns;
  if (fired) el = duration_ns ? el % duration_ns : 0ull;
  t.elapsed_ns[r] = el;
  if (!fired) return;
  const float x = pos[r], z = pos[cap + r];
  const int64_t S = t.stride;
  {
    const uint32_t k = t.npos[r];
    const int64_t slot = int64_t(k % uint32_t(t.capacity));
    t.pos_xy[(slot * 2 + 0) * S + r] = x;
    t.pos_xy[(slot * 2 + 1) * S + r] = z;
    t.npos[r] = k + 1u;
  }
  if (t.has_prev[r]) {
    const double dt = now - t.prev_t[r];
    const float fdt = float(dt);
    const uint32_t k = t.nvel[r];
    const int64_t slot = int64_t(k % uint32_t(t.capacity));
    t.vel_xy[(slot * 2 + 0) * S + r] = __fdiv_rn(__fsub_rn(x, t.prev_xy[r]), fdt);
    t.vel_xy[(slot * 2 + 1) * S + r] = __fdiv_rn(__fsub_rn(z, t.prev_xy[S + r]), fdt);
    t.vel_t[slot * S + r] = now;
    t.vel_over[slot * S + r] = dt;
    t.nvel[r] = k + 1u;
  }
  t.prev_xy[r] = x;
  t.prev_xy[S + r] = z;
  t.prev_t[r] = now;
  t.has_prev[r] = 1;
}
#endif

}  // namespace gbp
