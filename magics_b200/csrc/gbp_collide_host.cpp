// gbp_collide_host.cpp — the environment-collision predicate of k_env_collisions evaluated on the HOST, for the
// evaluation side of a run: which collider did a robot hit, and the Aabb intersection the reference records for it
// (planner/collisions.rs:402-431, :700-716; export.rs:171-206).  Collisions are rare events, the monitor's kernels
// only count them; magics_b200/collisions.py turns a counter that moved into (robot, obstacle) / (robot, robot)
// entries by asking this file about the handful of robots concerned.  The predicate is gbp_collide.cuh itself
// (collider_hits_ball, f32, every operation rounded separately through volatile temporaries) — the same source the
// device runs, so host and device agree bit for bit; no CUDA call is made here and no device is needed.
#include <algorithm>
#include <cmath>
#include <cstdint>

#define GBP_DEV inline
#include "gbp_collide.cuh"

#include "../../include/gbp_b200.h"

namespace {

bool to_dev(const gbp_collider_t &c, int32_t num_vertices, gbp::ColliderDev &d) {
  if (c.kind < 0 || c.kind > 3) return false;
  if (c.kind >= 2 && (c.num_vertices < 3 || c.first_vertex < 0 || int64_t(c.first_vertex) + c.num_vertices > num_vertices ||
                      (c.kind == 2 && c.num_vertices != 3)))
    return false;
  // Isometry2::new(translation, angle): UnitComplex::new(angle) = (cos, sin) in f32, as gbp_world_set_environment_colliders
  d = {c.kind, c.translation[0], c.translation[1], std::cos(c.angle), std::sin(c.angle), c.radius,
       c.half_extents[0], c.half_extents[1], c.first_vertex, c.num_vertices};
  return true;
}

}  // namespace

extern "C" {

int gbp_collider_hits_ball(const gbp_collider_t *collider, int32_t num_vertices, const float *vertices_xy, int32_t m,
                           const float *robots_xz, const float *radii, uint8_t *out) {
  gbp::ColliderDev d;
  if (!collider || m < 0 || (m > 0 && (!robots_xz || !radii || !out)) || !to_dev(*collider, num_vertices, d) ||
      (collider->kind >= 2 && !vertices_xy))
    return GBP_ERR_BAD_ARGUMENT;
  for (int32_t k = 0; k < m; ++k)
    out[k] = gbp::collider_hits_ball(d, vertices_xy, robots_xz[2 * k], robots_xz[2 * k + 1], radii[k]) ? 1 : 0;
  return 0;
}

// Collider::aabb = shape.compute_aabb(&isometry) (gbp_global_planner/src/lib.rs:94-98).  parry2d (third party, DESIGN
// section 2): Ball -> centre +- radius; Cuboid -> centre +- |R| half_extents (Isometry::absolute_transform_vector);
// Triangle / ConvexPolygon -> component-wise min / max of the transformed vertices R p + t (UnitComplex * Vector2 =
// (re x - im y, im x + re y), then the translation).
int gbp_collider_aabb(const gbp_collider_t *collider, int32_t num_vertices, const float *vertices_xy, float *mins_maxs) {
  gbp::ColliderDev d;
  if (!collider || !mins_maxs || !to_dev(*collider, num_vertices, d) || (collider->kind >= 2 && !vertices_xy))
    return GBP_ERR_BAD_ARGUMENT;
  using gbp::cl_add;
  using gbp::cl_mul;
  using gbp::cl_sub;
  if (d.kind == gbp::kColliderBall || d.kind == gbp::kColliderCuboid) {
    float hx = d.radius, hy = d.radius;
    if (d.kind == gbp::kColliderCuboid) {
      const float ar = std::fabs(d.re), ai = std::fabs(d.im);
      hx = cl_add(cl_mul(ar, d.hx), cl_mul(ai, d.hy));
      hy = cl_add(cl_mul(ai, d.hx), cl_mul(ar, d.hy));
    }
    mins_maxs[0] = cl_sub(d.tx, hx);
    mins_maxs[1] = cl_sub(d.ty, hy);
    mins_maxs[2] = cl_add(d.tx, hx);
    mins_maxs[3] = cl_add(d.ty, hy);
    return 0;
  }
  for (int k = 0; k < d.nv; ++k) {
    const float *p = vertices_xy + 2 * (d.v0 + k);
    const float x = cl_add(cl_sub(cl_mul(d.re, p[0]), cl_mul(d.im, p[1])), d.tx);
    const float y = cl_add(cl_add(cl_mul(d.im, p[0]), cl_mul(d.re, p[1])), d.ty);
    mins_maxs[0] = k ? std::min(mins_maxs[0], x) : x;
    mins_maxs[1] = k ? std::min(mins_maxs[1], y) : y;
    mins_maxs[2] = k ? std::max(mins_maxs[2], x) : x;
    mins_maxs[3] = k ? std::max(mins_maxs[3], y) : y;
  }
  return 0;
}

}  // extern "C"
