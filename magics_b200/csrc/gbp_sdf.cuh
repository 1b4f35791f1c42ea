// gbp_sdf.cuh — Environment -> SDF image on the device: the step before the
// Obstacle factor's pixel lookup.
//
//   k_env_raster   env_to_png::env_to_image       (crates/env_to_png/src/lib.rs:166-207) with
//                  image_to_tile_coords :249-258, image_to_tile_units :213-223, offset_modulus
//                  :241-243, is_tile_obstacle :340-479, is_placeable_obstacle :283-336 with the shapes of
//                  crates/gbp_environment/src/lib.rs:118-435.  One thread per pixel, f32 / f64 exactly
//                  as written.
//   k_blur_rows / k_blur_cols   `image::imageops::blur` (image 0.25.1, third party): separable
//                  Gaussian, support 2*sigma, vertical pass u8 -> f32, horizontal pass f32 -> u8
//                  with clamp and round-half-away.  The unnormalised tap weights depend only on the
//                  integer tap offset; the host evaluates them once (libm expf, like the reference
//                  process would) and the kernels normalise per window in tap order.
// Everything is HBM/L2-bound byte work: one thread per output pixel, the tap loop walks a column
// (coalesced across the warp) or a row (served by L1).  Compiled with -fmad=false: no contraction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbp {

struct EnvParams {
  int32_t nrows, ncols;
  uint32_t res;  // pixels per tile
  float tile_size, path_width, expansion;
};

__device__ __forceinline__ float env_offset_modulus(float value, float modulus) {
  return __fadd_rn(__fdiv_rn(-__fsub_rn(__fmul_rn(ceilf(__fdiv_rn(value, modulus)), modulus), value), modulus), 1.0f);
}

// is_tile_obstacle (:340-479); `tile` is the Unicode code point of the box-drawing character
__device__ __forceinline__ bool env_tile_obstacle(uint32_t tile, float path_width_in, float px, float py,
                                                  float expansion) {
  const float path_width = __fsub_rn(path_width_in, expansion);
  const float ow = __fdiv_rn(__fsub_rn(1.0f, path_width), 2.0f);  // obstacle_width
  const float opw = __fadd_rn(ow, path_width);                    // obstacle_and_path_width
  const float half_lo = __fsub_rn(0.5f, __fdiv_rn(expansion, 2.0f));
  const float half_hi = __fadd_rn(0.5f, __fdiv_rn(expansion, 2.0f));
  const bool xl = px < ow, xh = px > opw, yl = py < ow, yh = py > opw;
  switch (tile) {
    case 0x2500: return yl || yh;                    // horizontal
    case 0x2502: return xl || xh;                    // vertical
    case 0x2574: return yl || yh || px > half_lo;    // left stub
    case 0x2576: return yl || yh || px < half_hi;    // right stub
    case 0x2577: return xl || xh || py < half_hi;    // down stub
    case 0x2575: return xl || xh || py > half_lo;    // up stub
    case 0x250C: return xl || yl || (xh && yh);      // corner down-right
    case 0x2510: return xh || yl || (xl && yh);      // corner down-left
    case 0x2514: return xl || yh || (xh && yl);      // corner up-right
    case 0x2518: return xh || yh || (xl && yl);      // corner up-left
    case 0x252C: return yl || (yh && (xl || xh));    // T down
    case 0x2534: return yh || (yl && (xl || xh));    // T up
    case 0x251C: return xl || (xh && (yl || yh));    // T right
    case 0x2524: return xh || (xl && (yl || yh));    // T left
    case 0x253C: return (xl || xh) && (yl || yh);    // cross
    case 0x20: return true;                          // empty tile: all obstacle
    default: return false;
  }
}

// One placeable obstacle, prepared on the host (every sin/cos of the shape's own geometry is a
// function of the YAML parameters only; the per-pixel tests below are plain arithmetic).
struct EnvShape {
  int32_t kind, tile_row, tile_col, n;  // n: vertices of a (regular) polygon
  int64_t poff;                         // first vertex in the vertex array (pairs of doubles)
  float tx, ty;                         // translation as f32 (Vec2::from(RelativePoint))
  float qs, qw;                         // sin / cos of half the z rotation (glam Quat::from_rotation_z)
  float r2;                             // circle: (radius + expansion)^2 as f32
  float tri[6];                         // triangle: a, b, c
  double hw, hh;                        // rectangle: half_width, half_height (width / 4, height / 4)
};
enum { kShapeCircle = 0, kShapeTriangle = 1, kShapeRegularPolygon = 2, kShapePolygon = 3, kShapeRectangle = 4 };

// Triangle::inside helper `sign` (gbp_environment/src/lib.rs:229-231), f32
__device__ __forceinline__ float env_sign(float p1x, float p1y, float p2x, float p2y, float p3x, float p3y) {
  return __fsub_rn(__fmul_rn(__fsub_rn(p1x, p3x), __fsub_rn(p2y, p3y)), __fmul_rn(__fsub_rn(p2x, p3x), __fsub_rn(p1y, p3y)));
}

// is_placeable_obstacle (env_to_png/src/lib.rs:283-336) for the pixel at fraction (px, py) of tile (tx, ty)
__device__ bool env_placeable_obstacle(const EnvShape *__restrict__ shapes, int nshapes, const double *__restrict__ verts,
                                       uint32_t tile_x, uint32_t tile_y, float px, float py) {
  for (int k = 0; k < nshapes; ++k) {
    const EnvShape &o = shapes[k];
    if (uint32_t(o.tile_col) != tile_x || uint32_t(o.tile_row) != tile_y) continue;
    // translated = percentage - translation; rotated = Quat::from_rotation_z(angle) * (x, y, 0)
    const float x = __fsub_rn(px, o.tx), y = __fsub_rn(py, o.ty);
    const float c = __fsub_rn(__fmul_rn(o.qw, o.qw), __fmul_rn(o.qs, o.qs));  // w*w - b.b
    const float w2 = __fmul_rn(o.qw, 2.0f);
    const float rx = __fadd_rn(__fadd_rn(__fmul_rn(x, c), 0.0f), __fmul_rn(-__fmul_rn(o.qs, y), w2));
    const float ry = __fadd_rn(__fadd_rn(__fmul_rn(y, c), 0.0f), __fmul_rn(__fmul_rn(o.qs, x), w2));
    bool inside = false;
    switch (o.kind) {
      case kShapeCircle:  // :138-141
        inside = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)) <= o.r2;
        break;
      case kShapeTriangle: {  // :214-226
        const float d1 = env_sign(rx, ry, o.tri[0], o.tri[1], o.tri[2], o.tri[3]);
        const float d2 = env_sign(rx, ry, o.tri[2], o.tri[3], o.tri[4], o.tri[5]);
        const float d3 = env_sign(rx, ry, o.tri[4], o.tri[5], o.tri[0], o.tri[1]);
        const bool has_neg = d1 < 0.0f || d2 < 0.0f || d3 < 0.0f;
        const bool has_pos = d1 > 0.0f || d2 > 0.0f || d3 > 0.0f;
        inside = !(has_neg && has_pos);
        break;
      }
      case kShapeRegularPolygon: {  // :298-313 (the point is doubled)
        const double X = double(rx) * 2.0, Y = double(ry) * 2.0;
        int j = o.n - 1;
        for (int i = 0; i < o.n; ++i) {
          const double xi = verts[2 * (o.poff + i)], yi = verts[2 * (o.poff + i) + 1];
          const double xj = verts[2 * (o.poff + j)], yj = verts[2 * (o.poff + j) + 1];
          if ((yi < Y && yj >= Y) || (yj < Y && yi >= Y)) {
            if (xi + (Y - yi) / (yj - yi) * (xj - xi) < X) inside = !inside;
          }
          j = i;
        }
        break;
      }
      case kShapePolygon: {  // is_point_in_polygon :412-428
        const double X = double(rx), Y = double(ry);
        int j = o.n - 1;
        for (int i = 0; i < o.n; ++i) {
          const double ix = verts[2 * (o.poff + i)], iy = verts[2 * (o.poff + i) + 1];
          const double jx = verts[2 * (o.poff + j)], jy = verts[2 * (o.poff + j) + 1];
          if ((iy > Y) != (jy > Y) && X < (jx - ix) * (Y - iy) / (jy - iy) + ix) inside = !inside;
          j = i;
        }
        break;
      }
      case kShapeRectangle: {  // :347-359 (x against the height, y against the width, as written)
        const double X = double(rx), Y = double(ry);
        inside = X >= -o.hh && X <= o.hh && Y >= -o.hw && Y <= o.hw;
        break;
      }
      default: break;
    }
    if (inside) return true;
  }
  return false;
}

__global__ void k_env_raster(EnvParams e, const uint32_t *__restrict__ tiles, const EnvShape *__restrict__ shapes,
                             int nshapes, const double *__restrict__ verts, uint8_t *__restrict__ img) {
  const uint32_t W = uint32_t(e.ncols) * e.res, H = uint32_t(e.nrows) * e.res;
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const float fres = float(e.res);
  const uint32_t tx = uint32_t(floorf(__fdiv_rn(float(x), fres))), ty = uint32_t(floorf(__fdiv_rn(float(y), fres)));
  const float ux = __fmul_rn(__fdiv_rn(__fadd_rn(float(x), 0.5f), fres), e.tile_size);
  const float uy = __fmul_rn(__fdiv_rn(__fadd_rn(float(y), 0.5f), fres), e.tile_size);
  const float px = env_offset_modulus(ux, e.tile_size), py = env_offset_modulus(uy, e.tile_size);
  const uint32_t tile = (tx < uint32_t(e.ncols) && ty < uint32_t(e.nrows)) ? tiles[ty * e.ncols + tx] : 0x20u;
  const bool obstacle = env_tile_obstacle(tile, e.path_width, px, py, e.expansion) ||
                        env_placeable_obstacle(shapes, nshapes, verts, tx, ty, px, py);
  img[size_t(y) * W + x] = obstacle ? 0 : 255;
}

// Window of output index o along an axis of length n (image 0.25.1 sample.rs): [left, right).
__device__ __forceinline__ void blur_window(uint32_t o, uint32_t n, float support, int64_t &left, int64_t &right) {
  const float c = __fadd_rn(float(o), 0.5f);
  int64_t l = int64_t(floorf(__fsub_rn(c, support)));
  l = l < 0 ? 0 : (l > int64_t(n) - 1 ? int64_t(n) - 1 : l);
  int64_t r = int64_t(ceilf(__fadd_rn(c, support)));
  r = r < l + 1 ? l + 1 : (r > int64_t(n) ? int64_t(n) : r);
  left = l;
  right = r;
}

// wtab[d + D] = gaussian(d, sigma) for integer tap offsets d in [-D, D]
// vertical_sample: u8 image -> f32 image
__global__ void k_blur_rows(const uint8_t *__restrict__ src, float *__restrict__ dst, uint32_t W, uint32_t H,
                            float support, const float *__restrict__ wtab, int D) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (x >= W || oy >= H) return;
  int64_t left, right;
  blur_window(oy, H, support, left, right);
  float sum = 0.0f;
  for (int64_t i = left; i < right; ++i) sum = __fadd_rn(sum, wtab[int(i - int64_t(oy)) + D]);
  float t = 0.0f;
  for (int64_t i = left; i < right; ++i) {
    const float w = __fdiv_rn(wtab[int(i - int64_t(oy)) + D], sum);
    t = __fadd_rn(t, __fmul_rn(float(src[size_t(i) * W + x]), w));
  }
  dst[size_t(oy) * W + x] = t;
}

// horizontal_sample: f32 image -> u8 image, clamp to [0, 255], round half away from zero
__global__ void k_blur_cols(const float *__restrict__ src, uint8_t *__restrict__ dst, uint32_t W, uint32_t H,
                            float support, const float *__restrict__ wtab, int D) {
  const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (ox >= W || y >= H) return;
  int64_t left, right;
  blur_window(ox, W, support, left, right);
  float sum = 0.0f;
  for (int64_t i = left; i < right; ++i) sum = __fadd_rn(sum, wtab[int(i - int64_t(ox)) + D]);
  float t = 0.0f;
  for (int64_t i = left; i < right; ++i) {
    const float w = __fdiv_rn(wtab[int(i - int64_t(ox)) + D], sum);
    t = __fadd_rn(t, __fmul_rn(src[size_t(y) * W + size_t(i)], w));
  }
  const float cl = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
  dst[size_t(y) * W + ox] = uint8_t(roundf(cl));
}

__global__ void k_gray_to_rgb(const uint8_t *__restrict__ g, uint8_t *__restrict__ rgb, size_t n) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint8_t v = g[t];
  rgb[3 * t] = v;
  rgb[3 * t + 1] = v;
  rgb[3 * t + 2] = v;
}

}  // namespace gbp
