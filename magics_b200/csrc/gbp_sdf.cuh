// gbp_sdf.cuh — Environment -> SDF image on the device: the step before the
// Obstacle factor's pixel lookup.
//
//   k_env_raster   env_to_png::env_to_image       (crates/env_to_png/src/lib.rs:166-207) with
//                  image_to_tile_coords :249-258, image_to_tile_units :213-223, offset_modulus
//                  :241-243, is_tile_obstacle :340-479.  One thread per pixel, f32 exactly as written.
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

__global__ void k_env_raster(EnvParams e, const uint32_t *__restrict__ tiles, uint8_t *__restrict__ img) {
  const uint32_t W = uint32_t(e.ncols) * e.res, H = uint32_t(e.nrows) * e.res;
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const float fres = float(e.res);
  const uint32_t tx = uint32_t(floorf(__fdiv_rn(float(x), fres))), ty = uint32_t(floorf(__fdiv_rn(float(y), fres)));
  const float ux = __fmul_rn(__fdiv_rn(__fadd_rn(float(x), 0.5f), fres), e.tile_size);
  const float uy = __fmul_rn(__fdiv_rn(__fadd_rn(float(y), 0.5f), fres), e.tile_size);
  const float px = env_offset_modulus(ux, e.tile_size), py = env_offset_modulus(uy, e.tile_size);
  const uint32_t tile = (tx < uint32_t(e.ncols) && ty < uint32_t(e.nrows)) ? tiles[ty * e.ncols + tx] : 0x20u;
  img[size_t(y) * W + x] = env_tile_obstacle(tile, e.path_width, px, py, e.expansion) ? 0 : 255;
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
