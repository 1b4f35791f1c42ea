// gbp_topology.cuh — InterRobot factor creation / deletion on the device.
//
// Replaces update_robot_neighbours (planner/robot.rs:1362-1384, an all-pairs
// O(N^2) f32 distance test), delete_interrobot_factors (:1386-1439) and
// create_interrobot_factors (:1441-1586) by a sort-based spatial hash:
//   1. cell = floor(pos / (1.001 * comms radius)); key = hash(cell) (k_cell_keys)
//   2. radix sort (key, robot) (cub::DeviceRadixSort)
//   3. per robot: scan the 3x3 neighbouring cells by binary search in the sorted
//      keys, test the reference's exact f32 predicate !(R < |a-b|), count, then
//      fill; each robot's list is sorted by robot id (BTreeSet order)
//   4. diff against the previous CSR: surviving edges keep their state, new
//      edges get robot_number in the reference's creation order (robots in id
//      order -> new neighbours ascending -> i = 1..V-1, robot.rs:1500-1541);
//      the assignment kernels live in gbp_shard.cuh because the order is global
//      across shards.
// Neighbour lists hold GLOBAL robot ids; on a single GPU id == slot.
// Connectivity, ordering and robot_number are bit-exact; nothing is approximate.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "gbp_math.cuh"
#include "gbp_store.cuh"

namespace gbp {

__host__ __device__ inline uint32_t cell_hash(int32_t cx, int32_t cz) {
  uint32_t h = uint32_t(cx) * 0x9E3779B1u ^ (uint32_t(cz) * 0x85EBCA77u + 0xC2B2AE3Du);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// glam Vec3::distance on Transform.translation (x, -1.5, z) in f32, no FMA
// (robot.rs:1373-1374): sqrt((dx*dx + dy*dy) + dz*dz) with dy = 0.
__device__ __forceinline__ bool within_comms(float ax, float az, float bx, float bz, float R) {
  const float dx = __fsub_rn(ax, bx), dz = __fsub_rn(az, bz);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(0.0f, 0.0f)), __fmul_rn(dz, dz));
  const float d = __fsqrt_rn(d2);
  return !(R < d);
}

// Cell of a despawned robot (gone[r] != 0): no query ever visits it, and its own query is skipped — the
// reference's update_robot_neighbours only iterates entities that still exist (robot.rs:1362-1384).
constexpr int32_t kNoCell = INT32_MIN;

__global__ void k_cell_keys(int32_t n, const float *__restrict__ px, const float *__restrict__ pz,
                            const float *__restrict__ gone, double cell, int32_t *cx, int32_t *cz, uint32_t *keys,
                            int32_t *idx) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t x = int32_t(floor(double(px[r]) / cell)), z = int32_t(floor(double(pz[r]) / cell));
  if (gone && gone[r] != 0.0f) x = z = kNoCell;
  cx[r] = x;
  cz[r] = z;
  keys[r] = cell_hash(x, z);
  idx[r] = r;
}

// ---- sharded worlds: only robots near this shard's own robots enter the hash -------------------------
// A shard receives every robot's position (16 bytes each) but only those inside the bounding box of its own
// robots, grown by one comms radius, can be a neighbour of one of them.  k_own_bbox reduces the box (floats
// ordered as integers), k_cell_keys_near appends the candidates — own robots included — to a compact
// (key, index) list; the slots past the candidates keep the sentinel key / dummy index the host pre-filled.
__device__ __forceinline__ int32_t float_order(float f) {
  const int32_t b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__global__ void k_own_bbox(int32_t first, int32_t count, const float *__restrict__ px, const float *__restrict__ pz,
                           const float *__restrict__ gone, int32_t *box /* min x, min z, max x, max z (ordered ints) */) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int32_t r = first + t;
  if (gone && gone[r] != 0.0f) return;
  const float x = px[r], z = pz[r];
  if (!(x == x) || !(z == z)) return;  // NaN: no cell either way (the spatial hash never reproduced `!(R < NaN)`)
  atomicMin(&box[0], float_order(x));
  atomicMin(&box[1], float_order(z));
  atomicMax(&box[2], float_order(x));
  atomicMax(&box[3], float_order(z));
}
__global__ void k_cell_keys_near(int32_t n, const float *__restrict__ px, const float *__restrict__ pz,
                                 const float *__restrict__ gone, double cell, float margin,
                                 const int32_t *__restrict__ box, int32_t *cx, int32_t *cz, uint32_t *keys,
                                 int32_t *idx, int32_t cap, int32_t *count) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float fx = px[r], fz = pz[r];
  int32_t x = int32_t(floor(double(fx) / cell)), z = int32_t(floor(double(fz) / cell));
  const bool dead = gone && gone[r] != 0.0f;
  if (dead) x = z = kNoCell;
  cx[r] = x;
  cz[r] = z;
  if (dead) return;
  // compare in the ordered-integer domain of the box, with the margin applied in float (a superset is fine)
  const bool near = float_order(fx + margin) >= box[0] && float_order(fz + margin) >= box[1] &&
                    float_order(fx - margin) <= box[2] && float_order(fz - margin) <= box[3];
  if (!near) return;
  const int32_t j = atomicAdd(count, 1);
  if (j < cap) {
    keys[j] = cell_hash(x, z);
    idx[j] = r;
  }
}
__global__ void k_fill_u32(uint32_t *a, int32_t *b, int32_t n, uint32_t va, int32_t vb) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    a[t] = va;
    b[t] = vb;
  }
}

__device__ __forceinline__ int32_t lower_bound_u32(const uint32_t *a, int32_t n, uint32_t key) {
  int32_t lo = 0, hi = n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Robots within comms range of robot r (index into the position arrays): the 3x3 cells around r's cell, each
// located by binary search in the sorted keys, the reference's exact f32 predicate on every occupant.
// Up to `cap` of them are written to `dst` in discovery order; returns how many there are.
__device__ __forceinline__ int32_t search_neighbours(int32_t nall, int32_t r, const float *__restrict__ px,
                                                     const float *__restrict__ pz, const int32_t *__restrict__ cx,
                                                     const int32_t *__restrict__ cz,
                                                     const uint32_t *__restrict__ keys_sorted,
                                                     const int32_t *__restrict__ idx_sorted, float R, int32_t *dst,
                                                     int32_t cap) {
  const float ax = px[r], az = pz[r];
  const int32_t mx = cx[r], mz = cz[r];
  int32_t n = 0;
  for (int dz = -1; dz <= 1 && mx != kNoCell; ++dz)
    for (int dx = -1; dx <= 1; ++dx) {
      const int32_t qx = mx + dx, qz = mz + dz;
      const uint32_t key = cell_hash(qx, qz);
      for (int32_t j = lower_bound_u32(keys_sorted, nall, key); j < nall && keys_sorted[j] == key; ++j) {
        const int32_t o = idx_sorted[j];
        if (o == r || cx[o] != qx || cz[o] != qz) continue;
        if (!within_comms(ax, az, px[o], pz[o], R)) continue;
        if (n < cap) dst[n] = o;
        ++n;
      }
    }
  return n;
}

// Neighbour lists in two kernels with ONE search per robot: k_neighbours_find counts the neighbours of each of the
// robots [first, first+count) and parks up to kNbrPark of them per robot; after the exclusive scan of the counts
// k_neighbours_fill moves the parked ids into the CSR (searching again only for a robot with more than kNbrPark
// neighbours) and sorts each list by robot id (BTreeSet<Entity> order, robot.rs:1369-1382).
constexpr int kNbrPark = 24;
__global__ void k_neighbours_find(int32_t nall, int32_t first, int32_t count, const float *__restrict__ px,
                                  const float *__restrict__ pz, const int32_t *__restrict__ cx,
                                  const int32_t *__restrict__ cz, const uint32_t *__restrict__ keys_sorted,
                                  const int32_t *__restrict__ idx_sorted, float R, int64_t *cnt, int32_t *park) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  cnt[t] = search_neighbours(nall, first + t, px, pz, cx, cz, keys_sorted, idx_sorted, R,
                             park + int64_t(t) * kNbrPark, kNbrPark);
}
__global__ void k_neighbours_fill(int32_t nall, int32_t first, int32_t count, const float *__restrict__ px,
                                  const float *__restrict__ pz, const int32_t *__restrict__ cx,
                                  const int32_t *__restrict__ cz, const uint32_t *__restrict__ keys_sorted,
                                  const int32_t *__restrict__ idx_sorted, float R, const int64_t *__restrict__ off,
                                  const int32_t *__restrict__ park, int32_t *nbr, int64_t nbr_cap) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  if (off[count] > nbr_cap) return;  // host regrows and relaunches
  const int64_t base = off[t];
  const int32_t n = int32_t(off[t + 1] - base);
  int32_t *a = nbr + base;
  if (n <= kNbrPark) {
    for (int32_t u = 0; u < n; ++u) a[u] = park[int64_t(t) * kNbrPark + u];
  } else {
    search_neighbours(nall, first + t, px, pz, cx, cz, keys_sorted, idx_sorted, R, a, n);
  }
  // insertion sort by robot id
  for (int32_t u = 1; u < n; ++u) {
    const int32_t v = a[u];
    int32_t w = u - 1;
    while (w >= 0 && a[w] > v) {
      a[w + 1] = a[w];
      --w;
    }
    a[w + 1] = v;
  }
}

// For every own robot r in [0, n) (global id g0 + r): match its new neighbour list against
// the old one; both hold GLOBAL ids in ascending order.
//   map[e]    old edge index of new edge e, or -1 when the edge is new
//   newcnt[r] number of new edges of r, nlow[r] neighbours with a lower id
__global__ void k_edge_diff(int32_t n, int32_t g0, const int64_t *__restrict__ noff,
                            const int32_t *__restrict__ nnbr, const int64_t *__restrict__ ooff,
                            const int32_t *__restrict__ onbr, int32_t n_old, int64_t *map, int64_t *newcnt,
                            int32_t *nlow, int64_t cap) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (noff[n] > cap) {
    newcnt[r] = 0;
    return;
  }
  int64_t lo0 = 0, hi0 = 0;
  if (r < n_old && ooff) {
    lo0 = ooff[r];
    hi0 = ooff[r + 1];
  }
  int64_t fresh = 0;
  int32_t low = 0;
  for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
    const int32_t a = nnbr[e];
    if (a < g0 + r) ++low;
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (onbr[mid] < a) lo = mid + 1;
      else hi = mid;
    }
    const bool found = lo < hi0 && onbr[lo] == a;
    map[e] = found ? lo : -1;
    if (!found) ++fresh;
  }
  newcnt[r] = fresh;
  nlow[r] = low;
}

// ---- strict_reference_quirks: delete_interrobot_factors exactly as written --------------
// The reference collects (robot, lost neighbour) pairs into a HashMap keyed by robot (robot.rs:1391-1404): only the
// LAST lost neighbour of a robot — the largest id, lost neighbours are inserted in ascending order — survives, and a
// lost pair (r, a) is deleted (both directions, every factor set between the two: factorgraph.rs:380-436) iff
// a == maxlost[r] or r == maxlost[a].  Every robot drops all its lost neighbours from `robots_connected_with` either
// way, so an undeleted pair stays behind as a "zombie" factor set (e_frozen bit 2): still iterated, no longer listed
// as a connection, and joined by a second set when the two robots meet again.
constexpr uint8_t kEdgeZombie = 4;

__device__ __forceinline__ bool in_sorted(const int32_t *a, int64_t lo, int64_t hi, int32_t v) {
  const int64_t end = hi;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo < end && a[lo] == v;
}
// maxlost[r]: largest neighbour r is connected with (old non-zombie edge) that is not within range any more; -1: none.
__global__ void k_quirk_lost_max(int32_t n, const int64_t *__restrict__ ooff, const int32_t *__restrict__ onbr,
                                 const uint8_t *__restrict__ oflags, const float *__restrict__ gone,
                                 const int64_t *__restrict__ woff, const int32_t *__restrict__ wnbr, int32_t *maxlost) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t ml = -1;
  if (ooff && !(gone && gone[r] != 0.0f))  // a despawned robot is not in the query: it contributes no pair
    for (int64_t e = ooff[r]; e < ooff[r + 1]; ++e) {
      if (oflags[e] & kEdgeZombie) continue;
      const int32_t a = onbr[e];
      if (!in_sorted(wnbr, woff[r], woff[r + 1], a)) ml = a > ml ? a : ml;
    }
  maxlost[r] = ml;
}
// The new row of robot r: its old edges that are not deleted (in their old order: neighbour id, then creation)
// merged with a fresh edge for every robot within range it is not connected with.  FILL = false: row length,
// new-edge count and nlow; FILL = true: neighbour ids, map (old edge index or -1) and zombie flags.
template <bool FILL>
__global__ void k_quirk_rows(int32_t n, const int64_t *__restrict__ ooff, const int32_t *__restrict__ onbr,
                             const uint8_t *__restrict__ oflags, const float *__restrict__ gone,
                             const int64_t *__restrict__ woff, const int32_t *__restrict__ wnbr,
                             const int32_t *__restrict__ maxlost, int64_t *cnt_or_off, int32_t *nnbr, int64_t *map,
                             uint8_t *zombie, int64_t *newcnt, int32_t *nlow) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int64_t i = 0, oe = 0;
  if (ooff && !(gone && gone[r] != 0.0f)) {  // a despawned robot's graph went with it: its row empties
    i = ooff[r];
    oe = ooff[r + 1];
  }
  int64_t j = woff[r];
  const int64_t we = woff[r + 1];
  int64_t out = FILL ? cnt_or_off[r] : 0, fresh = 0;
  const int64_t out0 = out;
  int32_t low = 0;
  const int32_t mlr = maxlost[r];
  while (i < oe || j < we) {
    const int32_t ao = i < oe ? onbr[i] : INT32_MAX, aw = j < we ? wnbr[j] : INT32_MAX;
    const int32_t a = ao < aw ? ao : aw;
    const bool in_w = aw == a;
    const bool covered = mlr == a || maxlost[a] == r;
    bool had_normal = false;
    for (; i < oe && onbr[i] == a; ++i) {
      const bool zomb = (oflags[i] & kEdgeZombie) != 0;
      had_normal |= !zomb;
      if (covered) continue;  // every factor set between the two robots goes
      if (FILL) {
        nnbr[out] = a;
        map[out] = i;
        zombie[out] = (zomb || !in_w) ? 1 : 0;
      }
      ++out;
      low += a < r;
    }
    if (in_w) {
      if (!had_normal) {  // within range and not connected: create_interrobot_factors makes a new set
        if (FILL) {
          nnbr[out] = a;
          map[out] = -1;
          zombie[out] = 0;
        }
        ++out;
        ++fresh;
        low += a < r;
      }
      ++j;
    }
  }
  if (!FILL) {
    cnt_or_off[r] = out - out0;
  } else {
    newcnt[r] = fresh;
    nlow[r] = low;
  }
}

// Mirror messages (and frozen means) of surviving edges move to their new slot.
// New edges start with an Empty mirror message (add_external_edge,
// factorgraph.rs:340-353) and the receiver's current belief mean as the message
// its variable "sent" the factor (robot.rs:1557-1585, variable.rs:234-240).
__global__ void k_mirror_move(Store s, int p, int64_t total, int32_t Vm1, const int64_t *__restrict__ noff,
                              const int64_t *__restrict__ map, const double *__restrict__ omir,
                              const double *__restrict__ ofrozen, int64_t oEV, double *nmir,
                              double *nfrozen, int64_t nEV) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t e = t / Vm1, i = t - e * Vm1;
  const int64_t old = map[e];
  if (old < 0) {
    nmir[t] = empty_marker();
    // receiver robot r: noff[r] <= e < noff[r+1]
    int32_t lo = 0, hi = s.Nloc;
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (noff[mid + 1] <= e) lo = mid + 1;
      else hi = mid;
    }
    const int64_t vi = int64_t(lo) * s.V + (i + 1);
    const double *rec = s.latest[lo] ? s.bel_ext : s.pub[p];
    nfrozen[t] = rec[s.at<kRec>(20, vi)];
    nfrozen[nEV + t] = rec[s.at<kRec>(21, vi)];
    return;
  }
  const int64_t o = old * Vm1 + i;
#pragma unroll
  for (int k = 0; k < 6; ++k) nmir[k * nEV + t] = omir[k * oEV + o];
  nfrozen[t] = ofrozen[o];
  nfrozen[nEV + t] = ofrozen[oEV + o];
}

}  // namespace gbp
