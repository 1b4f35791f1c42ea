// gbp_shard.cuh — device side of the multi-GPU sharding.
//
// Robots are owned by contiguous global-id ranges in rank order (`gfirst`), which is
// the reference's robot order (bevy Entity order, id.rs:25-61).  A shard iterates its
// own robots (slots [0, Nloc)) and keeps GHOST slots [Nloc, N) for robots of other
// shards that are within comms range of one of its own.  Because the engine pulls
// (gbp_iterate.cuh), the only thing a shard ever needs from a ghost A is what A's
// InterRobot factors hold from A's own variables: A's published belief record, its
// epoch and A's antenna/idle bits.  That record is the halo; it is refreshed once
// per sub-step (after every internal variable iteration / prior change), which
// carries exactly the traffic of the reference's two delivery loops
// (robot.rs:1814-1831 factor->variable, :1843-1858 variable->factor) and of the
// horizon `change_prior` messages (robot.rs:2272-2282).
//
// The kernels here
//   * mark ghosts / build per-peer send lists from the neighbour lists (k_cross_mark,
//     k_ghost_fill, k_sendlist_fill, k_edge_slots),
//   * move the robot_number of cross-shard InterRobot factors from the shard that owns
//     the factor to the shard that evaluates it (k_cross_pack / k_edge_pull), so that
//     RobotNumberGenerator order (robot.rs:121-144, :1527) stays globally exact,
//   * pack / unpack the per-sub-step halo (k_halo_pack / k_halo_unpack).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "gbp_store.cuh"

namespace gbp {

constexpr int kMaxShards = 16;
constexpr unsigned long long kRelBit = 1ULL << 63;  // cross value is relative to the sender's base

struct ShardInfo {
  int32_t ws, rank;
  int32_t gfirst[kMaxShards + 1];  // shard q owns global ids [gfirst[q], gfirst[q+1])
};

__host__ __device__ inline int owner_of(const ShardInfo &sh, int32_t gid) {
  int q = 0;
  while (q + 1 < sh.ws && gid >= sh.gfirst[q + 1]) ++q;
  return q;
}

// Start offsets (in records) of the per-peer blocks of a halo / cross buffer.
struct PeerOffsets {
  int64_t start[kMaxShards + 1];
};

__device__ __forceinline__ int block_of(const PeerOffsets &po, int ws, int64_t j) {
  int q = 0;
  while (q + 1 < ws && j >= po.start[q + 1]) ++q;
  return q;
}

// Per own robot r: count neighbours owned by each other shard, flag r for that
// shard's send list, flag the neighbour as a ghost.  Neighbour lists hold global ids.
__global__ void k_cross_mark(ShardInfo sh, int32_t nloc, const int64_t *__restrict__ noff,
                             const int32_t *__restrict__ ngid, int64_t cap, int32_t *gflag,
                             int32_t *sflag, int64_t *ccnt) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  int32_t cnt[kMaxShards];
#pragma unroll
  for (int q = 0; q < kMaxShards; ++q) cnt[q] = 0;
  if (noff[nloc] <= cap) {
    const int32_t g0 = sh.gfirst[sh.rank], g1 = sh.gfirst[sh.rank + 1];
    for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
      const int32_t a = ngid[e];
      if (a >= g0 && a < g1) continue;
      cnt[owner_of(sh, a)] += 1;
      gflag[a] = 1;
    }
  }
  for (int q = 0; q < sh.ws; ++q) {
    sflag[int64_t(q) * nloc + r] = cnt[q] > 0 ? 1 : 0;
    ccnt[int64_t(q) * nloc + r] = cnt[q];
  }
}

// out[0] = E1, out[1] = new directed pairs of this shard, out[2] = ghosts, out[3] = error flag,
// out[4 + q] = ghost block start of shard q, out[4 + (ws+1) + q] = send block start,
// out[4 + 2(ws+1) + q] = cross-edge block start   (q = 0..ws)
// out[4 + 3(ws+1)] = robots that passed the candidate filter of the neighbour search (k_cell_keys_near)
__global__ void k_shard_result(ShardInfo sh, int32_t nloc, const int64_t *noff, const int64_t *newoff,
                               const int32_t *gslot, const int32_t *soff, const int64_t *coff,
                               const int32_t *err, const int32_t *ncand, int64_t *out) {
  const int q = threadIdx.x;
  if (q == 0) {
    out[4 + 3 * (kMaxShards + 1)] = ncand ? *ncand : 0;
    out[0] = noff[nloc];
    out[1] = newoff[nloc];
    out[2] = gslot ? gslot[sh.gfirst[sh.ws]] : 0;
    out[3] = *err;  // set by k_edge_pull of an earlier pass when a cross-shard lookup failed
  }
  if (q <= sh.ws && gslot) {
    out[4 + q] = gslot[sh.gfirst[q]];
    out[4 + (sh.ws + 1) + q] = soff[int64_t(q) * nloc];
    out[4 + 2 * (sh.ws + 1) + q] = coff[int64_t(q) * nloc];
  }
}

// Ghost slot of every flagged global id: gid and radius of the slot.
__global__ void k_ghost_fill(int32_t ntot, int32_t nloc, const int32_t *__restrict__ gflag,
                             const int32_t *__restrict__ gslot, const float *__restrict__ gradius,
                             int32_t *slot_gid, float *slot_radius) {
  const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ntot || !gflag[g]) return;
  const int32_t slot = nloc + gslot[g];
  slot_gid[slot] = g;
  slot_radius[slot] = gradius[g];
}

__global__ void k_sendlist_fill(int32_t ws, int32_t nloc, const int32_t *__restrict__ sflag,
                                const int32_t *__restrict__ soff, int32_t *sendlist) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= int64_t(ws) * nloc) return;
  if (sflag[t]) sendlist[soff[t]] = int32_t(t % nloc);
}

// Union of the per-peer send lists: flags (one byte per own robot, set through its word) and a list without
// duplicates (any order) — the robots whose internal half runs first so that their records can travel while
// the other robots are iterated (group_launch).
__global__ void k_border_list(int64_t nsend, const int32_t *__restrict__ sendlist, uint32_t *words, int32_t *list,
                              int32_t *count) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= nsend) return;
  const int32_t r = sendlist[j];
  const uint32_t bit = 1u << (8 * (r & 3));
  if (!(atomicOr(&words[r >> 2], bit) & bit)) list[atomicAdd(count, 1)] = r;
}

// Neighbour global id -> slot (own: gid - g0; ghost: Nloc + rank among ghosts).
__global__ void k_edge_slots(int64_t E, int32_t g0, int32_t nloc, const int32_t *__restrict__ ngid,
                             const int32_t *__restrict__ gslot, int32_t *enbr) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t a = ngid[e];
  enbr[e] = (a >= g0 && a < g0 + nloc) ? a - g0 : nloc + gslot[a];
}

// Owner side of the cross-shard robot_number exchange.  For every edge (r -> x) with x
// on shard q the pair key (gid r, gid x) and the number of r's factor toward x: the
// absolute value for a surviving factor, kRelBit | (rank among this shard's new
// directed pairs) for a new one (the receiver adds the shard's base, which depends
// on every lower shard's count and travels in the same exchange).
__global__ void k_cross_pack(ShardInfo sh, int32_t nloc, const int64_t *__restrict__ noff,
                             const int32_t *__restrict__ ngid, const int64_t *__restrict__ map,
                             const int64_t *__restrict__ newoff, const uint64_t *__restrict__ o_own,
                             const int64_t *__restrict__ coff, uint64_t *keys, uint64_t *vals) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  const int32_t g0 = sh.gfirst[sh.rank], g1 = sh.gfirst[sh.rank + 1];
  int32_t cnt[kMaxShards];
#pragma unroll
  for (int q = 0; q < kMaxShards; ++q) cnt[q] = 0;
  int64_t fresh = 0;
  for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
    const int32_t a = ngid[e];
    const int64_t old = map[e];
    const uint64_t v = old >= 0 ? o_own[old] : (kRelBit | uint64_t(newoff[r] + fresh));
    if (old < 0) ++fresh;
    if (a >= g0 && a < g1) continue;
    const int q = owner_of(sh, a);
    const int64_t at = coff[int64_t(q) * nloc + r] + cnt[q];
    cnt[q] += 1;
    keys[at] = (uint64_t(uint32_t(g0 + r)) << 32) | uint64_t(uint32_t(a));
    vals[at] = v;
  }
}

// Per new-CSR edge (r <- a), owner view: carry over or initialise the edge scalars and
// the number of r's OWN factor toward a (robot.rs:1500-1541: robots in id order -> new
// neighbours ascending -> i = 1..V-1; `base` = new directed pairs of all lower shards).
__global__ void k_edge_assign_own(int32_t nloc, int32_t V, const int64_t *__restrict__ noff,
                                  const int32_t *__restrict__ ngid, const int64_t *__restrict__ map,
                                  const int64_t *__restrict__ newoff, const float *__restrict__ gradius,
                                  double safety_mult, uint64_t counter0, int64_t base, uint32_t epoch,
                                  const uint64_t *__restrict__ o_own, const uint64_t *__restrict__ o_rnum,
                                  const uint32_t *__restrict__ o_birth, const uint8_t *__restrict__ o_frozen,
                                  uint64_t *e_own, double *e_dsafe, uint64_t *e_rnum, uint32_t *e_birth,
                                  uint8_t *e_frozen, const uint8_t *__restrict__ zombie = nullptr) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  int64_t fresh = 0;
  for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
    const int64_t old = map[e];
    e_dsafe[e] = safety_mult * double(gradius[ngid[e]]);
    if (old >= 0) {
      e_own[e] = o_own[old];
      e_rnum[e] = o_rnum[old];
      e_birth[e] = o_birth[old];
      // strict_reference_quirks: bit 2 marks a factor set its robots no longer list as a connection
      e_frozen[e] = uint8_t((o_frozen[old] & ~4) | ((zombie && zombie[e]) ? 4 : 0));
    } else {
      e_own[e] = counter0 + uint64_t(V - 1) * uint64_t(base + newoff[r] + fresh);
      ++fresh;
      e_birth[e] = epoch;
      e_frozen[e] = 1;  // the factor holds the receiver's belief at creation time
    }
  }
}

struct ShardBases {
  int64_t base[kMaxShards];  // new directed pairs created this tick by all lower shards
};

// Evaluator side: for every NEW edge (r <- a) fetch the number of a's factor toward r,
// from a's own row when a is local, else from the keys/values its owner sent.
__global__ void k_edge_pull(ShardInfo sh, int32_t nloc, int32_t V, const int64_t *__restrict__ noff,
                            const int32_t *__restrict__ ngid, const int64_t *__restrict__ map,
                            const uint64_t *__restrict__ e_own, uint64_t counter0, ShardBases bases,
                            PeerOffsets cross, const uint64_t *__restrict__ rkeys,
                            const uint64_t *__restrict__ rvals, uint64_t *e_rnum, int32_t *err) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  const int32_t g0 = sh.gfirst[sh.rank], g1 = sh.gfirst[sh.rank + 1];
  for (int64_t e = noff[r]; e < noff[r + 1]; ++e) {
    if (map[e] >= 0) continue;
    const int32_t a = ngid[e];
    if (a >= g0 && a < g1) {
      const int32_t al = a - g0;
      int64_t lo = noff[al], hi = noff[al + 1];
      const int64_t end = hi;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ngid[mid] < g0 + r) lo = mid + 1;
        else hi = mid;
      }
      // strict_reference_quirks: older sets of the same pair come first; the new edge pairs with the new one
      while (lo < end && ngid[lo] == g0 + r && map[lo] >= 0) ++lo;
      if (lo < end && ngid[lo] == g0 + r) e_rnum[e] = e_own[lo];
      else atomicExch(err, 1);
    } else {
      const int q = owner_of(sh, a);
      const uint64_t key = (uint64_t(uint32_t(a)) << 32) | uint64_t(uint32_t(g0 + r));
      int64_t lo = cross.start[q], hi = cross.start[q + 1];
      const int64_t end = hi;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rkeys[mid] < key) lo = mid + 1;
        else hi = mid;
      }
      if (lo < end && rkeys[lo] == key) {
        const uint64_t v = rvals[lo];
        e_rnum[e] = (v & kRelBit) ? counter0 + uint64_t(V - 1) * uint64_t(bases.base[q] + int64_t(v & ~kRelBit)) : v;
      } else {
        atomicExch(err, 2);
      }
    }
  }
}

// ---- per-sub-step halo ----------------------------------------------------------
// One record per (robot, variable i >= 1): eta4, Lambda16, position mean (22 doubles)
// + the record's epoch; per robot two more doubles: antenna/idle bits, and the f32 Transform
// (x, z) packed into one double (the collision monitor tests ghosts at their current position).
// Inside a peer block the layout is plane-major so that consecutive threads touch consecutive doubles.
constexpr int kHaloPlanes = 23;
__host__ __device__ inline int64_t halo_doubles_per_robot(int V) { return int64_t(kHaloPlanes) * (V - 1) + 2; }

__global__ void k_halo_pack(Store s, int p, int ws, PeerOffsets po, const int32_t *__restrict__ sendlist,
                            double *__restrict__ buf) {
  const int Vm1 = s.V - 1;
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= po.start[ws] * Vm1) return;
  const int64_t j = t / Vm1;
  const int i = 1 + int(t - j * Vm1);
  const int q = block_of(po, ws, j);
  const int64_t cnt = po.start[q + 1] - po.start[q], jj = j - po.start[q];
  const int64_t ps = cnt * Vm1, at = jj * Vm1 + (i - 1);
  double *blk = buf + po.start[q] * halo_doubles_per_robot(s.V);
  const int32_t r = sendlist[j];
  const int64_t vi = int64_t(r) * s.V + i;
  const double *rec = s.pub[p];
#pragma unroll
  for (int k = 0; k < 22; ++k) blk[k * ps + at] = rec[s.at<kRec>(k, vi)];
  blk[22 * ps + at] = double(s.pub_epoch[p][vi]);
  if (i == 1) {
    blk[kHaloPlanes * ps + jj] = double(int(s.antenna[r] != 0) | (int(s.idle[r] != 0) << 1));
    const unsigned long long xz = (unsigned long long)__float_as_uint(s.pos[r]) |
                                  ((unsigned long long)__float_as_uint(s.pos[s.cap + r]) << 32);
    blk[kHaloPlanes * ps + cnt + jj] = __longlong_as_double((long long)xz);
  }
}

__global__ void k_halo_unpack(Store s, int p, int ws, PeerOffsets po, const double *__restrict__ buf) {
  const int Vm1 = s.V - 1;
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= po.start[ws] * Vm1) return;
  const int64_t j = t / Vm1;
  const int i = 1 + int(t - j * Vm1);
  const int q = block_of(po, ws, j);
  const int64_t cnt = po.start[q + 1] - po.start[q], jj = j - po.start[q];
  const int64_t ps = cnt * Vm1, at = jj * Vm1 + (i - 1);
  const double *blk = buf + po.start[q] * halo_doubles_per_robot(s.V);
  const int64_t slot = int64_t(s.Nloc) + j;
  const int64_t vi = slot * s.V + i;
  double *rec = s.pub[p];
#pragma unroll
  for (int k = 0; k < 22; ++k) rec[s.at<kRec>(k, vi)] = blk[k * ps + at];
  s.pub_epoch[p][vi] = uint32_t(blk[22 * ps + at]);
  if (i == 1) {
    const int bits = int(blk[kHaloPlanes * ps + jj]);
    s.antenna[slot] = uint8_t(bits & 1);
    s.idle[slot] = uint8_t((bits >> 1) & 1);
    const unsigned long long xz = (unsigned long long)__double_as_longlong(blk[kHaloPlanes * ps + cnt + jj]);
    s.pos[slot] = __uint_as_float(unsigned(xz & 0xffffffffull));
    s.pos[s.cap + slot] = __uint_as_float(unsigned(xz >> 32));
  }
}

// update_robot_robot_collisions (planner/collisions.rs:72-143) restricted to connected pairs (every
// colliding pair is connected as long as radius_a + radius_b <= comms radius, which the host checks):
// parry2d BoundingSphere::intersects in f32 — |c_b - c_a|^2 <= (r_a + r_b)^2 — and the Free/Colliding
// state machine of CollisionHistory::update (:472-488).  Every directed edge carries the pair's state
// (both directions evolve identically, so a pair split over two shards needs no message); a Hit
// increments the own robot's counter, and the global counter once per pair (lower id's side).
__global__ void k_robot_collisions(Store s, const int32_t *__restrict__ egid, int32_t g0,
                                   unsigned long long *totals /* [0] hits, [1] colliding pairs now */) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= s.Nloc) return;
  const float ax = s.pos[r], az = s.pos[s.cap + r], ar = s.radius[r];
  unsigned hits = 0, pair_hits = 0, pair_now = 0;
  for (int64_t e = s.eoff[r]; e < s.eoff[r + 1]; ++e) {
    const int32_t a = s.enbr[e];
    if (s.e_frozen[e] & 4) continue;  // not a connection (strict_reference_quirks); a touching pair has a live edge too
    const float dx = __fsub_rn(s.pos[a], ax), dz = __fsub_rn(s.pos[s.cap + a], az);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz));
    const float sr = __fadd_rn(ar, s.radius[a]);
    const bool now = d2 <= __fmul_rn(sr, sr);
    const uint8_t cur = s.e_frozen[e];
    const bool was = (cur & 2) != 0;
    if (now != was) s.e_frozen[e] = uint8_t((cur & 1) | (now ? 2 : 0));
    const bool lower = g0 + r < egid[e];
    if (now && !was) {  // CollisionStatus::Hit
      hits += 1;
      if (lower) pair_hits += 1;
    }
    if (now && lower) pair_now += 1;
  }
  if (hits) s.coll_hits[r] += hits;
  if (pair_hits) atomicAdd(&totals[0], (unsigned long long)pair_hits);
  if (pair_now) atomicAdd(&totals[1], (unsigned long long)pair_now);
}

}  // namespace gbp
