// gbp_store.cuh — structure-of-arrays device store of every robot's factor graph.
//
// Replaces the reference's one-FactorGraph-component-per-robot storage
// (factorgraph/factorgraph.rs:74-120: petgraph nodes + BTreeMap inboxes of boxed
// ndarray payloads) by flat planes in HBM.  A "plane" is one scalar component
// for every variable of every robot, index vi = robot*V + i, so a warp reading
// one component touches consecutive doubles.
//
// Per-variable arrays with P > 1 components are TILED (GBP_TILED, the default): 32
// consecutive variable slots form a tile, a tile holds its P component rows of 32
// doubles (256 B) back to back, element (k, vi) sits at ((vi >> 5) * P + k) * 32 + (vi & 31).
// A thread computes one base address per array and reaches every component of its
// record with an immediate offset (k * 256 B) instead of a 64-bit multiply-add per
// load; a record tile is one contiguous block (pub: 6 KiB), which is also the unit a
// bulk copy can move.  -DGBP_TILED=0 keeps whole planes of stride NV (k * NV + vi).
//
// What is stored is the minimum from which every inbox of the reference can be
// rebuilt bit-for-bit:
//   pub[p]   belief record (eta4, Lambda16, mu4) of each variable as of its last
//            internal variable iteration / change_prior — this IS the message
//            every own factor and every neighbour's InterRobot factor holds from
//            it, up to subtracting the factor's own last message (variable.rs:301-330)
//   m_*      the factor->variable messages of the own Dynamic/Obstacle/Tracking
//            factors (variable inbox entries)
//   mir      the messages a robot's variables hold from the InterRobot factors
//            owned by its neighbours ("mirror" factors), one per (edge, i)
//   mu_ext   position mean each variable last sent to external factors (per-edge override
//            mu_frozen while an edge has not received the latest one)
// pub is double buffered: a fused pass reads neighbours' records from pub[p]
// while writing its own new record to pub[1-p].
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbp {

constexpr int kEll = 8;  // edge heads per robot in the fixed-position copy (Store::ell_*)

constexpr int kRec = 24;  // eta 0..3, lambda 4..19 (row major), mu 20..23

#ifndef GBP_TILED
#define GBP_TILED 1
#endif
constexpr int kTile = 32;  // variable slots per tile; NV is kept a multiple of it

struct Store {
  int32_t N;     // robots resident on this device (local + ghost slots)
  int32_t Nloc;  // robots this device iterates (slots [0, Nloc))
  int32_t V;     // variables per robot
  int64_t NV;    // variable-slot capacity: capacity * V rounded up to a whole tile

  // Index of component k of variable slot vi in a per-variable array of P components.
  template <int P>
  __host__ __device__ __forceinline__ int64_t at(int k, int64_t vi) const {
#if GBP_TILED
    return (vi >> 5) * int64_t(P * kTile) + (vi & 31) + int64_t(k) * kTile;
#else
    return int64_t(k) * NV + vi;
#endif
  }

  // ---- per variable -------------------------------------------------------
  double *prior_eta;     // [4][NV]   VariablePrior.information_vector
  double *prior_lam;     // [NV]      diagonal of VariablePrior.precision_matrix
  double *pub[2];        // [24][NV]  belief record seen by factors (double buffered)
  uint32_t *pub_epoch[2];  // [NV]    epoch of the last write of that record (0 = never)
  double *bel_ext;       // [24][NV]  belief after the last external variable iteration
  double *mu_ext;        // [2][NV]   position mean last delivered to external factors
  double *cov;           // [16][NV]  VariableBelief.covariance_matrix
  uint8_t *valid;        // [NV]      VariableBelief.valid
  uint8_t *cov_lazy;     // [NV]      1: cov / valid are not stored — the covariance is inv4 of the variable's current
                         //           precision (bel_ext if latest[r], else pub[p]) and valid is 1; set by k_iterate_axis,
                         //           resolved by materialise_cov (gbp_iterate.cuh) and k_gather_beliefs
  double *m_dynL[2];     // [20][NV]  message from Dynamic factor i-1 (eta4, Lambda16); double buffered with pub:
  double *m_dynR[2];     // [20][NV]  message from Dynamic factor i         an internal half reads [p], writes [1-p]
  double *m_obs;         // [4][NV]   Obstacle message as (J0, J1, J2=J3, v0)
  double *m_trk;         // [3][NV]   Tracking message as (J0, J1, v0)
  double *dyn_c;         // [4][NV]   Dynamic factor i: delta_t (f32 widened) and q11, q12, q22 (gbp_math.cuh dyn_q)
  const double *dyn_tab; // [V][4]    the same four constants by variable index while every robot of the world has
                         //           the same delta_t per factor (one radius: robot.rs:1225,1232), else null
  uint32_t *trk_record;  // [NV]      Tracking.record
  int32_t *trk_timeout;  // [NV]      TrackingFactor.timeout (-1 = None)
  uint8_t *trk_seed;     // [NV]      1: the Tracking factor's inbox still holds the belief it was created with
                         //           (factorgraph.rs:310-322); 0 after FactorGraph::reset_variables emptied it
  float *trk_last;       // [2][NV]   LastMeasurement.pos (f32)
  double *trk_value;     // [NV]      LastMeasurement.value

  // ---- per robot ----------------------------------------------------------
  float *radius;          // [cap]
  float *t0;              // [cap]
  float *pos;             // [2][cap]  Transform.translation x / z
  uint8_t *antenna;       // [cap]     RadioAntenna.active
  uint8_t *idle;          // [cap]     mission.state.idle()
  uint8_t *finished;      // [cap]     FinishedPath
  float *gone;            // [cap]     1.0f: the robot's entity has been despawned (gbp_world_remove_robots): it no longer
                          //           appears in the neighbour search and is never iterated; its slot keeps its last state
  uint8_t *latest;        // [cap]     0: pub[p] holds the current belief, 1: bel_ext
  uint32_t *iter_factor;  // [cap]     FactorGraph.iteration_count.factor
  uint8_t *mode;          // [cap]     0: the robot's x and y chains are decoupled, k_iterate_axis iterates it;
                          //           1: k_iterate does; 2: k_iterate does and found it decoupled once
                          //           (invariant of mode 0 and the hand-back rule: gbp_iterate_axis.cuh)
  int32_t *gen_list;      // [cap]     robots k_iterate has to run in the current launch (filled by k_iterate_axis)
  int32_t *gen_count;     // [2]       length of gen_list, indexed by launch parity
  int32_t *gid;           // [cap]     global robot id of the slot (== slot on one GPU)
  int32_t *next_wp;       // [cap]
  int32_t *wp_off;        // [cap+1]
  float *wp_xy;           // [2*total] interleaved x,y
  int64_t cap;            // robot capacity (stride of per-robot planes)

  // ---- edges: (receiver robot B, neighbour A), CSR by B, A ascending ------
  int64_t *eoff;       // [Nloc+1]
  int32_t *nlow;       // [Nloc] number of neighbours with a lower global id (inbox order, id.rs:25-61)
  int32_t *enbr;       // [E]   slot of A
  double *e_dsafe;     // [E]   safety distance of A's factor: multiplier * radius_A
  uint64_t *e_rnum;    // [E]   robot_number of A's factor toward B at i = 1
  uint32_t *e_birth;   // [E]   epoch at which the edge was created
  uint8_t *e_frozen;   // [E]   bit 0: A's factor holds an older mean of B than mu_ext (see mu_frozen);
                       //       bit 1: CollisionState::Colliding of the pair (planner/collisions.rs:455-493)
  // The first kEll edge heads of every robot once more, at fixed positions (slot k of robot r at r * kEll + k): the hot
  // kernel can ask for them in its first wave of loads instead of waiting for eoff[r] first.  Rebuilt from the CSR
  // whenever it changes (k_ell_fill); a missing edge has ell_nbr = -1.
  int32_t *ell_nbr;    // [cap * kEll] neighbour slot
  uint32_t *ell_birth; // [cap * kEll]
  double *ell_dsafe;   // [cap * kEll]
  uint8_t *e_act;      // [E]   1: the neighbour's radio is on and it is not idle — written by k_edge_messages for the
                       //       robots k_iterate runs in the same sub-step, read there (scratch: not carried over a rebuild)
  uint32_t *coll_hits; // [cap] per robot: collisions it has been part of (RobotRobotCollisions::get)
  double *mu_frozen;   // [2][E*(V-1)] position mean of B's variable that A's factor holds while the
                       //   edge is frozen: B's belief at edge creation (robot.rs:1557-1585), or the
                       //   last mean delivered before A's antenna went off (robot.rs:1851)
  double *mir;         // [6][E*(V-1)]  eta0, eta1, lam00, lam01, lam10, lam11
  int64_t E;
  int64_t EV;          // plane stride of mir = Ecap*(V-1)

  // ---- environment ----------------------------------------------------------
  const uint8_t *sdf;  // red channel, row 0 = top
  int32_t sdf_w, sdf_h;
  double world_w, world_h, jac_delta;
  double sdf_xo, sdf_yo, sdf_xs, sdf_ys;  // ObstacleFactor::measure's offsets (world / 2) and scales (pixels / world)

  // ---- scalars widened once (robot.rs:1238-1240,1272,1318,1517-1522) ------
  double qs_dyn;   // 1/sigma_dynamics^2
  double lm_ir;    // 1/sigma_interrobot^2
  double lm_obs;   // 1/sigma_obstacle^2
  double lm_trk;   // 1/sigma_tracking^2
  double tiny_scale;  // f64::from(1e-6_f32)
  double trk_switch_padding, trk_attraction;
  uint8_t en_dyn, en_ir, en_obs, en_trk;
};

}  // namespace gbp
