/* gbp_b200.h — C ABI of the B200-native Gaussian Belief Propagation engine.
 *
 * This is the drop-in boundary for the hot path of AU-Master-Thesis/magics
 * (per-tick GBP iteration over every robot's factor graph + InterRobot factor
 * creation + Obstacle SDF lookup).  The reference has no FFI of its own: the seam
 * is the public method set of its `FactorGraph` component and the FixedUpdate
 * systems of `RobotPlugin`.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference checkout,
 * crates/magics/src/...).  A per-robot API would serialise the GPU, so each
 * reference method becomes ONE world-level call that does the same thing for
 * every robot at once; the Rust `FactorGraph` facade forwards to these
 * (INTEGRATION.md shows the -sys binding).
 *
 * Conventions
 *   - plain C types only, no torch / CUDA types in any signature;
 *   - all pointers are HOST pointers borrowed for the duration of the call;
 *   - every function returns 0 on success, <0 on error (gbp_last_error());
 *     numerical degeneracy is never an error, it stays in-band (validity flags)
 *     exactly as in the reference (variable.rs:276-297, marginalise_factor_distance.rs:79-81);
 *   - a world handle is not re-entrant: one host thread per handle;
 *   - robots are identified by their index in insertion order; this index is
 *     the total order the reference derives from bevy `Entity` (id.rs:25-61) and
 *     is what fixes inbox order, InterRobot slot order and robot_number order.
 */
#ifndef GBP_B200_H
#define GBP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBP_DOFS 4 /* factorgraph/mod.rs:21 */

/* gbp_config::GbpIterationScheduleKind (gbp_config/src/lib.rs:360-376) */
enum gbp_schedule_kind {
  GBP_SCHEDULE_CENTERED = 0,
  GBP_SCHEDULE_INTERLEAVE_EVENLY = 1,
  GBP_SCHEDULE_SOON_AS_POSSIBLE = 2,
  GBP_SCHEDULE_LATE_AS_POSSIBLE = 3,
  GBP_SCHEDULE_HALF_BEGINNING_HALF_END = 4
};

enum gbp_status {
  GBP_OK = 0,
  GBP_ERR_BAD_HANDLE = -1,
  GBP_ERR_BAD_ARGUMENT = -2,
  GBP_ERR_CUDA = -3,
  GBP_ERR_NCCL = -4,
  GBP_ERR_STATE = -5
};

/* The scalars of gbp_config::Config that cross the boundary.
 * f32 fields are f32 in the reference and are widened with f64::from exactly
 * where the reference widens them (robot.rs:1238-1240,1272,1318,1517-1522). */
typedef struct gbp_config {
  int32_t num_variables;           /* V = variable_timesteps.len() (spawner.rs:559-569) */
  float sigma_factor_dynamics;     /* GbpSection (gbp_config/src/lib.rs:544-594) */
  float sigma_factor_interrobot;
  float sigma_factor_obstacle;
  float sigma_factor_tracking;
  float safety_distance_multiplier; /* robot.inter-robot-safety-distance-multiplier */
  float comms_radius;               /* robot.communication.radius (robot.rs:1372) */
  float target_speed;               /* robot.target-speed (robot.rs:2207, 1225) */
  float delta_t;                    /* 1/simulation.hz as f32: Time::delta_seconds (robot.rs:2205,2309) */
  float tracking_switch_padding;      /* TrackingSection (lib.rs:502-537) */
  float tracking_attraction_distance;
  uint8_t enable_dynamic;           /* FactorsEnabledSection (lib.rs:454-494) */
  uint8_t enable_interrobot;
  uint8_t enable_obstacle;
  uint8_t enable_tracking;
  int32_t schedule_kind;            /* enum gbp_schedule_kind */
  int32_t iterations_internal;      /* GbpIterationSchedule.internal (cast `as u8`, robot.rs:1782) */
  int32_t iterations_external;
  double world_width;               /* obstacle::WorldSize (robot.rs:1258-1263) */
  double world_height;
  /* 0 (default): every InterRobot factor between two robots that have left each other's comms range is deleted.
   * 1: delete_interrobot_factors exactly as written (robot.rs:1386-1439): the (robot, lost
   *    neighbour) pairs go through a HashMap keyed by robot, so a robot that loses several neighbours in one tick
   *    keeps only its LAST (largest id) pair; a lost pair (a, b) is deleted iff b is a's last lost neighbour or a is
   *    b's.  The others stay as factor sets nobody lists in `robots_connected_with` any more; they keep being
   *    iterated, and when the pair meets again a second set is created next to them.  Single-GPU worlds only. */
  int32_t strict_reference_quirks;
} gbp_config_t;

typedef struct gbp_world gbp_world_t; /* opaque */

/* Thread-local description of the last error on this thread. */
const char *gbp_last_error(void);

/* ---- gbp_schedule crate ------------------------------------------------ */
/* GbpSchedule::schedule(GbpScheduleParams{internal,external})
 * (gbp_schedule/src/schedules/mod.rs:60-81; dispatch gbp_config/src/lib.rs:378-400).
 * Writes max(internal,external) entries into out_internal/out_external (0/1)
 * and returns that count (<= 255), or <0 on error. */
int gbp_schedule(int32_t kind, uint8_t internal, uint8_t external,
                 uint8_t *out_internal, uint8_t *out_external);

/* utils::get_variable_timesteps (utils.rs:35-75).  Returns the count written
 * (capacity must be >= lookahead_multiple*(n+1)); <0 on error. */
int gbp_variable_timesteps(uint32_t lookahead_horizon, uint32_t lookahead_multiple,
                           uint32_t *out, int32_t capacity);

/* ---- world lifetime ---------------------------------------------------- */
/* Replaces scenario load + reset_robot_number_generator (robot.rs:121-144).
 * device: CUDA device ordinal. */
gbp_world_t *gbp_world_create(const gbp_config_t *cfg, int32_t device);
void gbp_world_destroy(gbp_world_t *w);

/* Replaces the `Sdf(SdfImage)` resource cloned into every ObstacleFactor
 * (simulation_loader.rs:52, robot.rs:1274): RGB8, row 0 = top, uploaded once. */
int gbp_world_set_sdf(gbp_world_t *w, const uint8_t *rgb8, int32_t width, int32_t height);

/* `gbp_environment::Obstacle` (crates/gbp_environment/src/lib.rs:532-572) with its `PlaceableShape`
 * (:437-443: Circle :118-142, Triangle :154-227, RegularPolygon :233-314, Rectangle :316-361, Polygon
 * :363-435).  Angles in radians as the YAML holds them (`Angle(f64)`). */
enum {
  GBP_SHAPE_CIRCLE = 0,
  GBP_SHAPE_TRIANGLE = 1,
  GBP_SHAPE_REGULAR_POLYGON = 2,
  GBP_SHAPE_POLYGON = 3,
  GBP_SHAPE_RECTANGLE = 4
};
typedef struct gbp_obstacle {
  int32_t kind;               /* GBP_SHAPE_* */
  int32_t tile_row, tile_col; /* tile-coordinates */
  double rotation;            /* rotation.as_radians() */
  double tx, ty;              /* translation (RelativePoint) */
  double radius;              /* circle, triangle (inscribed circle), regular polygon */
  double angle_a, angle_b;    /* triangle: angles.A, angles.B */
  int32_t sides;              /* regular polygon */
  int32_t n_points;           /* polygon */
  int64_t point_offset;       /* polygon: first point in gbp_environment_t::polygon_points */
  double width, height;       /* rectangle */
} gbp_obstacle_t;

/* `gbp_environment::Environment` as far as SDF generation reads it
 * (crates/gbp_environment/src/lib.rs:40-75,940-971; settings.sdf: resolution, expansion, blur). */
typedef struct gbp_environment {
  int32_t nrows, ncols;      /* tiles.grid: rows of box-drawing characters */
  const uint32_t *tiles;     /* [nrows*ncols] Unicode code points, row-major (TileGrid::get_tile) */
  float tile_size;           /* tiles.settings.tile-size */
  float path_width;          /* tiles.settings.path-width, a fraction of the tile */
  uint32_t resolution;       /* tiles.settings.sdf.resolution: pixels per tile */
  float expansion;           /* tiles.settings.sdf.expansion */
  float blur;                /* tiles.settings.sdf.blur (sigma = blur * resolution pixels) */
  int32_t n_obstacles;       /* env.obstacles.len() */
  const gbp_obstacle_t *obstacles;  /* [n_obstacles] in file order (first hit wins, lib.rs:293-333) */
  const double *polygon_points;     /* (x, y) pairs of every polygon obstacle */
} gbp_environment_t;

/* Replaces env_to_png::env_to_sdf_image (crates/env_to_png/src/lib.rs:149-163: env_to_image :166-207 +
 * image::imageops::blur), called at scenario load (simulation_loader.rs:153-161): rasterises the tile grid
 * and the placeable obstacles and blurs the image on the device.  rgb8: [nrows*resolution][ncols*resolution][3], R = G = B, row 0 = top. */
int gbp_env_to_sdf_image(const gbp_environment_t *env, int32_t device, uint8_t *rgb8);
/* Same, but the image never leaves the device: it becomes the world's SDF (gbp_world_set_sdf). */
int gbp_world_set_sdf_from_environment(gbp_world_t *w, const gbp_environment_t *env);

/* Replaces RobotBundle::new for n robots (robot.rs:1134-1355): V variables
 * (prior 1e30*I on first/last, non-finite -> 0 on the rest, variable.rs:146-148),
 * V-1 Dynamic, V-2 Obstacle, V-2 Tracking factors.
 *   radii[n]            f32 robot radius (t0 = radius/2/target_speed in f32, :1225)
 *   timesteps[V]        variable_timesteps (shared by the world)
 *   init_means[n*V*4]   f64 initial variable means (already widened from f32, :1212-1217)
 *   positions[n*2]      f32 Transform.translation (x, z) (spawner.rs:530)
 *   wp_offsets[n+1], wp_xy[2*wp_offsets[n]]  f32 waypoint polyline of each robot
 *                       (mission route; also the TrackingFactor path, :1316-1322)
 * Robot ids are assigned consecutively in insertion order.  Sharded worlds: before gbp_world_commit_shards every
 * shard takes its own robots; afterwards (robots spawned while the simulation runs) only the last shard does — the
 * new ids lie above every existing one — and every rank calls gbp_world_commit_shards again before the next
 * collective call. */
int gbp_world_add_robots(gbp_world_t *w, int32_t n, const float *radii,
                         const uint32_t *timesteps, const double *init_means,
                         const float *positions, const int32_t *wp_offsets,
                         const float *wp_xy);

int32_t gbp_world_num_robots(const gbp_world_t *w);

/* ---- multi-GPU: one swarm spatially partitioned over several B200s -------------
 * The reference iterates all robots in one process (Bevy `Query<&mut FactorGraph>`,
 * robot.rs:1769-1861); cross-robot traffic is the two delivery loops of
 * iterate_gbp_v2 (robot.rs:1814-1831, :1843-1858) and the horizon change_prior
 * messages (robot.rs:2272-2282).  Here each GPU ("shard") owns a contiguous range of
 * robot ids in rank order and keeps ghost copies of the robots of other shards
 * that are within comms range of its own.  The per-sub-step halo (published belief
 * records of border robots) moves by NCCL send/recv over NVLink; it is the only
 * communication on the iteration path.  Connectivity, inbox order and robot_number
 * stay globally bit-exact.
 *
 * One process per GPU:
 *   rank 0: gbp_comm_unique_id(id); broadcast id to every rank by any means;
 *   every rank: w = gbp_world_create_shard(cfg, device, rank, world_size, id);
 *               gbp_world_add_robots(w, <the robots this rank owns>) ...;
 *               gbp_world_commit_shards(w);
 *   then the per-tick calls below, made by every rank in the same order
 *   (update_topology / update_prior_* / iterate* / *_iteration / step are collective;
 *   set_comms and change_prior_of_variable act on the own robots but must be
 *   called by every rank in the same tick, with m = 0 if a rank has nothing to change).
 * Robot indices in per-robot arrays are LOCAL (0..num_robots-1); neighbour ids returned
 * by gbp_world_read_connections are GLOBAL (local index + gbp_world_first_global_id). */
#define GBP_COMM_ID_BYTES 128
int gbp_comm_unique_id(uint8_t *id /* [GBP_COMM_ID_BYTES] */);
gbp_world_t *gbp_world_create_shard(const gbp_config_t *cfg, int32_t device, int32_t rank,
                                    int32_t world_size, const uint8_t *id);
/* The same sharding code path with every shard in THIS process on ONE device (transfers are
 * device-to-device copies on a shared stream): out[world_size].  Collective calls made
 * on any member run for the whole group; destroy every member. */
int gbp_world_create_local_shards(const gbp_config_t *cfg, int32_t device, int32_t world_size,
                                  gbp_world_t **out);
/* Collective; after the last gbp_world_add_robots: fixes the global ids (shard q owns
 * [first_global_id, first_global_id + num_robots) in rank order). */
int gbp_world_commit_shards(gbp_world_t *w);
int64_t gbp_world_first_global_id(const gbp_world_t *w);
int64_t gbp_world_num_robots_global(const gbp_world_t *w);
int32_t gbp_world_num_ghosts(const gbp_world_t *w);

/* ---- per-tick systems (RobotPlugin FixedUpdate chain, robot.rs:85-108) -- */
/* update_robot_neighbours + delete_interrobot_factors + create_interrobot_factors
 * (robot.rs:1362-1586; FactorGraph::delete_interrobot_factors_connected_to
 * factorgraph.rs:380-436; add_internal_edge/add_external_edge :304-353).
 * Sort-based spatial hash; connectivity, creation order and robot_number are
 * bit-exact with the reference's all-pairs search.
 * The search itself (it reads positions and the live factor lists only) is started by the preceding
 * gbp_world_update_prior_of_current_state / gbp_world_step on a side stream and runs next to the iterations; this call
 * waits for it, reads its sizes and applies it.  A search whose inputs changed since (robots added or removed) is
 * thrown away and repeated here; GBP_TOPO_EARLY=0 in the environment always searches here. */
int gbp_world_update_topology(gbp_world_t *w);

/* update_failed_comms result (robot.rs:1593-1601): antenna_active[n] (0/1), and
 * mission.state.idle() (robot.rs:1791,1806): idle[n] (0/1).  NULL = all active / none idle. */
int gbp_world_set_comms(gbp_world_t *w, const uint8_t *antenna_active, const uint8_t *idle);

/* Mission::next_waypoint (robot.rs:2214): index into each robot's polyline of the
 * waypoint the horizon moves toward; <0 or >= len means "no more waypoints". */
int gbp_world_set_waypoint_index(gbp_world_t *w, const int32_t *next_index);

/* reached_waypoint (robot.rs:2080-2176) for every robot, single-route missions: compares the estimated
 * position of one variable (f32 Vec2, variable.rs:127-130) with the next waypoint
 * (glam distance_squared in f32) and advances the waypoint index (Route::advance, robot.rs:431-438).
 * A criterion is ReachedWhenIntersects (gbp_config formation section): which variable and which distance. */
enum gbp_intersects_with { GBP_INTERSECTS_CURRENT = 0, GBP_INTERSECTS_HORIZON = 1, GBP_INTERSECTS_VARIABLE = 2 };
enum gbp_intersection_distance { GBP_DISTANCE_ROBOT_RADIUS = 0, GBP_DISTANCE_METER = 1 };
typedef struct gbp_reached_when {
  int32_t intersects_with; /* enum gbp_intersects_with */
  int32_t variable_index;  /* for GBP_INTERSECTS_VARIABLE; out of range -> last variable */
  int32_t distance;        /* enum gbp_intersection_distance */
  float meter;             /* for GBP_DISTANCE_METER */
} gbp_reached_when_t;
/* taskpoint: criterion for every waypoint but the last (mission.taskpoint_reached_when_intersects);
 * finished: for the last one (mission.finished_when_intersects).  out_reached[n] (may be NULL) is set
 * to 1 for the robots that advanced — the RobotReachedWaypoint events.  Runs for the own robots of
 * this world; in the reference it is a FixedUpdate system unordered with respect to the GBP chain. */
int gbp_world_reached_waypoint(gbp_world_t *w, const gbp_reached_when_t *taskpoint,
                               const gbp_reached_when_t *finished, uint8_t *out_reached);
/* Mission/Route target_index of every robot (what gbp_world_set_waypoint_index sets). */
int gbp_world_read_waypoint_index(gbp_world_t *w, int32_t *next_index);

/* update_robot_robot_collisions + RobotRobotCollisions bookkeeping (planner/collisions.rs:72-143,
 * :146-200, :455-493): bounding-sphere test of every connected pair (parry2d
 * BoundingSphere::intersects, f32) and the Free/Colliding state machine; a Free -> Colliding
 * transition is one collision.  Pairs are taken from the comms-radius connectivity, so
 * radius_a + radius_b <= comms radius is required (true for every shipped scenario).  Collective
 * for a sharded world; the counts are this shard's share (a pair counts on the shard of its
 * lower robot id): num_collisions = RobotRobotCollisions::num_collisions, colliding_now = pairs in
 * state Colliding.  Either pointer may be NULL (no host sync then). */
int gbp_world_update_robot_collisions(gbp_world_t *w, int64_t *num_collisions, int64_t *colliding_now);
/* The two counters of this shard as of the last update (no kernel runs). */
int gbp_world_read_collision_totals(gbp_world_t *w, int64_t *num_collisions, int64_t *colliding_now);
/* RobotRobotCollisions::get(entity) for every own robot: per_robot[n]. */
int gbp_world_read_robot_collisions(gbp_world_t *w, uint32_t *per_robot);

/* ---- robot-environment collisions (planner/collisions.rs:368-455) ------------------------------
 * The `Colliders` resource (gbp_global_planner::Colliders, filled by environment/map_generator.rs:141-536,
 * :537-1298): one convex parry2d shape per collider in its own frame plus an Isometry2.  kind 0 = Ball
 * (radius), 1 = Cuboid (half_extents), 2 = Triangle (3 vertices), 3 = ConvexPolygon (num_vertices >= 3,
 * counter-clockwise as ConvexPolygon::from_convex_hull leaves them); vertices index into `vertices_xy`. */
typedef struct gbp_collider {
  int32_t kind;
  float translation[2]; /* isometry.translation (world x, z) */
  float angle;          /* isometry.rotation angle, radians */
  float radius;
  float half_extents[2];
  int32_t first_vertex, num_vertices;
} gbp_collider_t;
/* Replaces the collider set and clears every robot's collision history (RobotEnvironmentCollisions::clear on
 * LoadSimulation / ReloadSimulation, collisions.rs:40-46).  Per handle: a shard tests its own robots. */
int gbp_world_set_environment_colliders(gbp_world_t *w, int32_t n, const gbp_collider_t *colliders,
                                        int32_t num_vertices, const float *vertices_xy);
/* update_robot_environment_collisions (collisions.rs:368-431): parry2d intersection_test of every own robot's Ball
 * at its Transform against every collider, CollisionHistory::update per (robot, collider) (:455-493); a
 * Free -> Colliding transition is one collision.  num_collisions = RobotEnvironmentCollisions::num_collisions(),
 * colliding_now = pairs in state Colliding; either pointer may be NULL (no host sync then). */
int gbp_world_update_environment_collisions(gbp_world_t *w, int64_t *num_collisions, int64_t *colliding_now);
/* RobotEnvironmentCollisions::get(entity) for every own robot: per_robot[n]. */
int gbp_world_read_environment_collisions(gbp_world_t *w, uint32_t *per_robot);

/* Which collider, and where (host functions, no device needed; gbp_collide_host.cpp).  The monitors above count; the
 * reference also keeps one entry per pair that ever hit with the Aabb intersection of every hit (collisions.rs:402-431,
 * :700-716, exported by export.rs:171-206, :552-555).  Hits are rare, so the host derives those entries from the
 * counters that moved (magics_b200/collisions.py) and asks these two about the few robots concerned.
 * hits_ball: out[k] = parry2d intersection_test(collider, Ball(radii[k]) at robots_xz[k]) — the very predicate
 * k_env_collisions runs, compiled for the host from the same header.  aabb: Collider::aabb()
 * (gbp_global_planner/src/lib.rs:94-98) as mins_maxs = {min x, min z, max x, max z}. */
int gbp_collider_hits_ball(const gbp_collider_t *collider, int32_t num_vertices, const float *vertices_xy, int32_t m,
                           const float *robots_xz, const float *radii, uint8_t *out);
int gbp_collider_aabb(const gbp_collider_t *collider, int32_t num_vertices, const float *vertices_xy, float *mins_maxs);

/* ---- PositionTracker / VelocityTracker (planner/tracking.rs:36-260) ------------------------------
 * Every robot carries two ring buffers of `capacity` samples and a repeating Timer of `sample_ns`
 * (spawner.rs:627-628: 10000 samples, 100 ms).  Memory: 32 bytes x capacity x robots on the device. */
int gbp_world_set_tracking_buffers(gbp_world_t *w, int32_t capacity, uint64_t sample_ns);
/* track_positions + track_velocities for one FixedUpdate (tracking.rs:117-137, :226-260): call once per tick after
 * update_prior_of_current_state and before the next set_comms.  delta_ns = Time<Fixed>::delta(), elapsed_seconds =
 * Time::elapsed_seconds_f64().  Robots whose Transform changed this tick (not idle, or added since the last call) tick
 * their Timer; when it fires the position (x, z) and the velocity over the time since the previous sample are pushed. */
int gbp_world_track(gbp_world_t *w, uint64_t delta_ns, double elapsed_seconds);
/* The raw rings of the own robots, slot-major: num_*[n] = samples pushed so far (sample k sits in slot
 * k % capacity); positions_xy / velocities_xy [capacity][2][n], velocity_timestamp / velocity_measured_over
 * [capacity][n] (VelocityMeasurement.timestamp, .measured_over in seconds).  Any pointer may be NULL. */
int gbp_world_read_tracks(gbp_world_t *w, uint32_t *num_positions, float *positions_xy, uint32_t *num_velocities,
                          float *velocities_xy, double *velocity_timestamp, double *velocity_measured_over);

/* update_prior_of_horizon_state (robot.rs:2182-2283) for every robot. */
int gbp_world_update_prior_of_horizon_state(gbp_world_t *w);
/* update_prior_of_current_state_v3 (robot.rs:2286-2338) for every robot;
 * also advances the f32 Transform used by the neighbour search. */
int gbp_world_update_prior_of_current_state(gbp_world_t *w);

/* FactorGraph::change_prior_of_variable (factorgraph.rs:494-528) for one
 * variable index of every robot listed: robots[m], new_means[m*4]. */
int gbp_world_change_prior_of_variable(gbp_world_t *w, int32_t variable_index, int32_t m,
                                       const int32_t *robots, const double *new_means);

/* Global-planner hand-off (planner/robot.rs:655-776: an RRT* path has arrived for a robot).
 * set_tracking_path: FactorGraph::modify_tracking_factors(|t| t.set_tracking_path(..)) (factorgraph.rs:1467-1477,
 *   factor/tracking.rs:134-136) + Route::update_waypoints (robot.rs:389-392): the polyline of each listed
 *   robot is replaced (wp_offsets[m+1], wp_xy as in gbp_world_add_robots; two or more points) and its next
 *   waypoint index becomes 1.
 * reset_variables: FactorGraph::reset_variables(means, first_last_sigma, inbetween_sigma) (factorgraph.rs:1541-1564):
 *   means[m*V*4]; belief mean and precision (sigma on the diagonal) replaced, information vector kept, every
 *   inbox of the robot's variables AND of its factors emptied.  The reference passes (1e30, INFINITY).
 * reset_tracking_factors: FactorGraph::reset_tracking_factors (factorgraph.rs:1566-1590): timeout of 10
 *   iterations on every tracking factor.
 * On a sharded world reset_variables is collective: every rank calls it in the same tick with its own robots
 * (m = 0 if it has none); the ids of the reset robots travel to the other shards, which freeze their own InterRobot
 * factors toward them at the zero linearisation point the emptied inbox gives (factor/mod.rs:336-349). */
int gbp_world_set_tracking_path(gbp_world_t *w, int32_t m, const int32_t *robots, const int32_t *wp_offsets,
                                const float *wp_xy);
int gbp_world_reset_variables(gbp_world_t *w, int32_t m, const int32_t *robots, const double *means,
                              double first_last_sigma, double inbetween_sigma);
int gbp_world_reset_tracking_factors(gbp_world_t *w, int32_t m, const int32_t *robots);

/* iterate_gbp_v2 (robot.rs:1769-1861) with the config's schedule. */
int gbp_world_iterate(gbp_world_t *w);
/* Same with an explicit schedule: n sub-steps of (internal[i], external[i]). */
int gbp_world_iterate_schedule(gbp_world_t *w, int32_t n, const uint8_t *internal,
                               const uint8_t *external);

/* The four FactorGraph half-iterations, each for every (non-idle / antenna-on)
 * robot, with the delivery loops of iterate_gbp_v2 folded in:
 *   internal_factor_iteration   factorgraph.rs:688-714
 *   internal_variable_iteration factorgraph.rs:762-790
 *   external_factor_iteration   factorgraph.rs:719-760 + delivery robot.rs:1814-1831
 *   external_variable_iteration factorgraph.rs:794-826 + delivery robot.rs:1843-1858
 * The internal pair and the external pair must each be called in this order
 * (as iterate_gbp_v2 does); the engine executes a pair as one fused pass when
 * the second half is requested and returns GBP_ERR_STATE on any other order.
 * The state "after the factor half, before the variable half" is therefore never
 * materialised: while a pair is open (state kept per swarm, so any shard of a group
 * may close what another opened) every read-back and every state-changing call —
 * read_beliefs, set_comms, change_prior_of_variable, reset_variables, remove_robots,
 * update_topology, iterate* — returns GBP_ERR_STATE instead of showing or changing
 * something the reference would not. */
int gbp_world_internal_factor_iteration(gbp_world_t *w);
int gbp_world_internal_variable_iteration(gbp_world_t *w);
int gbp_world_external_factor_iteration(gbp_world_t *w);
int gbp_world_external_variable_iteration(gbp_world_t *w);

/* One full tick = FixedUpdate chain items 1-7: update_topology, the two prior updates (one launch
 * for both when V >= 3 and message counting is off: they touch different variables of a robot), iterate. */
int gbp_world_step(gbp_world_t *w);

/* ---- setters used by the UI hooks (ui/settings.rs:437,495,590) ---------- */
/* FactorGraph::change_factor_enabled (factorgraph.rs:1529-1539); kind: 0 dyn 1 ir 2 obs 3 trk.
 * Disabling is exact (the factors stop updating and stop receiving; what they sent stays in the variables' inboxes).
 * RE-enabling differs for one factor update: the reference's factor resumes with the inbox it held when it was
 * switched off (a disabled factor drops what it is sent, factor/mod.rs:308-310), the engine rebuilds every factor
 * inbox from the current records. */
int gbp_world_change_factor_enabled(gbp_world_t *w, int32_t kind, uint8_t enabled);
/* FactorGraph::update_inter_robot_safety_distance_multiplier (factorgraph.rs:892) */
int gbp_world_set_safety_distance_multiplier(gbp_world_t *w, float multiplier);
int gbp_world_set_schedule(gbp_world_t *w, int32_t kind, int32_t internal, int32_t external);

/* ---- read-back (visualisers / export read these fields) ----------------- */
/* VariableBelief of every variable (variable.rs:39-54): any pointer may be NULL.
 *   eta[n*V*4], lam[n*V*16] row-major, mean[n*V*4], cov[n*V*16], valid[n*V] */
int gbp_world_read_beliefs(gbp_world_t *w, double *eta, double *lam, double *mean,
                           double *cov, uint8_t *valid);
/* The same read-back without stalling the engine: the gather runs on the engine's stream, the
 * device->host copies on a second stream, so the kernels of the next tick overlap the PCIe
 * transfer (host buffers should be page-locked, gbp_host_alloc_pinned).  The buffers are valid
 * after gbp_world_readback_wait; at most one read-back is in flight (a second call waits for the first). */
int gbp_world_read_beliefs_async(gbp_world_t *w, double *eta, double *lam, double *mean,
                                 double *cov, uint8_t *valid);
int gbp_world_readback_wait(gbp_world_t *w);
/* Transform.translation (x, z) of every robot, f32[n*2]. */
int gbp_world_read_positions(gbp_world_t *w, float *xy);
/* RobotDespawned (planner/robot.rs:2171-2172, despawn_entity_after): the listed robots (LOCAL indices) leave the
 * simulation.  From the next gbp_world_update_topology on they are in nobody's comms range, so every InterRobot
 * factor to or from them is deleted by the regular path (robot.rs:1386-1439); they are never iterated again and
 * gbp_world_reached_waypoint reports nothing for them.  Indices are stable: the slot stays, frozen in its last
 * state, and still counts in gbp_world_num_robots; gbp_world_read_removed tells which slots are dead.  On a
 * sharded world each rank removes its own robots; the peers learn it with the positions of the next topology pass. */
int gbp_world_remove_robots(gbp_world_t *w, int32_t m, const int32_t *robots);
int gbp_world_read_removed(gbp_world_t *w, uint8_t *removed);
/* MessageCount (factorgraph/mod.rs:103-137; FactorGraph::messages_sent / messages_received, factorgraph.rs:876-890,
 * read by export.rs:434-439 and the robot diagnostics).  Off by default; turn it on right after gbp_world_create,
 * before robots are added (the reference counts from graph creation on).  The counts are kept by small accounting
 * kernels next to every half-step, prior update and topology change - from who takes part, with the reference's
 * rules: a message counts as sent per inbox key of the updating node (not in change_prior), as received in
 * receive_message_from (a disabled factor swallows it uncounted), the counters of a deleted InterRobot factor go
 * with it.  counts[4 * r + k]: sent.internal, sent.external, received.internal, received.external of robot r.
 * Single-GPU worlds only. */
int gbp_world_set_message_counting(gbp_world_t *w, int32_t on);
int gbp_world_read_message_counts(gbp_world_t *w, int64_t *counts);
/* State of every Tracking factor, as the visualisers and the RRT* hand-off read it (factor/tracking.rs:62-90,
 * `Tracking.record`, `LastMeasurement { pos: Vec2, value }`), one entry per (robot,
 * variable): record[n*V], last_pos f32[n*V*2], last_value[n*V].  Variables 0 and V-1 have no Tracking factor;
 * their entries keep the initial values (record 0, the initial mean, 0.0). */
int gbp_world_read_tracking(gbp_world_t *w, int64_t *record, float *last_pos, double *last_value);
/* RobotConnections.robots_connected_with as CSR: offsets[n+1], neighbours (global robot
 * ids) sorted ascending; robot_number[e] = RobotNumberGenerator value of the FIRST (i=1)
 * InterRobot factor robot r created toward neighbours[e] (robot.rs:1527).
 * Pass capacity of the neighbour arrays; returns edge count or <0. */
int64_t gbp_world_read_connections(gbp_world_t *w, int64_t *offsets, int32_t *neighbours,
                                   int64_t *robot_number, int64_t capacity);
/* ObstacleFactor::measure pixel lookup (obstacle.rs:141-188) for m points:
 * xy[m*2] f64 -> px[m], py[m] (u32 after the saturating cast) and value[m]. */
int gbp_world_sdf_lookup(gbp_world_t *w, int32_t m, const double *xy, uint32_t *px,
                         uint32_t *py, double *value);
/* FactorGraph::factor_count / node_count style totals (factorgraph.rs:247-276):
 * out[0]=variables out[1]=dynamic out[2]=obstacle out[3]=tracking out[4]=interrobot. */
int gbp_world_node_counts(gbp_world_t *w, int64_t out[5]);
/* Which iterate kernel runs a robot is decided on the device per launch: k_iterate_axis (two lanes per
 * variable) while the robot's x and y chains are decoupled — no InterRobot factor inside its safety
 * distance, flat SDF, no Tracking message — and k_iterate otherwise; both produce the bits of the
 * reference's update order.  general_only = 1 sends every robot through k_iterate
 * (A/B tests); 2 does the same and runs a whole gbp_world_iterate as ONE cooperative launch (k_tick_fused) while the
 * swarm fits the GPU at once (a few thousand robots on one GPU, no message counting / profiling) — for the
 * reference's own scenarios of 10 - 50 robots, where a tick is otherwise 30 launches of microseconds each.
 * read: robots of this shard currently assigned to each kernel. */
int gbp_world_set_iterate_path(gbp_world_t *w, int32_t general_only);
int gbp_world_read_iterate_path(gbp_world_t *w, int64_t *robots_axis, int64_t *robots_general);
/* number of GPU kernels this handle has launched so far (bench `gpu_launches`) */
int64_t gbp_world_kernel_launches(const gbp_world_t *w);
/* Optional per-launch CUDA-event timing on the engine's stream (bench.py's
 * roofline leg): accumulated count / device milliseconds per kernel family. */
enum gbp_profile_kind {
  GBP_PROFILE_ITERATE_INT = 0,     /* k_iterate_axis<EXT=0,INT=1> (k_edge_messages + k_iterate when general_only) */
  GBP_PROFILE_ITERATE_EXT = 1,     /* k_iterate_axis<EXT=1,INT=0> */
  GBP_PROFILE_ITERATE_EXT_INT = 2, /* k_iterate_axis<EXT=1,INT=1>, the dominant kernel */
  GBP_PROFILE_TOPOLOGY = 3,        /* whole gbp_world_update_topology */
  GBP_PROFILE_PRIORS = 4,          /* horizon + current prior kernels */
  GBP_PROFILE_HALO = 5,            /* the send/recv part of the per-sub-step halo exchange */
  GBP_PROFILE_ITERATE_GENERAL = 6, /* k_edge_messages + k_iterate over the robots k_iterate_axis handed over (any EXT/INT) */
  GBP_PROFILE_TOPO_POSITIONS = 7,  /* sharded worlds: the exchange of every robot's position / radius / despawned flag */
  GBP_PROFILE_TOPO_SEARCH = 8,     /* neighbour search + diff against the live edges, up to the size read-back */
  GBP_PROFILE_TOPO_APPLY = 9,      /* sharded: header / cross-shard robot_number exchange; then the new edge set */
  GBP_PROFILE_ITERATE_BORDER = 10, /* sharded: k_iterate_axis over the border robots (any EXT/INT), launched before the rest */
  GBP_PROFILE_KINDS = 11
};
int gbp_world_set_profiling(gbp_world_t *w, int32_t on);
int gbp_world_read_profile(gbp_world_t *w, int32_t kind, int64_t *count, double *total_ms);
/* Page-locked host memory for the buffers passed to the upload / read-back calls
 * (any host pointer is accepted; pinned ones make the copies DMA at PCIe speed). */
void *gbp_host_alloc_pinned(size_t bytes);
void gbp_host_free_pinned(void *p);
/* device time helpers for bench.py: record/elapsed on the engine's own stream. */
int gbp_world_sync(gbp_world_t *w);
int gbp_world_timer_start(gbp_world_t *w);
int gbp_world_timer_stop_ms(gbp_world_t *w, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* GBP_B200_H */
