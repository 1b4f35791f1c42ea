"""Mission / Route clocks of the robots (`Mission::local` and `Mission::global`), kept on the host next to the engine.

The engine advances a robot's waypoint index on the device (`gbp_world_reached_waypoint`, planner/robot.rs:2080-2176) and
tells the caller who advanced; WHEN a route / mission started and finished is bookkeeping of the reference's `Mission`
component (robot.rs:815-1012) and `Route` (robot.rs:331-490) that only the exporter reads (export.rs:381-409).  This file
restates that bookkeeping so `magics_b200.export` can fill `mission.started_at / finished_at / routes` the way the
reference does — including `Route::advance` adding the route's start time to an already absolute clock (robot.rs:436-438).

Times are `Time<Fixed>::elapsed()` durations in integer nanoseconds, converted like `Duration::as_secs_f64`.
Missions built by `Mission::global` (one route per pair of taskpoints, filled in by the global planner) are
`GlobalMissionState` + `MissionClock.progress`; the RRT* SEARCH itself (the third-party `rrt` crate behind
gbp_global_planner/src/rrtstar.rs) is outside this repo's scope — the caller supplies the planner.  What happens when a
planned path arrives for a robot — the host arithmetic in front of `set_tracking_path` / `reset_variables` /
`reset_tracking_factors` — is `path_arrival` / `apply_global_paths` below.
"""
from __future__ import annotations

from dataclasses import dataclass, field


def secs_f64(elapsed_ns: int) -> float:
    """`Duration::as_secs_f64`: whole seconds plus nanoseconds / 1e9, each converted on its own."""
    elapsed_ns = int(elapsed_ns)
    return float(elapsed_ns // 1_000_000_000) + float(elapsed_ns % 1_000_000_000) / 1e9


@dataclass
class RouteClock:
    """`Route` (robot.rs:331-346): waypoints, the index of the next one, start and finish time."""
    waypoints: list
    started_at: float
    target_index: int = 1  # Route::new (robot.rs:383): the first waypoint is the initial pose
    finished_at: float | None = None

    def is_completed(self) -> bool:
        return self.target_index >= len(self.waypoints)

    def advance(self, elapsed_ns: int) -> None:
        """Route::advance (robot.rs:432-439)."""
        if self.target_index < len(self.waypoints):
            self.target_index += 1
        if self.is_completed() and self.finished_at is None:
            self.finished_at = secs_f64(elapsed_ns) + self.started_at


@dataclass
class MissionState:
    """`Mission::local` (robot.rs:835-857): one route over all waypoints, taskpoints = (first, last)."""
    route: RouteClock
    started_at: float
    finished_at: float | None = None
    completed: bool = False
    taskpoints: list = field(default_factory=list)

    def advance_to_next_waypoint(self, elapsed_ns: int) -> None:
        """Mission::advance_to_next_waypoint / next_route (robot.rs:995-1006, 961-993) for an Active local mission:
        the only route completing ends the mission at the fixed clock's `elapsed()`."""
        if self.completed:
            return
        self.route.advance(elapsed_ns)
        if self.route.is_completed():
            self.completed = True
            self.finished_at = secs_f64(elapsed_ns)

    @property
    def routes(self) -> list:
        return [self.route]

    @property
    def idle(self) -> bool:
        return False  # Mission::local starts Active (robot.rs:853)


@dataclass
class GlobalMissionState:
    """`Mission::global` (robot.rs:859-905) with the state machine `progress_missions` drives for the RrtStar planning
    strategy (robot.rs:562-812): one route per pair of consecutive taskpoints, created when the previous one completes;
    a new route starts Idle, asks the global planner for a path (Idle { waiting_for_waypoints: true }), becomes Active
    when the path has arrived and replaced the route's two waypoints, and hands over to the next route — or completes
    the mission — when its last waypoint is reached."""
    taskpoints: list
    started_at: float
    routes: list = field(default_factory=list)
    active_route: int = 0
    finished_at: float | None = None
    state: str = "idle"  # "idle" | "waiting" | "active" | "completed"
    pending_path: object = None  # the planner's answer, held until the next progress pass (the task is polled per frame)

    def __post_init__(self):
        if len(self.taskpoints) < 2:
            raise ValueError("a mission has at least two taskpoints (min_len_vec::TwoOrMore)")
        if not self.routes:
            self.routes.append(RouteClock(list(self.taskpoints[:2]), self.started_at))  # Mission::new :881-890

    @property
    def route(self) -> RouteClock:
        return self.routes[min(self.active_route, len(self.routes) - 1)]

    @property
    def completed(self) -> bool:
        return self.state == "completed"

    @property
    def idle(self) -> bool:
        """MissionState::idle(): the robot neither iterates nor moves (robot.rs:1794-1851, :2212, :2303)."""
        return self.state in ("idle", "waiting")

    def next_route(self, elapsed_ns: int) -> None:
        """Mission::next_route (robot.rs:961-993)."""
        if self.completed:
            return
        self.active_route += 1
        if self.active_route >= len(self.taskpoints) - 1:
            self.state = "completed"
            self.finished_at = secs_f64(elapsed_ns)
        else:
            k = self.active_route
            self.routes.append(RouteClock(list(self.taskpoints[k:k + 2]), secs_f64(elapsed_ns)))
            self.state = "idle"

    def advance_to_next_waypoint(self, elapsed_ns: int) -> None:
        """Mission::advance_to_next_waypoint (robot.rs:995-1006): only an Active mission advances."""
        if self.state != "active":
            return
        self.route.advance(elapsed_ns)
        if self.route.is_completed():
            self.next_route(elapsed_ns)


class MissionClock:
    """One `MissionState` per robot of a world, in the world's robot order.

    spawn(...) when robots are added (with the `started_at` the spawner hands to `Mission::local`,
    `time.elapsed_seconds_f64()` at spawn, spawner.rs / robot.rs:1145), observe(...) after every
    `reached_waypoint` call with the flags it returned."""

    def __init__(self):
        self.missions: list[MissionState] = []

    def spawn(self, waypoints, started_at: float, planning_strategy: str = "only-local") -> None:
        """waypoints: per new robot a sequence of (x, y) with at least two entries; planning_strategy "only-local"
        (Mission::local) or "rrt-star" (Mission::global), robot.rs:1287-1300."""
        for wps in waypoints:
            wps = [(float(p[0]), float(p[1])) for p in wps]
            if len(wps) < 2:
                raise ValueError("a route has at least two waypoints (min_len_vec::TwoOrMore)")
            if planning_strategy == "rrt-star":
                self.missions.append(GlobalMissionState(taskpoints=wps, started_at=float(started_at)))
            else:
                self.missions.append(MissionState(RouteClock(wps, float(started_at)), float(started_at),
                                                  taskpoints=[wps[0], wps[-1]]))

    def observe(self, reached, elapsed_ns: int) -> None:
        """reached: per robot, whether `reached_waypoint` advanced it in the tick whose fixed clock reads elapsed_ns."""
        if len(reached) != len(self.missions):
            raise ValueError(f"{len(reached)} flags for {len(self.missions)} missions")
        for m, hit in zip(self.missions, reached):
            if hit:
                m.advance_to_next_waypoint(elapsed_ns)

    def next_waypoint_index(self):
        """Route::target_index per robot — equals the engine's `read_waypoint_index` at every tick."""
        return [m.route.target_index for m in self.missions]

    def mission_data(self, robot: int, now_ns: int) -> dict:
        """`MissionData` as the exporter fills it (export.rs:381-409): unfinished clocks read the export's `now`."""
        m = self.missions[robot]
        now = secs_f64(now_ns)
        return {
            "waypoints": [[p[0], p[1]] for p in m.taskpoints],
            "started_at": m.started_at,
            "finished_at": m.finished_at if m.finished_at is not None else now,
            "routes": [{"waypoints": [[p[0], p[1]] for p in r.waypoints], "started_at": r.started_at,
                        "finished_at": r.finished_at if r.finished_at is not None else now} for r in m.routes],
        }

    def idle_mask(self):
        """MissionState::idle() per robot: the `idle` argument of `set_comms`."""
        return [1 if m.idle else 0 for m in self.missions]

    def progress(self, world, elapsed_ns: int, planner, target_speed: float, planning_horizon: float, colliders=(),
                 rng=None) -> None:
        """`progress_missions` (robot.rs:562-812) for the RrtStar missions of `world`'s robots, once per frame:
        Idle -> the planner is asked for a path from taskpoint k to k + 1, `planner(start, end, colliders, rng)` ->
        sequence of (x, y) or None (the reference spawns an async RRT* task; the search itself is not part of this repo)
        -> Idle { waiting }; waiting -> the path arrives on the next pass: `apply_global_paths`, Active (a failed search
        goes back to Idle and is tried again); Active with a completed route -> `next_route`."""
        arrived, paths = [], []
        for r, m in enumerate(self.missions):
            if not isinstance(m, GlobalMissionState) or m.completed:
                continue
            if m.state == "waiting":
                if m.pending_path is None or len(m.pending_path) < 2:
                    m.state = "idle"  # Err(e) / empty Path: "try again" (robot.rs:786-791)
                else:
                    arrived.append(r)
                    paths.append(m.pending_path)
                    m.state = "active"
                m.pending_path = None
            elif m.state == "idle":
                k = m.active_route
                m.pending_path = planner(m.taskpoints[k], m.taskpoints[k + 1], colliders, rng)
                m.state = "waiting"
            elif m.state == "active" and m.route.is_completed():
                m.next_route(elapsed_ns)
        if arrived:
            apply_global_paths(world, arrived, paths, target_speed, planning_horizon, self)


# ---- the global planner's path arrives (planner/robot.rs:655-776) ---------------------------------------------------
def path_arrival(path, target_speed: float, planning_horizon: float, num_variables: int):
    """What `update_robot_mission` computes on the host when a robot's RRT* task has finished, f32 as written:

    waypoints (m, 4)  per path point `from` (the last one paired with itself): (from, target_speed * normalize(from - to)),
                      NaN directions -> 0 (robot.rs:655-668).  The direction is `from - to` in the reference, i.e. the
                      velocity part points back along the path: kept;
    tracking  (m, 2)  the positions of those waypoints — `set_tracking_path` of every tracking factor (:676-684);
    means  (V, 4) f64 the linearisation points `reset_variables` receives (:694-758): with start / next the first two
                      waypoints AS 4-VECTORS, dir = next - start, next' = start + min(speed * horizon, 0.9 |dir|) *
                      normalize(dir); variable i sits at lerp(start.xy, next'.xy, i / V) with velocity
                      (target_speed * normalize(dir)).xy.

    The caller then hands `tracking` to `World.set_tracking_path` (which also replaces the route's waypoints and sets
    the next index to 1, Route::update_waypoints :389-392), `means` to `World.reset_variables(…, 1e30, inf)` and calls
    `World.reset_tracking_factors` — `apply_global_paths` below.  glam (Vec2 / Vec4 normalize, length, lerp) is a
    third-party crate: restated with every f32 operation rounded on its own, parity unpinned in the last bit."""
    import numpy as np

    f = np.float32
    p = np.asarray(path, f).reshape(-1, 2)
    if p.shape[0] < 2:
        raise ValueError("a path has at least two points (min_len_vec::TwoOrMore)")
    speed = f(target_speed)
    wps = np.zeros((p.shape[0], 4), f)
    for k in range(p.shape[0]):
        frm, to = p[k], p[min(k + 1, p.shape[0] - 1)]
        d = frm - to
        n = np.sqrt(f(d[0] * d[0] + d[1] * d[1]))
        with np.errstate(divide="ignore", invalid="ignore"):
            d = d * (f(1.0) / n)  # glam Vec2::normalize: self * length_recip()
        if np.isnan(d).any():
            d = np.zeros(2, f)
        wps[k] = (frm[0], frm[1], speed * d[0], speed * d[1])
    start, nxt = wps[0], wps[1]
    dirv = nxt - start
    with np.errstate(divide="ignore", invalid="ignore"):
        length = np.sqrt(f(f(dirv[0] * dirv[0] + dirv[2] * dirv[2]) + f(dirv[1] * dirv[1] + dirv[3] * dirv[3])))
        dn = dirv / length  # glam Vec4 (SSE2): self / sqrt(dot)
    l = f(speed * f(planning_horizon))
    mx = f(length * f(0.9))
    s = l if l < mx else mx
    nxt2 = start + s * dn
    vel = speed * dn
    V = int(num_variables)
    means = np.zeros((V, 4), np.float64)
    for i in range(V):
        r = f(i) / f(V)
        pos = start[:2] + (nxt2[:2] - start[:2]) * r  # glam lerp: self + (rhs - self) * s
        means[i] = (pos[0], pos[1], vel[0], vel[1])
    return wps, wps[:, :2].copy(), means


def apply_global_paths(world, robots, paths, target_speed: float, planning_horizon: float, clock: "MissionClock | None" = None):
    """The device side of the hand-off for the listed robots (robot.rs:676-766): new tracking path, variables reset
    around the first path segment, tracking factors reset; the robots' routes (and their host clocks) restart at index 1."""
    import numpy as np

    robots = [int(r) for r in robots]
    if not robots:
        return
    arrivals = [path_arrival(p, target_speed, planning_horizon, world.V) for p in paths]
    world.set_tracking_path(robots, [a[1] for a in arrivals])
    world.reset_variables(robots, np.stack([a[2] for a in arrivals]), 1e30, float("inf"))
    world.reset_tracking_factors(robots)
    if clock is not None:
        for r, a in zip(robots, arrivals):
            route = clock.missions[r].route  # Route::update_waypoints (robot.rs:389-392)
            route.waypoints = [(float(x), float(y)) for x, y in a[1]]
            route.target_index = 1
