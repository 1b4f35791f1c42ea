"""Mission / Route clocks of single-route (`Mission::local`) robots, kept on the host next to the engine.

The engine advances a robot's waypoint index on the device (`gbp_world_reached_waypoint`, planner/robot.rs:2080-2176) and
tells the caller who advanced; WHEN a route / mission started and finished is bookkeeping of the reference's `Mission`
component (robot.rs:815-1012) and `Route` (robot.rs:331-490) that only the exporter reads (export.rs:381-409).  This file
restates that bookkeeping so `magics_b200.export` can fill `mission.started_at / finished_at / routes` the way the
reference does — including `Route::advance` adding the route's start time to an already absolute clock (robot.rs:436-438).

Times are `Time<Fixed>::elapsed()` durations in integer nanoseconds, converted like `Duration::as_secs_f64`.
Missions built by `Mission::global` (one route per pair of taskpoints, filled in by the RRT* planner) are outside this
repo's scope.
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


class MissionClock:
    """One `MissionState` per robot of a world, in the world's robot order.

    spawn(...) when robots are added (with the `started_at` the spawner hands to `Mission::local`,
    `time.elapsed_seconds_f64()` at spawn, spawner.rs / robot.rs:1145), observe(...) after every
    `reached_waypoint` call with the flags it returned."""

    def __init__(self):
        self.missions: list[MissionState] = []

    def spawn(self, waypoints, started_at: float) -> None:
        """waypoints: per new robot a sequence of (x, y) with at least two entries."""
        for wps in waypoints:
            wps = [(float(p[0]), float(p[1])) for p in wps]
            if len(wps) < 2:
                raise ValueError("a route has at least two waypoints (min_len_vec::TwoOrMore)")
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
            "routes": [{"waypoints": [[p[0], p[1]] for p in m.route.waypoints], "started_at": m.route.started_at,
                        "finished_at": m.route.finished_at if m.route.finished_at is not None else now}],
        }
