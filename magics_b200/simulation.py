"""Headless run of one of the reference's scenarios: the systems around the GBP iteration in the reference's order.

The reference has no headless mode (`--headless` is parsed and ignored, cli.rs:65-67 / main.rs:382-401); its experiment
scripts drive the windowed app and read the exported JSON.  `Simulation` plays the app for one scenario directory:

  per fixed step (Time<Fixed>, 1 / simulation.hz)
    1. FormationSpawner clock -> `spawn_formation` for every formation that is due   (spawner.rs:186-323, :415-600)
    2. `reached_waypoint` on the device, Mission / Route clocks on the host, `progress_missions` for formations planned
       by the global planner (path hand-off, idle robots), despawn          (robot.rs:2080-2176, :331-490, :562-812)
    3. `update_failed_comms` draws the antennas                                       (robot.rs:1592-1601)
    4. the RobotPlugin chain: neighbours, InterRobot factor deletion / creation, the two prior updates,
       `iterate_gbp_v2` = `world.step()`                                             (robot.rs:85-108)
    5. robot-robot / robot-environment collision monitors, entry lists               (planner/collisions.rs)
    6. position / velocity trackers                                                   (planner/tracking.rs:117-260)
  `export()`: the reference's ExportData (export.rs:112-277) with missions, obstacles and collision entries.

`world` is anything with the method surface of `magics_b200.World` — the tests drive the CPU oracle through the same
class.  Random inputs (radii, random placements, antenna draws) come from the numpy Generator, not from the reference's
WyRand stream.  The reference runs its collision systems in `Update` (once per rendered frame) and everything else in
`FixedUpdate`; here every system runs once per fixed step.
"""
from __future__ import annotations

import numpy as np

from .collisions import CollisionLog
from .environment import colliders as environment_colliders
from .export import export_data
from .mission import MissionClock, secs_f64


class Simulation:
    def __init__(self, scenario, world, rng=None, *, environment_collisions: bool = True, tracker_capacity: int = 10000,
                 tracker_sample_ns: int = 100_000_000, prng_seed: int = 0, global_planner=None):
        self.scenario, self.world = scenario, world
        self.rng = rng if rng is not None else np.random.default_rng(prng_seed)
        self.prng_seed = int(prng_seed)
        self.dt_ns = int(round(1e9 / scenario.hz))
        self.tick_count = 0
        self.clock = MissionClock()
        self.radii = np.zeros(0, np.float32)
        self.gone = np.zeros(0, bool)
        self.skipped_spawns = 0
        # formations with `planning-strategy: rrt-star` wait for a path from taskpoint to taskpoint:
        # `global_planner(start, end, colliders, rng) -> [(x, y), ...] | None` (magics_b200.planner.RRTStarPlanner is one;
        # the reference's search is the third-party `rrt` crate), default a straight line
        self.global_planner = global_planner or (lambda start, end, colliders, rng: [start, end])
        self._any_global = False
        # (taskpoint, finished) criterion of every robot's formation; `reached_waypoint` takes ONE pair per call
        self.criteria: list = []
        self.robot_criterion = np.zeros(0, np.int64)
        self.colliders = environment_colliders(scenario.env) if environment_collisions else []
        if self.colliders:
            world.set_environment_colliders(self.colliders)
        self.log = CollisionLog(world, self.radii, self.colliders)
        world.set_tracking_buffers(capacity=tracker_capacity, sample_ns=tracker_sample_ns)
        self._events: dict = {}
        self._events_until = 0

    @classmethod
    def on_gpu(cls, scenario, device: int = 0, **kw) -> "Simulation":
        """A `magics_b200.World` on `device` with the SDF generated on the device from the scenario's environment."""
        from .world import World

        world = World(scenario.cfg, device=device)
        world.set_sdf_from_environment(scenario.env)
        return cls(scenario, world, **kw)

    # ---- one fixed step ---------------------------------------------------------------------------------------------
    @property
    def elapsed_ns(self) -> int:
        return self.tick_count * self.dt_ns

    def _due(self, tick: int) -> list:
        if tick >= self._events_until:  # the spawner clock, computed a minute of simulated time at a time
            upto = max(2 * self._events_until, int(60 * self.scenario.hz), tick + 1)
            self._events = {}
            for t, k in self.scenario.spawn_events(upto):
                self._events.setdefault(t, []).append(k)
            self._events_until = upto
        return self._events.get(tick, [])

    def _spawn(self, tick: int) -> None:
        for k in self._due(tick):
            sw = self.scenario.spawn(k, self.rng)
            if sw is None:  # "failed to spawn formation ..., skipping" (spawner.rs:460-468) or `robots: 0`
                self.skipped_spawns += 1
                continue
            sw.add_to(self.world, set_sdf=False)
            self.radii = np.concatenate([self.radii, sw.radii])
            self.log.add_robots(sw.radii)
            self.gone = np.concatenate([self.gone, np.zeros(sw.n, bool)])
            fm = self.scenario.formations[k]
            crit = (fm.reached_when, fm.finished_when)
            if crit not in self.criteria:
                self.criteria.append(crit)
            self.robot_criterion = np.concatenate([self.robot_criterion, np.full(sw.n, self.criteria.index(crit))])
            strategy = fm.planning_strategy
            self._any_global |= strategy == "rrt-star"
            self.clock.spawn([sw.wp_xy[sw.wp_offsets[r]:sw.wp_offsets[r + 1]] for r in range(sw.n)],
                             started_at=secs_f64(self.elapsed_ns), planning_strategy=strategy)

    def _reached_waypoint(self) -> np.ndarray:
        """`reached_waypoint` for every robot under ITS formation's criteria.  The engine call applies one criterion pair
        to all robots and touches nothing but the waypoint index, so with several pairs in play (`Collaborative GP`) it
        runs once per pair from the same starting indices and every robot keeps the outcome of its own pair."""
        w = self.world
        if len(self.criteria) == 1:
            return np.asarray(w.reached_waypoint(*self.criteria[0]), bool)
        cur = np.array(w.read_waypoint_index(), np.int32)
        flags = np.zeros(cur.shape[0], bool)
        for g, crit in enumerate(self.criteria):
            members = self.robot_criterion == g
            if not members.any():
                continue
            w.set_waypoint_index(cur)  # undo what the previous pair's call did to the other robots
            reached = np.asarray(w.reached_waypoint(*crit), bool)
            new = np.asarray(w.read_waypoint_index(), np.int32)
            cur[members], flags[members] = new[members], reached[members]
        w.set_waypoint_index(cur)
        return flags

    def tick(self) -> None:
        sc, w = self.scenario, self.world
        self._spawn(self.tick_count)
        self.tick_count += 1
        if w.num_robots == 0:
            return
        self.clock.observe(self._reached_waypoint(), self.elapsed_ns)
        if self._any_global:  # progress_missions (robot.rs:562-812)
            self.clock.progress(w, self.elapsed_ns, self.global_planner, sc.cfg.target_speed, sc.planning_horizon,
                                self.colliders, self.rng)
        if sc.despawn:
            done = np.array([m.completed for m in self.clock.missions], bool) & ~self.gone
            if done.any():
                w.remove_robots(np.flatnonzero(done).astype(np.int32))
                self.gone |= done
        if sc.failure_rate > 0.0 or self._any_global:
            w.set_comms(antenna_active=sc.draw_antennas(w.num_robots, self.rng),
                        idle=np.asarray(self.clock.idle_mask(), np.uint8) if self._any_global else None)
        w.step()
        self.log.update_robot_collisions()
        if self.colliders:
            self.log.update_environment_collisions()
        w.track(self.dt_ns, secs_f64(self.elapsed_ns))

    def run(self, ticks: int | None = None, max_ticks: int = 100_000) -> int:
        """`ticks` fixed steps, or (ticks=None) until every spawned robot has completed its mission and nothing is left
        to spawn in the next minute of simulated time — what the experiment scripts wait for before they export.
        Returns the number of steps taken."""
        start = self.tick_count
        if ticks is not None:
            for _ in range(ticks):
                self.tick()
            return ticks
        while self.tick_count - start < max_ticks:
            self.tick()
            horizon = self.tick_count + int(60 * self.scenario.hz)
            pending = any(self._due(t) for t in range(self.tick_count, min(horizon, self._events_until)))
            if len(self.gone) and all(m.completed for m in self.clock.missions) and not pending:
                break
        return self.tick_count - start

    # ---- results ------------------------------------------------------------------------------------------------------
    def export(self, **kw) -> dict:
        """ExportData of the run so far; `makespan` is the virtual clock at export time, as in export.rs:357."""
        args = dict(scenario=self.scenario.name, makespan=secs_f64(self.elapsed_ns), delta_t=1.0 / self.scenario.hz,
                    prng_seed=self.prng_seed, radii=self.radii, missions=self.clock, now_ns=self.elapsed_ns,
                    colliders=self.colliders, collision_log=self.log)
        args.update(kw)
        return export_data(self.world, **args)


def main(argv=None) -> int:
    """python -m magics_b200.simulation <scenario directory> [--ticks N] [--export out.json] [--seed S] [--device D]

    Runs the scenario on the GPU engine (there is no CPU fallback) and writes the reference's ExportData."""
    import argparse
    import json

    from .scenarios import ReferenceScenario

    ap = argparse.ArgumentParser(prog="magics_b200.simulation", description=main.__doc__)
    ap.add_argument("scenario", help="a config/scenarios/<name>/ directory of the reference (config.toml, formation.yaml, "
                                     "environment.yaml)")
    ap.add_argument("--ticks", type=int, default=None, help="fixed steps to run (default: until every mission is complete)")
    ap.add_argument("--max-ticks", type=int, default=100_000)
    ap.add_argument("--export", default=None, help="path of the JSON to write (default: export_<scenario>_<seed>.json)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-environment-collisions", action="store_true")
    args = ap.parse_args(argv)
    sc = ReferenceScenario.from_directory(args.scenario)
    planner = None
    if any(f.planning_strategy == "rrt-star" for f in sc.formations):
        from .planner import RRTStarPlanner

        planner = RRTStarPlanner.from_config(sc.rrt)  # this repo's RRT* behind the reference's CollisionProblem
    sim = Simulation.on_gpu(sc, device=args.device, prng_seed=args.seed, global_planner=planner,
                            environment_collisions=not args.no_environment_collisions)
    steps = sim.run(args.ticks, max_ticks=args.max_ticks)
    data = sim.export()
    path = args.export or f"export_{sc.name.lower()}_{args.seed}.json"
    with open(path, "w") as f:
        json.dump(data, f)
    print(f"{sc.name}: {steps} steps, {sim.world.num_robots} robots spawned, {int(sim.gone.sum())} finished, "
          f"{len(data['collisions']['robots'])} robot-robot / {len(data['collisions']['environment'])} robot-environment "
          f"collision entries -> {path}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
