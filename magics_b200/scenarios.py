"""Synthetic swarm generators for the BASELINE.json configs.

Everything here is host-side INPUT construction in numpy: robot placement,
`variable_timesteps`, and the initial variable means of `RobotBundle::new`
(crates/magics/src/planner/robot.rs:1157-1192, f32 arithmetic).  The outputs
are plain arrays accepted by both `magics_b200.World.add_robots` and the test
oracle, so a parity test feeds identical bits to both sides.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from .config import GbpConfig
from .world import get_variable_timesteps

f32 = np.float32


@dataclass
class Swarm:
    cfg: GbpConfig
    radii: np.ndarray        # (n,) f32
    timesteps: np.ndarray    # (V,) u32
    init_means: np.ndarray   # (n, V, 4) f64
    positions: np.ndarray    # (n, 2) f32
    wp_offsets: np.ndarray   # (n+1,) i32
    wp_xy: np.ndarray        # (sum, 2) f32
    sdf: np.ndarray | None = None  # (h, w, 3) u8
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.radii.shape[0])

    def add_to(self, world, set_sdf=True):
        if set_sdf and self.sdf is not None:
            world.set_sdf(self.sdf)
        world.add_robots(self.radii, self.timesteps, self.init_means, self.positions, self.wp_offsets, self.wp_xy)
        return world

    def slice(self, lo: int, hi: int) -> "Swarm":
        a, b = int(self.wp_offsets[lo]), int(self.wp_offsets[hi])
        return replace(self, radii=self.radii[lo:hi], init_means=self.init_means[lo:hi],
                       positions=self.positions[lo:hi], wp_offsets=(self.wp_offsets[lo:hi + 1] - a).astype(np.int32),
                       wp_xy=self.wp_xy[a:b])


def lookahead_horizon(target_speed: float, planning_horizon: float) -> int:
    """(target_speed * planning_horizon) as u32, f32 product (spawner.rs:563-564)."""
    return int(f32(target_speed) * f32(planning_horizon))


def initial_means(start_xy: np.ndarray, goal_xy: np.ndarray, timesteps: np.ndarray, target_speed: float,
                  planning_horizon: float) -> np.ndarray:
    """RobotBundle::new initial variable means (robot.rs:1157-1192) for routes start -> goal.

    Vec4 state (x, y, vx, vy) in f32; the initial velocity points at the first
    waypoint with magnitude target_speed (spawner.rs:470-482) and the last
    waypoint inherits it (spawner.rs:528-530), so start2goal has zero velocity part.
    """
    s = np.asarray(start_xy, f32)
    g = np.asarray(goal_xy, f32)
    d = g - s
    length = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32)
    safe = np.where(length > 0, length, f32(1))
    unit = (d / safe[:, None]).astype(f32)
    unit[length == 0] = 0
    vel = (unit * f32(target_speed)).astype(f32)
    start4 = np.concatenate([s, vel], axis=1).astype(f32)
    reach = np.minimum(length, f32(planning_horizon) * f32(target_speed)).astype(f32)
    s2g4 = np.concatenate([d, np.zeros_like(d)], axis=1).astype(f32)
    inv_len = (f32(1) / safe).astype(f32)
    horizon4 = (start4 + reach[:, None] * (s2g4 * inv_len[:, None]).astype(f32)).astype(f32)
    ts = np.asarray(timesteps, np.uint32)
    frac = (ts.astype(f32) / f32(ts[-1])).astype(f32)  # variable_timestep as f32 / last as f32
    diff = (horizon4 - start4).astype(f32)
    means = (start4[:, None, :] + (diff[:, None, :] * frac[None, :, None]).astype(f32)).astype(f32)
    return means.astype(np.float64)


def _finish(cfg, radii, starts, goals, timesteps, planning_horizon, sdf=None, name="", waypoints=None, meta=None):
    n = starts.shape[0]
    if waypoints is None:
        wp_xy = np.stack([starts, goals], axis=1).reshape(-1, 2).astype(f32)
        wp_off = (np.arange(n + 1) * 2).astype(np.int32)
        first = goals
    else:
        wp_off = np.zeros(n + 1, np.int32)
        wp_off[1:] = np.cumsum([len(w) for w in waypoints])
        wp_xy = np.concatenate(waypoints, axis=0).astype(f32)
        first = np.stack([w[1] for w in waypoints]).astype(f32)
    means = initial_means(starts, first, timesteps, cfg.target_speed, planning_horizon)
    return Swarm(cfg=cfg, radii=np.asarray(radii, f32), timesteps=np.asarray(timesteps, np.uint32), init_means=means,
                 positions=np.asarray(starts, f32), wp_offsets=wp_off, wp_xy=wp_xy, sdf=sdf, name=name,
                 meta=meta or {})


def white_sdf(w=200, h=200):
    return np.full((h, w, 3), 255, np.uint8)


def synthetic_sdf(w, h, seed=0, n_rect=24, blur_px=4.0):
    """Deterministic SDF-like image: dark rectangles on white, Gaussian blurred.
    (Stand-in for env_to_png::env_to_sdf_image, which is SURVEY §8 next-2.)"""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    img = np.full((h, w), 255.0)
    for _ in range(n_rect):
        rw, rh = rng.integers(max(2, w // 40), max(3, w // 8)), rng.integers(max(2, h // 40), max(3, h // 8))
        x0, y0 = rng.integers(0, w - rw), rng.integers(0, h - rh)
        img[y0:y0 + rh, x0:x0 + rw] = 0.0
    img = gaussian_filter(img, blur_px)
    u8 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return np.repeat(u8[:, :, None], 3, axis=2)


def circle(n=30, circle_radius=50.0, robot_radius=1.5, cfg: GbpConfig | None = None, planning_horizon=5.0,
           lookahead_multiple=3):
    """Config 1: `Circle Experiment` geometry with the default config.toml scalars."""
    cfg = cfg or GbpConfig(world_width=100.0, world_height=100.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, planning_horizon), lookahead_multiple)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    ang = (np.arange(n, dtype=np.float64) * (2.0 * np.pi / n))
    starts = np.stack([circle_radius * np.cos(ang), circle_radius * np.sin(ang)], axis=1).astype(f32)
    goals = (-starts).astype(f32)
    return _finish(cfg, np.full(n, robot_radius, f32), starts, goals, ts, planning_horizon, sdf=white_sdf(),
                   name=f"circle-{n}")


def rings(n=100_000, spacing=8.0, ring_gap=44.0, r0=400.0, cfg: GbpConfig | None = None, planning_horizon=5.0,
          lookahead_multiple=3, robot_radius=1.0, seed=0):
    """Config 4: synthetic circle swarm, concentric rings with `spacing` metres
    between robots along a ring (K ~ 4 at comms radius 20) and rings further
    apart than 2x the comms radius; goal = antipode on the own ring."""
    cfg = cfg or GbpConfig(target_speed=3.6, world_width=100.0, world_height=100.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, planning_horizon), lookahead_multiple)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    xs, ys = [], []
    left, k = n, 0
    while left > 0:
        rad = r0 + k * ring_gap
        m = min(left, max(8, int(2.0 * np.pi * rad / spacing)))
        ang = np.arange(m, dtype=np.float64) * (2.0 * np.pi / max(m, int(2.0 * np.pi * rad / spacing)))
        xs.append(rad * np.cos(ang))
        ys.append(rad * np.sin(ang))
        left -= m
        k += 1
    starts = np.stack([np.concatenate(xs), np.concatenate(ys)], axis=1).astype(f32)
    goals = (-starts).astype(f32)
    return _finish(cfg, np.full(n, robot_radius, f32), starts, goals, ts, planning_horizon, sdf=white_sdf(),
                   name=f"rings-{n}", meta={"rings": k})


def lattice(nx=1000, ny=1000, pitch=12.0, cfg: GbpConfig | None = None, planning_horizon=5.0, lookahead_multiple=3,
            robot_radius=1.0, goal_ahead=500.0, rows: tuple[int, int] | None = None):
    """Config 5: uniform grid, every interior robot has K = 8 at comms radius 20
    (4 at `pitch`, 4 at pitch*sqrt(2)); all head +x toward a goal 500 m ahead.
    Robot id = row-major (iy * nx + ix), so a contiguous id range is a horizontal slab of rows.
    `rows=(lo, hi)` builds only the robots of rows [lo, hi) of the same global lattice (what one
    rank of a multi-GPU run owns), bit-identical to the corresponding slice of the whole."""
    cfg = cfg or GbpConfig(target_speed=3.6, world_width=100.0, world_height=100.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, planning_horizon), lookahead_multiple)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    r0, r1 = (0, ny) if rows is None else rows
    ix, iy = np.meshgrid(np.arange(nx), np.arange(r0, r1))
    x = (ix.reshape(-1) - (nx - 1) / 2.0) * pitch
    y = (iy.reshape(-1) - (ny - 1) / 2.0) * pitch
    starts = np.stack([x, y], axis=1).astype(f32)
    goals = (starts + np.array([goal_ahead, 0.0], f32)).astype(f32)
    n = nx * (r1 - r0)
    return _finish(cfg, np.full(n, robot_radius, f32), starts, goals, ts, planning_horizon, sdf=white_sdf(),
                   name=f"lattice-{nx}x{ny}", meta={"nx": nx, "ny": ny, "pitch": pitch, "rows": (r0, r1)})


def dense_lattice(nx=500, ny=500, pitch=2.0, comms_radius=4.5, cfg: GbpConfig | None = None, planning_horizon=5.0,
                  lookahead_multiple=3, robot_radius=1.0, goal_ahead=500.0):
    """Stress workload (not a BASELINE config): a lattice tighter than the safety distance 2.2 * radius, so the
    InterRobot factors of the four nearest neighbours are ACTIVE (no `skip` exit), with K = 20 neighbours inside the
    comms radius, and rows alternately heading +x / -x so that neighbouring rows shear past each other: InterRobot
    factors are created and deleted every tick.  Everything the lattice-1M benchmark never exercises."""
    cfg = cfg or GbpConfig(target_speed=3.6, comms_radius=comms_radius, world_width=100.0, world_height=100.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, planning_horizon), lookahead_multiple)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    x = (ix.reshape(-1) - (nx - 1) / 2.0) * pitch
    y = (iy.reshape(-1) - (ny - 1) / 2.0) * pitch
    starts = np.stack([x, y], axis=1).astype(f32)
    sign = np.where(iy.reshape(-1) % 2 == 0, 1.0, -1.0)
    goals = (starts + np.stack([sign * goal_ahead, np.zeros_like(sign)], axis=1)).astype(f32)
    n = nx * ny
    return _finish(cfg, np.full(n, robot_radius, f32), starts, goals, ts, planning_horizon, sdf=white_sdf(),
                   name=f"dense-{nx}x{ny}", meta={"nx": nx, "ny": ny, "pitch": pitch})


def junction_twoway(per_lane=3, seed=0):
    """Config 2: `Structured Junction Twoway` scalars (all four factor kinds, V=12),
    12 lanes (4 arms x {left, straight, right}) through a '+' junction with a
    waypoint at the junction centre; lateral jitter from a fixed-seed PCG."""
    cfg = GbpConfig(sigma_factor_dynamics=0.1, sigma_factor_interrobot=0.005, sigma_factor_obstacle=0.005,
                    sigma_factor_tracking=0.15, safety_distance_multiplier=2.5, comms_radius=20.0, target_speed=5.0,
                    enable_tracking=1, world_width=100.0, world_height=100.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, 5.0), 3)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    rng = np.random.default_rng(seed)
    arms = np.array([[-45.0, 0.0], [45.0, 0.0], [0.0, -45.0], [0.0, 45.0]])
    wps, starts = [], []
    for a in range(4):
        for b in range(4):
            if a == b:
                continue
            for k in range(per_lane):
                direction = -arms[a] / np.linalg.norm(arms[a])
                normal = np.array([-direction[1], direction[0]])
                lateral = rng.uniform(-2.0, 2.0)
                s = arms[a] - direction * (6.0 * k) * 0 + direction * (-6.0 * k) + normal * lateral
                mid = normal * lateral * 0.5
                e = arms[b] + normal * rng.uniform(-2.0, 2.0)
                wps.append(np.array([s, mid, e], f32))
                starts.append(s)
    starts = np.array(starts, f32)
    n = starts.shape[0]
    # '+' shaped road: white (free) cross on dark background, blur sigma 2 px
    from scipy.ndimage import gaussian_filter

    img = np.zeros((200, 200))
    img[84:116, :] = 255.0
    img[:, 84:116] = 255.0
    u8 = np.clip(np.rint(gaussian_filter(img, 2.0)), 0, 255).astype(np.uint8)
    sdf = np.repeat(u8[:, :, None], 3, axis=2)
    goals = np.array([w[1] for w in wps], f32)
    return _finish(cfg, np.full(n, 1.0, f32), starts, goals, ts, 5.0, sdf=sdf, name=f"junction-twoway-{n}",
                   waypoints=wps)


def complex_environment(n=19, seed=0):
    """Config 3: `Collaborative Complex` sizes: world 250x175, SDF 2000x1400, V=12,
    obstacle factors on, tracking off; robots cross the map left <-> right."""
    cfg = GbpConfig(sigma_factor_interrobot=0.005, sigma_factor_obstacle=0.005, target_speed=5.0,
                    world_width=250.0, world_height=175.0)
    ts = get_variable_timesteps(lookahead_horizon(cfg.target_speed, 5.0), 3)
    cfg = replace(cfg, num_variables=int(ts.shape[0]))
    rng = np.random.default_rng(seed)
    y = np.linspace(-80.0, 80.0, n)
    side = np.where(np.arange(n) % 2 == 0, -1.0, 1.0)
    starts = np.stack([side * 118.0 + rng.uniform(-2, 2, n), y], axis=1).astype(f32)
    goals = np.stack([-side * 118.0, y[::-1]], axis=1).astype(f32)
    sdf = synthetic_sdf(2000, 1400, seed=seed, n_rect=40, blur_px=4.0)
    return _finish(cfg, np.full(n, 1.0, f32), starts, goals, ts, 5.0, sdf=sdf, name=f"complex-{n}")


# ---- the reference's own scenario inputs (BASELINE configs 1-3) -------------------------------------------
_SCHEDULES = {"centered": 0, "interleave-evenly": 1, "soon-as-possible": 2, "late-as-possible": 3,
              "half-beginning-half-end": 4}


def read_scenario_directory(path: str) -> dict:
    """The three input files of one `config/scenarios/<name>/` directory of the reference as plain data:
    `formation.yaml` (FormationGroup, gbp_config/src/formation.rs; serde's `!tag` variants kept as {"kind": tag, ...}),
    `environment.yaml` (gbp_environment/src/lib.rs) and the sections of `config.toml` the path reads ([gbp], [robot],
    [simulation].hz and the despawn flag; gbp_config/src/lib.rs).  This is the form tests/golden/scenarios.json holds
    (tests/golden/make_golden.py calls this function) and `ReferenceScenario` consumes."""
    import os
    import tomllib
    from dataclasses import asdict

    import yaml

    from .environment import Environment

    class Loader(yaml.SafeLoader):
        pass

    def tagged(loader, suffix, node):
        if isinstance(node, yaml.MappingNode):
            return {"kind": suffix, **loader.construct_mapping(node, deep=True)}
        if isinstance(node, yaml.SequenceNode):
            return {"kind": suffix, "value": loader.construct_sequence(node, deep=True)}
        return {"kind": suffix, "value": loader.construct_scalar(node)}

    Loader.add_multi_constructor("!", tagged)
    with open(os.path.join(path, "config.toml"), "rb") as f:
        cfg = tomllib.load(f)
    with open(os.path.join(path, "formation.yaml"), encoding="utf-8") as f:
        form = yaml.load(f.read(), Loader=Loader)
    with open(os.path.join(path, "environment.yaml"), encoding="utf-8") as f:
        env = Environment.from_yaml(f.read())
    name = os.path.basename(os.path.normpath(path))
    return {"source": f"config/scenarios/{name}/", "formations": form["formations"], "environment": asdict(env),
            "gbp": cfg["gbp"], "robot": cfg["robot"],
            "simulation": {k: cfg["simulation"][k] for k in ("hz", "despawn-robot-when-final-waypoint-reached")}}


class ReferenceScenario:
    """One of the reference's `config/scenarios/<name>/` directories, as extracted into tests/golden/scenarios.json by
    tests/golden/make_golden.py: the formation group (`formation.yaml`), the environment (`environment.yaml`) and the
    [gbp] / [robot] scalars of `config.toml`.  It plays the roles of `spawn_formation` (spawner.rs:415-600) and of the
    FormationSpawner clock for a run at the scenario's fixed rate: `spawn_events` says which formation spawns at
    which tick, `spawn` turns one formation into the arrays `World.add_robots` / the oracle take.

    Inputs the reference draws from its PRNG (robot radii in [radius.min, radius.max], random placement along a line
    segment) come from the numpy Generator passed to `spawn`."""

    def __init__(self, name: str, golden_path: str | None = None, data: dict | None = None):
        import json
        import os

        from .environment import Environment, Obstacle
        from .formation import formation_from_dict

        if data is None:
            path = golden_path or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                                               "golden", "scenarios.json")
            data = json.load(open(path))[name]
        d = data
        self.name = name
        e = dict(d["environment"])
        e["obstacles"] = [Obstacle(**{k: (tuple(map(tuple, v)) if k == "points" else tuple(v) if isinstance(v, list) else v)
                                      for k, v in o.items()}) for o in e.get("obstacles") or []]
        self.env = Environment(**e)
        self.formations = [formation_from_dict(f) for f in d["formations"]]
        g, r = d["gbp"], d["robot"]
        en = g.get("factors-enabled", {})
        it = g.get("iteration-schedule", {})
        self.world_w, self.world_h = self.env.world_size
        self.hz = float(d["simulation"]["hz"])
        self.despawn = bool(d["simulation"]["despawn-robot-when-final-waypoint-reached"])
        self.planning_horizon = float(r["planning-horizon"])
        self.lookahead_multiple = int(g.get("lookahead-multiple", 3))
        self.radius_range = (float(r["radius"]["min"]), float(r["radius"]["max"]))
        self.failure_rate = float(r["communication"].get("failure-rate", 0.0))
        self.rrt: dict = {}
        trk = g.get("tracking", {})
        cfg = GbpConfig(
            sigma_factor_dynamics=float(g["sigma-factor-dynamics"]), sigma_factor_interrobot=float(g["sigma-factor-interrobot"]),
            sigma_factor_obstacle=float(g["sigma-factor-obstacle"]), sigma_factor_tracking=float(g["sigma-factor-tracking"]),
            safety_distance_multiplier=float(r["inter-robot-safety-distance-multiplier"]),
            comms_radius=float(r["communication"]["radius"]), target_speed=float(r["target-speed"]),
            delta_t=1.0 / self.hz, tracking_switch_padding=float(trk.get("switch-padding", 1.0)),
            tracking_attraction_distance=float(trk.get("attraction-distance", 2.0)),
            enable_dynamic=int(en.get("dynamic", True)), enable_interrobot=int(en.get("interrobot", True)),
            enable_obstacle=int(en.get("obstacle", True)), enable_tracking=int(en.get("tracking", False)),
            schedule_kind=_SCHEDULES[it.get("schedule", "interleave-evenly")], iterations_internal=int(it.get("internal", 10)),
            iterations_external=int(it.get("external", 10)), world_width=self.world_w, world_height=self.world_h)
        self.timesteps = get_variable_timesteps(lookahead_horizon(cfg.target_speed, self.planning_horizon),
                                                self.lookahead_multiple)
        self.cfg = replace(cfg, num_variables=int(self.timesteps.shape[0]))
        crit = {(f.reached_when, f.finished_when) for f in self.formations}
        # gbp_world_reached_waypoint takes one criterion pair per call; the shipped scenarios use one per group
        self.reached_when, self.finished_when = next(iter(crit)) if len(crit) == 1 else (None, None)

    @classmethod
    def from_directory(cls, path: str) -> "ReferenceScenario":
        """Straight from a scenario directory of the reference (`config.toml`, `formation.yaml`, `environment.yaml`)."""
        import os
        import tomllib

        sc = cls(os.path.basename(os.path.normpath(path)), data=read_scenario_directory(path))
        with open(os.path.join(path, "config.toml"), "rb") as f:
            sc.rrt = tomllib.load(f).get("rrt", {})  # `[rrt]`: the global planner's parameters (gbp_config RRTSection)
        return sc

    def draw_antennas(self, n: int, rng) -> np.ndarray:
        """update_failed_comms (robot.rs:1592-1601) for one tick: `antenna.active = !prng.gen_bool(failure_rate)` per
        robot, in robot order — the `antenna_active` argument of `set_comms`.  The draws come from the numpy Generator
        (the reference's WyRand stream is third-party and not reproduced, like its other random inputs)."""
        if self.failure_rate <= 0.0:
            return np.ones(n, np.uint8)
        return (~(rng.random(n) < self.failure_rate)).astype(np.uint8)

    def spawn_events(self, ticks: int) -> list[tuple[int, int]]:
        """(tick, formation index) of every spawn in the first `ticks` fixed steps, in time then formation order."""
        ev = []
        for k, f in enumerate(self.formations):
            for t in f.spawn_times(ticks / self.hz):
                tick = int(np.ceil(t * self.hz - 1e-9))
                if tick < ticks:
                    ev.append((tick, k))
        return sorted(ev)

    def spawn(self, formation_index: int, rng) -> Swarm | None:
        """spawn_formation for one event: radii, Formation::as_positions, the route of every robot."""
        from .formation import as_positions, routes

        f = self.formations[formation_index]
        if f.robots == 0:
            return None  # `robots: 0` (Obstacle Shapes Showcase): the spawn loop has nothing to iterate over
        lo, hi = self.radius_range
        radii = np.asarray([f32(lo) if lo == hi else f32(rng.uniform(lo, hi)) for _ in range(f.robots)], f32)
        placed = as_positions(f, self.world_w, self.world_h, radii, rng)
        if placed is None:
            return None  # "failed to spawn formation ..., skipping" (spawner.rs:460-468)
        init, wps = placed
        r = routes(init, wps)
        return _finish(self.cfg, radii, init, np.stack([w[1] for w in r]).astype(f32), self.timesteps,
                       self.planning_horizon, sdf=None, name=f"{self.name}[{formation_index}]", waypoints=r)
