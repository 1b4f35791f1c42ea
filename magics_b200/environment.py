"""Host-side mirror of `gbp_environment::Environment` as far as SDF generation reads it
(crates/gbp_environment/src/lib.rs:40-75,940-971) and of `env_to_png::env_to_sdf_image`
(crates/env_to_png/src/lib.rs:149-163), which the reference calls at scenario load
(crates/magics/src/simulation_loader.rs:153-161).  The rasteriser and the blur run on the device
(magics_b200/csrc/gbp_sdf.cuh) behind `gbp_env_to_sdf_image` / `gbp_world_set_sdf_from_environment`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


SHAPE_KINDS = {"circle": 0, "triangle": 1, "regular-polygon": 2, "polygon": 3, "rectangle": 4}


class CObstacle(C.Structure):
    """include/gbp_b200.h gbp_obstacle_t"""
    _fields_ = [("kind", C.c_int32), ("tile_row", C.c_int32), ("tile_col", C.c_int32), ("rotation", C.c_double),
                ("tx", C.c_double), ("ty", C.c_double), ("radius", C.c_double), ("angle_a", C.c_double),
                ("angle_b", C.c_double), ("sides", C.c_int32), ("n_points", C.c_int32), ("point_offset", C.c_int64),
                ("width", C.c_double), ("height", C.c_double)]


class CEnvironment(C.Structure):
    """include/gbp_b200.h gbp_environment_t"""
    _fields_ = [("nrows", C.c_int32), ("ncols", C.c_int32), ("tiles", C.POINTER(C.c_uint32)),
                ("tile_size", C.c_float), ("path_width", C.c_float), ("resolution", C.c_uint32),
                ("expansion", C.c_float), ("blur", C.c_float), ("n_obstacles", C.c_int32),
                ("obstacles", C.POINTER(CObstacle)), ("polygon_points", C.POINTER(C.c_double))]


@dataclass
class Obstacle:
    """`gbp_environment::Obstacle` (crates/gbp_environment/src/lib.rs:532-572); angles in radians as in the YAML."""
    shape: str                     # one of SHAPE_KINDS
    row: int = 0
    col: int = 0
    translation: tuple = (0.5, 0.5)
    rotation: float = 0.0
    radius: float = 0.0            # circle, triangle, regular-polygon
    angles: tuple = (0.0, 0.0)     # triangle (A, B)
    sides: int = 0                 # regular-polygon
    points: tuple = ()             # polygon ((x, y), ...)
    width: float = 0.0             # rectangle
    height: float = 0.0


@dataclass
class Environment:
    """tiles.grid + tiles.settings (+ settings.sdf) of an `environment.yaml`."""
    grid: list[str]
    tile_size: float = 100.0
    path_width: float = 0.1
    resolution: int = 200
    expansion: float = 0.0
    blur: float = 0.0
    obstacles: list = field(default_factory=list)

    @classmethod
    def from_yaml(cls, text: str) -> "Environment":
        """Parses an `environment.yaml` (serde's externally tagged shapes: `shape: !circle {radius: ..}`)."""
        import yaml

        class Loader(yaml.SafeLoader):
            pass

        def tagged(loader, suffix, node):
            return {"kind": suffix, **loader.construct_mapping(node, deep=True)}

        Loader.add_multi_constructor("!", tagged)
        d = yaml.load(text, Loader=Loader)
        s = d["tiles"]["settings"]
        sdf = s.get("sdf", {})
        obstacles = []
        for o in d.get("obstacles") or []:
            sh = o["shape"]
            ang = sh.get("angles", {})
            obstacles.append(Obstacle(
                shape=sh["kind"], row=int(o["tile-coordinates"]["row"]), col=int(o["tile-coordinates"]["col"]),
                translation=(float(o["translation"]["x"]), float(o["translation"]["y"])), rotation=float(o["rotation"]),
                radius=float(sh.get("radius", 0.0)), angles=(float(ang.get("A", 0.0)), float(ang.get("B", 0.0))),
                sides=int(sh.get("sides", 0)), points=tuple((float(p["x"]), float(p["y"])) for p in sh.get("points", [])),
                width=float(sh.get("width", 0.0)), height=float(sh.get("height", 0.0))))
        return cls(grid=list(d["tiles"]["grid"]), tile_size=float(s["tile-size"]), path_width=float(s["path-width"]),
                   resolution=int(sdf.get("resolution", 200)), expansion=float(sdf.get("expansion", 0.0)),
                   blur=float(sdf.get("blur", 0.0)), obstacles=obstacles)

    @property
    def nrows(self) -> int:
        return len(self.grid)

    @property
    def ncols(self) -> int:
        return len(self.grid[0])  # TileGrid::ncols: chars().count() of the first row

    @property
    def image_shape(self) -> tuple[int, int]:
        return self.nrows * self.resolution, self.ncols * self.resolution

    @property
    def world_size(self) -> tuple[float, float]:
        """obstacle::WorldSize (robot.rs:1258-1263): (width, height) = tile_size * (ncols, nrows)."""
        return float(np.float32(self.tile_size)) * self.ncols, float(np.float32(self.tile_size)) * self.nrows

    def tile_codes(self) -> np.ndarray:
        """Unicode code points, row-major; short rows read as the reference's `get_tile` -> None would:
        they are rejected here instead of panicking inside the rasteriser."""
        rows = [[ord(ch) for ch in r] for r in self.grid]
        if any(len(r) != self.ncols for r in rows):
            raise ValueError("tile grid rows differ in length (env_to_image: 'Tile not found')")
        return np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1)

    def c_obstacles(self):
        """(CObstacle array, polygon points (m, 2) f64)"""
        arr = (CObstacle * max(1, len(self.obstacles)))()
        pts: list[tuple[float, float]] = []
        for k, o in enumerate(self.obstacles):
            if isinstance(o, dict):
                o = Obstacle(**o)
            arr[k] = CObstacle(SHAPE_KINDS[o.shape], o.row, o.col, o.rotation, o.translation[0], o.translation[1],
                               o.radius, o.angles[0], o.angles[1], o.sides, len(o.points), len(pts), o.width, o.height)
            pts.extend(o.points)
        return arr, np.ascontiguousarray(pts if pts else [(0.0, 0.0)], dtype=np.float64)

    def c_struct(self):
        """(CEnvironment, keep-alive objects)"""
        codes = self.tile_codes()
        obs, pts = self.c_obstacles()
        ce = CEnvironment(self.nrows, self.ncols, codes.ctypes.data_as(C.POINTER(C.c_uint32)), self.tile_size,
                          self.path_width, self.resolution, self.expansion, self.blur, len(self.obstacles),
                          C.cast(obs, C.POINTER(CObstacle)), pts.ctypes.data_as(C.POINTER(C.c_double)))
        return ce, (codes, obs, pts)


class CCollider(C.Structure):
    """include/gbp_b200.h gbp_collider_t"""
    _fields_ = [("kind", C.c_int32), ("translation", C.c_float * 2), ("angle", C.c_float), ("radius", C.c_float),
                ("half_extents", C.c_float * 2), ("first_vertex", C.c_int32), ("num_vertices", C.c_int32)]


@dataclass
class Collider:
    """One entry of `gbp_global_planner::Colliders`: a parry2d shape in its own frame + Isometry2 (translation, angle).
    kind: "ball" (radius), "cuboid" (half_extents), "triangle" / "convex-polygon" (points, counter-clockwise)."""
    kind: str
    translation: tuple = (0.0, 0.0)
    angle: float = 0.0
    radius: float = 0.0
    half_extents: tuple = (0.0, 0.0)
    points: tuple = ()


COLLIDER_KINDS = {"ball": 0, "cuboid": 1, "triangle": 2, "convex-polygon": 3}


def pack_colliders(colliders):
    """(CCollider array, vertices (m, 2) f32, plain (n, 9) f32 rows for the oracle)"""
    arr = (CCollider * max(1, len(colliders)))()
    verts: list[tuple[float, float]] = []
    rows = np.zeros((max(1, len(colliders)), 9), np.float32)
    for k, c in enumerate(colliders):
        kind = COLLIDER_KINDS[c.kind]
        arr[k] = CCollider(kind, (C.c_float * 2)(*c.translation), c.angle, c.radius, (C.c_float * 2)(*c.half_extents),
                           len(verts), len(c.points))
        rows[k] = (kind, c.translation[0], c.translation[1], c.angle, c.radius, c.half_extents[0], c.half_extents[1],
                   len(verts), len(c.points))
        verts.extend(c.points)
    return arr, np.ascontiguousarray(verts if verts else [(0.0, 0.0)], dtype=np.float32), rows


# Tile -> cuboids of `build_tile_grid` (crates/magics/src/environment/map_generator.rs:537-1298): per tile character a
# list of (x length, z length, x offset, z offset) in units of (T = tile_size, B = base_dim, P = pos_offset,
# W = path_width * tile_size) — dims as ("T" | "B" | "T/2" | "W"), offsets as signed multiples of P or T/4.
_TILE_CUBOIDS = {
    "─": [("T", "B", 0, "-P"), ("T", "B", 0, "+P")],
    "│": [("B", "T", "-P", 0), ("B", "T", "+P", 0)],
    "╴": [("T", "B", 0, "-P"), ("T", "B", 0, "+P"), ("T/2", "W", "+Q", 0)],
    "╶": [("T", "B", 0, "-P"), ("T", "B", 0, "+P"), ("T/2", "W", "-Q", 0)],
    "╷": [("B", "T", "-P", 0), ("B", "T", "+P", 0), ("W", "T/2", 0, "+Q")],
    "╵": [("B", "T", "-P", 0), ("B", "T", "+P", 0), ("W", "T/2", 0, "-Q")],
    "┌": [("B", "B", "+P", "-P"), ("B", "T", "-P", 0), ("T", "B", 0, "+P")],
    "┐": [("B", "B", "-P", "-P"), ("B", "T", "+P", 0), ("T", "B", 0, "+P")],
    "└": [("B", "B", "+P", "+P"), ("B", "T", "-P", 0), ("T", "B", 0, "-P")],
    "┘": [("B", "B", "-P", "+P"), ("B", "T", "+P", 0), ("T", "B", 0, "-P")],
    "┬": [("B", "B", "-P", "-P"), ("B", "B", "+P", "-P"), ("T", "B", 0, "+P")],
    "┴": [("B", "B", "-P", "+P"), ("B", "B", "+P", "+P"), ("T", "B", 0, "-P")],
    "├": [("B", "B", "+P", "-P"), ("B", "B", "+P", "+P"), ("B", "T", "-P", 0)],
    "┤": [("B", "B", "-P", "-P"), ("B", "B", "-P", "+P"), ("B", "T", "+P", 0)],
    "┼": [("B", "B", "-P", "-P"), ("B", "B", "+P", "-P"), ("B", "B", "-P", "+P"), ("B", "B", "+P", "+P")],
    " ": [("T", "T", 0, 0)],
}
_TILE_CUBOIDS["-"] = _TILE_CUBOIDS["─"]
_TILE_CUBOIDS["|"] = _TILE_CUBOIDS["│"]


def tile_colliders(env: "Environment") -> list:
    """The cuboid colliders `build_tile_grid` pushes for the tile grid (map_generator.rs:537-1298), f32 arithmetic in
    the reference's order.  A Bevy `Cuboid::new(x, h, z)` becomes a parry2d Cuboid through the fork's Bevy conversion
    (not in the reference tree): half extents (x / 2, z / 2) are assumed."""
    f = np.float32
    T, pw = f(env.tile_size), f(env.path_width)
    B = T * (f(1.0) - pw) / f(2.0)
    P = (pw * T + B) / f(2.0)  # path_width.mul_add(tile_size, base_dim) / 2.0 (fused in the reference; equal here
    #                             whenever path_width * tile_size is exact in f32)
    dims = {"T": T, "B": B, "T/2": T / f(2.0), "W": pw * T}
    offs = {0: f(0.0), "+P": P, "-P": -P, "+Q": T / f(4.0), "-Q": -(T / f(4.0))}
    gx = f(env.ncols) / f(2.0) - f(0.5)
    gz = -(f(env.nrows) / f(2.0) - f(0.5))
    out = []
    for y, row in enumerate(env.grid):
        for x, ch in enumerate(row):
            rules = _TILE_CUBOIDS.get(ch)
            if not rules:
                continue
            ox = (f(x) - gx) * T
            oz = (-f(y) - gz) * T
            for dx, dz, sx, sz in rules:
                out.append(Collider("cuboid", (float(ox + offs[sx]), float(oz + offs[sz])), 0.0,
                                    half_extents=(float(dims[dx] / f(2.0)), float(dims[dz] / f(2.0)))))
    return out


def _fma32(a, b, c):
    """f32 mul_add: the product of two f32 is exact in f64; one rounding after the sum."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def _convex_hull_ccw(pts):
    """Counter-clockwise convex hull without collinear points (what parry2d's ConvexPolygon::from_convex_hull keeps)."""
    p = sorted({(float(x), float(y)) for x, y in pts})
    if len(p) < 3:
        return p

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for q in p:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], q) <= 0:
            lower.pop()
        lower.append(q)
    for q in reversed(p):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], q) <= 0:
            upper.pop()
        upper.append(q)
    return lower[:-1] + upper[:-1]


def obstacle_colliders(env: "Environment") -> list:
    """The colliders `build_obstacles` pushes for the placeable obstacles (crates/magics/src/environment/
    map_generator.rs:141-536), f32 arithmetic in the reference's order: Circle -> Ball, Triangle -> Triangle (vertices
    rotated by Quat::from_rotation_y(pi/2 - rotation), isometry angle -rotation), RegularPolygon / Polygon ->
    ConvexPolygon::from_convex_hull, Rectangle -> Cuboid.  As written in the reference, the circle's and the polygon's
    vertical placement is not mirrored like the other shapes' (`1 - y` resp. no sign flip, :186 / :389): kept.
    glam's quaternion rotations are restated as the plane rotations they are (the last bits may differ)."""
    f = np.float32
    T = f(env.tile_size)
    gox, goz = f(env.ncols) / f(2.0) - f(0.5), f(env.nrows) / f(2.0) - f(0.5)
    half = T / f(2.0)
    pi = f(np.pi)
    out = []
    for o in env.obstacles:
        if isinstance(o, dict):
            o = Obstacle(**o)
        ox, oz = (f(o.col) - gox) * T, (f(o.row) - goz) * T
        tx, ty = f(o.translation[0]), f(o.translation[1])
        x = _fma32(tx, T, ox) - half
        z_plain = _fma32(ty, T, oz) - half
        rot = f(o.rotation)
        if o.shape == "circle":
            z = _fma32(f(1.0) - ty, T, oz) - half
            out.append(Collider("ball", (float(x), float(z)), 0.0, radius=float(f(o.radius) * T)))
        elif o.shape == "triangle":
            a, b = f(o.angles[0]), f(o.angles[1])
            c = pi - (a + b)
            r = f(o.radius)
            hyp = [r / np.sin(a), r / np.sin(b), r / np.sin(c)]
            ang = [pi + a / f(2.0), -b / f(2.0), pi - b - c / f(2.0)]
            pts = [(f(np.cos(t)) * h, f(np.sin(t)) * h) for t, h in zip(ang, hyp)]  # Triangle::points
            pts = [(-px * T, py * T) for px, py in pts]
            th = pi / f(2.0) - rot  # Quat::from_rotation_y(th) on (px, 0, py): x' = x cos + z sin, z' = -x sin + z cos
            ct, st = f(np.cos(th)), f(np.sin(th))
            rp = tuple((float(px * ct + py * st), float(-px * st + py * ct)) for px, py in pts)
            out.append(Collider("triangle", (float(x), float(-z_plain)), float(th - pi / f(2.0)), points=rp))
        elif o.shape == "regular-polygon":
            n = int(o.sides)
            off = pi + (f(0.0) if n == 4 else (pi / f(2.0) if n % 2 else -pi / f(2.0)))
            ang = rot + off
            ca, sa = f(np.cos(ang)), f(np.sin(ang))
            scale = T / f(2.0)
            pts = []
            for i in range(n):  # RegularPolygon::point_at in f64, then `as f32`
                t = 2.0 * np.pi / n * i + np.pi / 4.0
                px, py = f(np.cos(t) * float(o.radius)), f(np.sin(t) * float(o.radius))
                pts.append(((px * ca - py * sa) * scale, (px * sa + py * ca) * scale))  # Quat::from_rotation_z(ang)
            hull = tuple(_convex_hull_ccw(pts))
            out.append(Collider("convex-polygon", (float(x), float(-z_plain)), float(ang), points=hull))
        elif o.shape == "polygon":
            pts = [(f(px) * T, f(py) * T) for px, py in o.points]
            out.append(Collider("convex-polygon", (float(x), float(z_plain)), 0.0, points=tuple(_convex_hull_ccw(pts))))
        elif o.shape == "rectangle":
            out.append(Collider("cuboid", (float(x), float(-z_plain)), 0.0,
                                half_extents=(float(f(o.width) * T / f(4.0)), float(f(o.height) * T / f(4.0)))))
        else:
            raise ValueError(f"unknown placeable shape {o.shape!r}")
    return out


def colliders(env: "Environment") -> list:
    """Everything `Colliders` holds after the map has been generated: tile walls, then placeable obstacles."""
    return tile_colliders(env) + obstacle_colliders(env)
