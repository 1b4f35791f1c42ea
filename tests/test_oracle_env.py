"""Oracle restatement of env_to_png::env_to_sdf_image (crates/env_to_png/src/lib.rs) against the crate's
own #[test]s (tests/golden/env_to_png.json, extracted by tests/golden/make_golden.py) and against
independent properties of the rasteriser and of the Gaussian blur (image 0.25.1, third party: parity of
the blur itself is unpinned, see oracle/gbp_oracle.cpp)."""
import ctypes as C
import json
import os

import numpy as np

from magics_b200.environment import Environment
from oracle import oracle as oo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "env_to_png.json")))
f32 = np.float32


def _env(name, **over):
    e = dict(GOLD["environments"][name])
    e.update(over)
    return Environment(**e)


def test_reference_kats():
    L = oo.lib()
    for k in GOLD["kats"]:
        a = k["args"]
        if k["fn"] == "image_to_tile_units":
            o = (C.c_float * 2)()
            L.gbpo_image_to_tile_units(a["px"], a["py"], a["resolution"], C.c_float(a["tile_size"]), o)
            # stale test: the function adds half a pixel since lib.rs:218
            want = [f32((f32(a["px"]) + f32(0.5)) / f32(a["resolution"]) * f32(a["tile_size"])),
                    f32((f32(a["py"]) + f32(0.5)) / f32(a["resolution"]) * f32(a["tile_size"]))]
            assert not k["consistent_with_code"] and [o[0], o[1]] == [float(want[0]), float(want[1])]
        elif k["fn"] == "tile_units_to_percentage":
            o = (C.c_float * 2)()
            L.gbpo_tile_units_to_percentage(C.c_float(a["x"]), C.c_float(a["y"]), C.c_float(a["tile_size"]), o)
            assert not k["consistent_with_code"]
            assert abs(o[0] - 0.23) < 1e-6 and abs(o[1] - 0.56) < 1e-6  # fraction into the tile
        elif k["fn"] == "image_to_tile_coords":
            o = (C.c_uint64 * 2)()
            L.gbpo_image_to_tile_coords(a["px"], a["py"], a["resolution"], o)
            assert k["consistent_with_code"] and [o[0], o[1]] == k["expected"]
        elif k["fn"] == "is_tile_obstacle":
            got = bool(L.gbpo_is_tile_obstacle(ord(a["tile"]), C.c_float(a["path_width"]), C.c_float(a["x"]),
                                               C.c_float(a["y"]), C.c_float(a["expansion"])))
            assert got == (k["expected"] if k["consistent_with_code"] else not k["expected"])


def test_cross_tile_without_blur_is_the_plus_shaped_road():
    env = _env("Structured Junction Twoway", blur=0.0)
    img = oo.env_to_sdf_image(env)
    assert img.shape == (200, 200, 3) and set(np.unique(img)) == {0, 255}
    assert np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 0], img[..., 2])
    g = img[..., 0]
    assert np.array_equal(g, g.T) and np.array_equal(g, g[::-1, :]) and np.array_equal(g, g[:, ::-1])
    # path_width - expansion = 0.15 of the tile is free around the centre lines: 30 of 200 pixels
    assert int((g[0] == 255).sum()) == 30 and int((g[100] == 255).sum()) == 200
    assert g[0, 0] == 0 and g[100, 0] == 255 and g[0, 100] == 255


def test_every_tile_kind_matches_its_connectivity():
    # a free pixel at the tile centre and at the middle of each edge the glyph connects to
    conn = {"─": "lr", "│": "ud", "╴": "l", "╶": "r", "╷": "d", "╵": "u", "┌": "dr", "┐": "dl", "└": "ur", "┘": "ul",
            "┬": "ldr", "┴": "lur", "├": "udr", "┤": "udl", "┼": "udlr", " ": ""}
    res = 40
    for ch, c in conn.items():
        g = oo.env_to_sdf_image(Environment(grid=[ch], tile_size=10.0, path_width=0.4, resolution=res))[..., 0]
        edge = {"l": g[res // 2, 0], "r": g[res // 2, res - 1], "u": g[0, res // 2], "d": g[res - 1, res // 2]}
        for k, v in edge.items():
            assert (v == 255) == (k in c), (ch, k)
        # the centre is road (stubs end exactly there: one of the four centre pixels is free)
        assert (g[res // 2 - 1:res // 2 + 1, res // 2 - 1:res // 2 + 1].max() == 255) == (ch != " ")
        assert g[0, 0] == 0 or ch not in conn  # corners are never road
    # unknown glyphs (e.g. the full block of the circle scenario) are free space
    assert np.all(oo.env_to_sdf_image(Environment(grid=["█"], resolution=20)) == 255)


def test_blur_matches_a_gaussian_filter_in_the_interior():
    from scipy.ndimage import gaussian_filter

    for name in ("Structured Junction Twoway", "Collaborative Complex"):
        env = _env(name, resolution=50 if name == "Collaborative Complex" else 200)
        if name == "Collaborative Complex":
            env.blur = 0.08  # sigma 4 px at 50 px per tile
        sharp = oo.env_to_sdf_image(Environment(**{**env.__dict__, "blur": 0.0}))[..., 0].astype(np.float64)
        got = oo.env_to_sdf_image(env)[..., 0].astype(np.int64)
        sigma = float(f32(env.blur) * f32(env.resolution))
        ref = np.rint(gaussian_filter(sharp, sigma, truncate=2.0, mode="nearest"))
        m = int(2 * sigma) + 2
        assert np.max(np.abs(got[m:-m, m:-m] - ref[m:-m, m:-m])) <= 1
        assert got.min() >= 0 and got.max() <= 255
    # a constant image stays constant, borders included (window weights are renormalised)
    assert np.all(oo.env_to_sdf_image(Environment(grid=["██", "██"], resolution=30, blur=0.1)) == 255)


def test_blur_below_one_pixel_is_skipped():
    env = Environment(grid=["┼"], resolution=50, path_width=0.3, blur=0.019)  # 0.95 px
    assert set(np.unique(oo.env_to_sdf_image(env))) == {0, 255}


# ---- placeable obstacles (is_placeable_obstacle, env_to_png/src/lib.rs:283-336) -----------------------------
def _blank(res=400, **kw):
    return Environment(grid=["█"], resolution=res, **kw)  # unknown glyph: no tile obstacle anywhere


def _dark_fraction(env):
    return float((oo.env_to_sdf_image(env)[..., 0] == 0).mean())


def test_obstacle_environments_parse_and_render():
    from magics_b200.environment import Obstacle

    for name, e in GOLD["environments_with_obstacles"].items():
        env = Environment(**{**e, "obstacles": [Obstacle(**o) for o in e["obstacles"]]})
        img = oo.env_to_sdf_image(env)
        bare = oo.env_to_sdf_image(Environment(**{**e, "obstacles": []}))
        assert img.shape == bare.shape and (img.astype(int) <= bare.astype(int)).all(), name  # obstacles only darken
        assert (img != bare).any(), name


def test_shape_areas_match_their_geometry():
    from magics_b200.environment import Obstacle

    e = 0.02
    # circle: radius + expansion (gbp_environment/src/lib.rs:128-141)
    f = _dark_fraction(_blank(expansion=e, obstacles=[Obstacle("circle", radius=0.2)]))
    assert abs(f - np.pi * (0.2 + e) ** 2) < 2e-3
    # rectangle: the test is against width / 4 and height / 4 after + 2 * expansion (:335-359)
    f = _dark_fraction(_blank(expansion=e, obstacles=[Obstacle("rectangle", width=0.8, height=0.4, rotation=0.3)]))
    assert abs(f - ((0.8 + 2 * e) / 2) * ((0.4 + 2 * e) / 2)) < 2e-3
    # regular polygon: the point is doubled, so the circumradius is (radius + 2 * expansion) / 2 (:250-313)
    for n in (3, 4, 5, 8):
        f = _dark_fraction(_blank(expansion=e, obstacles=[Obstacle("regular-polygon", sides=n, radius=0.5, rotation=1.0)]))
        rho = (0.5 + 2 * e) / 2
        assert abs(f - n / 2 * rho * rho * np.sin(2 * np.pi / n)) < 2e-3, n
    # polygon: vertices pushed away from their mean by 4 * expansion of the offset (:374-401), shoelace area
    pts = np.array([(-0.3, -0.2), (0.25, -0.25), (0.3, 0.1), (0.0, 0.3), (-0.35, 0.15)])
    grown = pts + (pts - pts.mean(0)) * 4 * e
    area = 0.5 * abs(np.dot(grown[:, 0], np.roll(grown[:, 1], -1)) - np.dot(grown[:, 1], np.roll(grown[:, 0], -1)))
    f = _dark_fraction(_blank(expansion=e, obstacles=[Obstacle("polygon", points=tuple(map(tuple, pts)), rotation=0.7)]))
    assert abs(f - area) < 2e-3
    # triangle: vertices at radius / sin(angle) along the directions of Triangle::points (:192-210; not the
    # incircle construction its doc comment promises), area by the shoelace formula in float64
    A, B = 1.0, 0.9
    Cc = np.pi - (A + B)
    r = 0.1 + e
    dirs = (np.pi + A / 2, -B / 2, np.pi - B - Cc / 2)
    v = np.array([[np.cos(d) * r / np.sin(a), np.sin(d) * r / np.sin(a)] for d, a in zip(dirs, (A, B, Cc))])
    area = 0.5 * abs(np.dot(v[:, 0], np.roll(v[:, 1], -1)) - np.dot(v[:, 1], np.roll(v[:, 0], -1)))
    f = _dark_fraction(_blank(expansion=e, obstacles=[Obstacle("triangle", radius=0.1, angles=(A, B), rotation=2.0)]))
    assert abs(f - area) < 2e-3


def test_obstacles_only_apply_to_their_tile_and_first_hit_wins():
    from magics_b200.environment import Obstacle

    env = Environment(grid=["██", "██"], resolution=50,
                      obstacles=[Obstacle("circle", row=1, col=0, radius=0.3), Obstacle("circle", row=0, col=1, radius=0.1)])
    g = oo.env_to_sdf_image(env)[..., 0]
    assert g[:50, :50].min() == 255 and g[50:, 50:].min() == 255  # tiles (0,0) and (1,1) untouched
    assert g[75, 25] == 0 and g[25, 75] == 0                      # centres of tiles (1,0) and (0,1)
    assert (g[50:, :50] == 0).sum() > (g[:50, 50:] == 0).sum()
