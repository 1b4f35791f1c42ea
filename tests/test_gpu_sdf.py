"""Device SDF generation (gbp_env_to_sdf_image / gbp_world_set_sdf_from_environment)
against the oracle restatement of env_to_png::env_to_sdf_image: every byte equal."""
import json
import os

import numpy as np
import pytest

from magics_b200 import Environment, GbpConfig, Obstacle, World, env_to_sdf_image
from oracle import oracle as oo
from oracle.oracle import OracleWorld

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "env_to_png.json")))


@pytest.mark.parametrize("name", sorted(GOLD["environments"]))
def test_scenario_environments_bit_exact(name):
    env = Environment(**GOLD["environments"][name])
    got, ref = env_to_sdf_image(env), oo.env_to_sdf_image(env)
    assert got.shape == ref.shape == (*env.image_shape, 3)
    assert np.array_equal(got, ref), f"{name}: {int((got != ref).sum())} bytes differ"
    assert len(np.unique(ref)) > 2  # blurred


ALL_TILES = ["─│╴╶╷╵┌┐", "└┘┬┴├┤┼ ", "█x┼─│┌┘ "]


@pytest.mark.parametrize("res,path_width,expansion,blur", [
    (40, 0.4, 0.0, 0.0), (37, 0.33, 0.07, 0.0), (40, 0.5, 0.1, 0.05), (64, 0.16, 0.01, 0.2), (25, 0.9, 0.3, 0.04),
    (16, 0.5, 0.0, 1.0),   # sigma = 16 px: windows wider than a tile, clamped at every border
    (7, 0.5, 0.5, 0.15),   # path closed by the expansion, sigma barely above one pixel
])
def test_all_tile_kinds_bit_exact(res, path_width, expansion, blur):
    env = Environment(grid=ALL_TILES, tile_size=12.5, path_width=path_width, resolution=res, expansion=expansion,
                      blur=blur)
    got, ref = env_to_sdf_image(env), oo.env_to_sdf_image(env)
    assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {ref.size} bytes differ"


def test_single_pixel_tiles_and_single_row():
    for grid in (["┼"], ["─┼─"], ["│", "┼", "│"]):
        for res in (1, 2, 3):
            env = Environment(grid=grid, resolution=res, path_width=0.5, blur=0.9)
            assert np.array_equal(env_to_sdf_image(env), oo.env_to_sdf_image(env))


def test_world_sdf_from_environment_feeds_the_obstacle_lookup():
    env = Environment(**GOLD["environments"]["Collaborative Complex"])
    ww, wh = env.world_size
    cfg = GbpConfig(world_width=ww, world_height=wh)
    g, o = World(cfg, device=0), OracleWorld(cfg)
    g.set_sdf_from_environment(env)
    o.set_sdf(oo.env_to_sdf_image(env))
    rng = np.random.default_rng(0)
    xy = np.stack([rng.uniform(-ww / 2 - 5, ww / 2 + 5, 20000), rng.uniform(-wh / 2 - 5, wh / 2 + 5, 20000)], axis=1)
    pg, po = g.sdf_lookup(xy), o.sdf_lookup(xy)
    for a, b in zip(pg, po):
        assert np.array_equal(a, b)


def test_bad_environments_are_rejected():
    with pytest.raises(RuntimeError):
        env_to_sdf_image(Environment(grid=["┼"], path_width=0.2, expansion=0.3))  # Percentage::new(-0.1) panics
    with pytest.raises(RuntimeError):
        env_to_sdf_image(Environment(grid=["┼"], blur=1.5))
    with pytest.raises(RuntimeError):
        env_to_sdf_image(Environment(grid=["█"], obstacles=[Obstacle("regular-polygon", sides=0, radius=0.1)]))
    with pytest.raises(ValueError):
        env_to_sdf_image(Environment(grid=["┼─", "┼"]))


def test_junction_scenario_on_the_generated_sdf():
    """Config 2 end to end: the reference environment of `Structured Junction Twoway` rasterised and blurred
    on the device, Obstacle factors reading it, beliefs equal to the oracle's on its own image."""
    from magics_b200 import scenarios
    from tests.parity import assert_beliefs_match

    env = Environment(**GOLD["environments"]["Structured Junction Twoway"])
    sw = scenarios.junction_twoway(per_lane=2)
    assert (sw.cfg.world_width, sw.cfg.world_height) == env.world_size
    g, o = World(sw.cfg, device=0), OracleWorld(sw.cfg)
    g.set_sdf_from_environment(env)
    o.set_sdf(oo.env_to_sdf_image(env))
    sw.add_to(g, set_sdf=False)
    sw.add_to(o, set_sdf=False)
    for _ in range(4):
        g.step()
        o.step()
    assert_beliefs_match(g.read_beliefs(), o.read_beliefs(), what="junction on generated SDF")


@pytest.mark.parametrize("name", sorted(GOLD["environments_with_obstacles"]))
def test_scenario_environments_with_placeable_obstacles_bit_exact(name):
    e = GOLD["environments_with_obstacles"][name]
    env = Environment(**{**e, "obstacles": [Obstacle(**o) for o in e["obstacles"]]})
    got, ref = env_to_sdf_image(env), oo.env_to_sdf_image(env)
    assert np.array_equal(got, ref), f"{name}: {int((got != ref).sum())} bytes differ"


def test_random_placeable_obstacles_bit_exact():
    rng = np.random.default_rng(5)
    kinds = ["circle", "triangle", "regular-polygon", "polygon", "rectangle"]
    for trial in range(6):
        obstacles = []
        for k in range(40):
            kind = kinds[k % 5]
            kw = dict(row=int(rng.integers(0, 2)), col=int(rng.integers(0, 3)),
                      translation=(float(rng.uniform(0, 1)), float(rng.uniform(0, 1))),
                      rotation=float(rng.uniform(0, 2 * np.pi)))
            if kind == "circle":
                kw["radius"] = float(rng.uniform(0.01, 0.2))
            elif kind == "triangle":
                a = float(rng.uniform(0.3, 1.3))
                kw.update(radius=float(rng.uniform(0.01, 0.08)), angles=(a, float(rng.uniform(0.3, np.pi - a - 0.3))))
            elif kind == "regular-polygon":
                kw.update(sides=int(rng.integers(3, 9)), radius=float(rng.uniform(0.05, 0.4)))
            elif kind == "polygon":
                m = int(rng.integers(3, 8))
                ang = np.sort(rng.uniform(0, 2 * np.pi, m))
                rad = rng.uniform(0.05, 0.3, m)
                kw["points"] = tuple((float(r * np.cos(a)), float(r * np.sin(a))) for r, a in zip(rad, ang))
            else:
                kw.update(width=float(rng.uniform(0.05, 0.8)), height=float(rng.uniform(0.05, 0.8)))
            obstacles.append(Obstacle(kind, **kw))
        env = Environment(grid=["█─┼", "│█ "], tile_size=20.0, path_width=0.3, resolution=int(rng.integers(30, 90)),
                          expansion=float(rng.choice([0.0, 0.025, 0.1])), blur=float(rng.choice([0.0, 0.03])),
                          obstacles=obstacles)
        got, ref = env_to_sdf_image(env), oo.env_to_sdf_image(env)
        assert np.array_equal(got, ref), f"trial {trial}: {int((got != ref).sum())} bytes differ"
