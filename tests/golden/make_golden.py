"""Extracts the reference's own known-answer tests for the GBP hot path into JSON
fixtures.  Run in the build container (needs /root/reference); the fixtures are
committed so the GPU box never reads the reference.

  python tests/golden/make_golden.py

Sources (all `#[cfg(test)]` blocks of the reference):
  crates/gbp_schedule/src/schedules/*.rs      -> schedules.json
  crates/magics/src/utils.rs:95-133           -> variable_timesteps.json
  crates/magics/src/factorgraph/factor/marginalise_factor_distance.rs:182-233
                                              -> marginalise.json
  crates/env_to_png/src/lib.rs:482-533 (tests) and config/scenarios/*/environment.yaml
                                              -> env_to_png.json
  config/scenarios/{Circle Experiment, Structured Junction Twoway, Collaborative Complex}/
      {formation.yaml, environment.yaml, config.toml}   (BASELINE configs 1-3: the scenario INPUT data)
                                              -> scenarios.json
"""
import json
import os
import re

REF = "/root/reference/crates"
HERE = os.path.dirname(os.path.abspath(__file__))

KIND = {"centered": 0, "interleave_evenly": 1, "soon_as_possible": 2, "late_as_possible": 3,
        "half_beginning_half_end": 4}


def strip_comments(src: str) -> str:
    return "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("//"))


def schedules():
    out = []
    for name, kind in KIND.items():
        path = f"{REF}/gbp_schedule/src/schedules/{name}.rs"
        src = strip_comments(open(path).read())
        tests = src.split("#[test]")[1:]
        for t in tests:
            fn = re.search(r"fn (\w+)\(", t).group(1)
            # a test may hold several `let config = ...; let mut schedule = ...;` groups
            groups = re.split(r"let config = GbpScheduleParams\s*\{", t)[1:]
            for g in groups:
                m = re.search(r"internal:\s*(\d+),\s*external:\s*(\d+)", g)
                if not m:
                    continue
                seq = re.findall(r"assert_eq!\(schedule\.next\(\),\s*Some\(ts\((true|false),\s*(true|false)\)\)\)", g)
                ends = "assert_eq!(schedule.next(), None)" in g
                if not seq and not ends:
                    continue
                out.append({"kind": kind, "schedule": name, "test": fn, "file": os.path.relpath(path, "/root/reference"),
                            "internal": int(m.group(1)), "external": int(m.group(2)),
                            "sequence": [[a == "true", b == "true"] for a, b in seq], "exhaustive": ends})
    return out


def timesteps():
    src = open(f"{REF}/magics/src/utils.rs").read()
    body = src[src.index("fn test_get_variable_timesteps"):]
    hs = [int(x) for x in re.findall(r"let lookahead_horizon = (\d+);", body)]
    ms = [int(x) for x in re.findall(r"let lookahead_multiple = (\d+);", body)]
    vs = [[int(y) for y in v.split(",") if y.strip()] for v in re.findall(r"vec!\[([\d,\s]+)\]", body)]
    assert len(hs) == len(ms) == len(vs) == 5
    return [{"lookahead_horizon": h, "lookahead_multiple": m, "expected": v} for h, m, v in zip(hs, ms, vs)]


def marginalise():
    # marginalise_factor_distance.rs:140-233: the 8x8 matrix 1..64 and the single-neighbour
    # pass-through case (information vector / precision of a 4-dof factor are returned unchanged, mean = 0)
    # the reference lays the matrix out as four 4x4 quadrants 1..16, 17..32, 33..48, 49..64
    quad = lambda k: [[float(16 * k + 4 * r + c + 1) for c in range(4)] for r in range(4)]
    ul, ur, ll, lr = quad(0), quad(1), quad(2), quad(3)
    full = [ul[r] + ur[r] for r in range(4)] + [ll[r] + lr[r] for r in range(4)]
    return {
        "matrix_8x8": full,
        "blocks_marg_idx_0": {"aa": ul, "ab": ur, "ba": ll, "bb": lr},
        "blocks_marg_idx_4": {"aa": lr, "ab": ll, "ba": ur, "bb": ul},
        # information_vector_length_equal_to_ndofs_do_nothing (:212-233)
        "single_neighbour": {"eta": [0.0, 1.0, 2.0, 3.0],
                             "lam": [[5.0, 0.2, 0.0, 0.0], [0.2, 5.0, 0.0, 0.0], [0.0, 0.0, 5.0, 0.3],
                                     [0.0, 0.0, 0.3, 5.0]]},
    }


def env_to_png():
    # crates/env_to_png/src/lib.rs:482-533: four #[test]s.  Read against the crate's own code, three of
    # their assertions are stale (image_to_tile_units adds 0.5 to the pixel index since :218;
    # tile_units_to_percentage returns the fraction INTO the tile, 0.23 not 0.3; a '─' tile only looks
    # at y, so (0.1, 0.6) is free): they are recorded as written, flagged `consistent_with_code`.
    src = open(f"{REF}/env_to_png/src/lib.rs").read()
    assert "fn test_image_to_tile_coords" in src and "assert_eq!(tile_coords.x, 1);" in src
    kats = [
        {"fn": "image_to_tile_units", "args": {"px": 23, "py": 56, "resolution": 100, "tile_size": 10.0},
         "expected": [2.3, 5.6], "consistent_with_code": False, "code_gives": "(px + 0.5) / res * tile = [2.35, 5.65]"},
        {"fn": "tile_units_to_percentage", "args": {"x": 2.3, "y": 5.6, "tile_size": 10.0},
         "expected": [0.3, 0.6], "consistent_with_code": False, "code_gives": "offset_modulus = [0.23, 0.56]"},
        {"fn": "image_to_tile_coords", "args": {"px": 134, "py": 240, "resolution": 100}, "expected": [1, 2],
         "consistent_with_code": True},
        {"fn": "is_tile_obstacle", "args": {"tile": "─", "path_width": 0.5, "x": 0.3, "y": 0.6, "expansion": 0.0},
         "expected": False, "consistent_with_code": True},
        {"fn": "is_tile_obstacle", "args": {"tile": "─", "path_width": 0.5, "x": 0.1, "y": 0.6, "expansion": 0.0},
         "expected": True, "consistent_with_code": False, "code_gives": "False: a horizontal tile only tests y"},
    ]
    # the tile-only environments of the BASELINE scenarios (obstacles: [])
    import yaml

    envs = {}
    for name in ("Structured Junction Twoway", "Collaborative Complex", "Junction Twoway", "Structured Junction"):
        path = f"/root/reference/config/scenarios/{name}/environment.yaml"
        d = yaml.safe_load(open(path))
        if d.get("obstacles"):
            continue
        s = d["tiles"]["settings"]
        envs[name] = {"grid": d["tiles"]["grid"], "tile_size": s["tile-size"], "path_width": s["path-width"],
                      "resolution": s["sdf"]["resolution"], "expansion": s["sdf"]["expansion"], "blur": s["sdf"]["blur"]}
    # environments with placeable obstacles (all five PlaceableShape variants occur), parsed with the
    # product's own YAML reader and stored as plain data
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from dataclasses import asdict

    from magics_b200.environment import Environment

    placed = {}
    for name in ("Obstacle Shapes Showcase", "Merge", "Environment Obstacles Experiment",
                 "Communications Failure Experiment", "Varying Network Connectivity Experiment"):
        env = Environment.from_yaml(open(f"/root/reference/config/scenarios/{name}/environment.yaml").read())
        placed[name] = asdict(env)
    return {"kats": kats, "environments": envs, "environments_with_obstacles": placed}


def scenarios():
    """The input files of the three BASELINE scenarios, as plain data: the formation group (parsed YAML with serde's
    tags kept as {"kind": tag, ...}), the environment (same form as env_to_png.json) and the scalars of config.toml
    the hot path reads ([gbp], [robot], [simulation].hz / despawn flag)."""
    import sys

    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from magics_b200.scenarios import read_scenario_directory

    out = {}
    for name in ("Circle Experiment", "Structured Junction Twoway", "Collaborative Complex"):
        out[name] = read_scenario_directory(f"/root/reference/config/scenarios/{name}")
    return out


if __name__ == "__main__":
    json.dump(scenarios(), open(os.path.join(HERE, "scenarios.json"), "w"), indent=1, ensure_ascii=False)
    json.dump(env_to_png(), open(os.path.join(HERE, "env_to_png.json"), "w"), indent=1, ensure_ascii=False)
    s = schedules()
    json.dump(s, open(os.path.join(HERE, "schedules.json"), "w"), indent=1)
    json.dump(timesteps(), open(os.path.join(HERE, "variable_timesteps.json"), "w"), indent=1)
    json.dump(marginalise(), open(os.path.join(HERE, "marginalise.json"), "w"), indent=1)
    print(len(s), "schedule cases")
