"""The C-ABI library loads and exports every symbol include/gbp_b200.h declares;
host-only entry points (schedule, timesteps) match the reference's own tests.
No compute call is made here (CPU suite)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import magics_b200
from magics_b200 import GbpConfig
from magics_b200.config import CConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gbp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gbp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = magics_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gbp_b200.h but not exported"


def test_integration_listing_names_every_function_of_the_header():
    """INTEGRATION.md section 1 is the `-sys` crate a maintainer pastes: no entry point may be missing from it."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    listing = text[text.index("```rust"):text.index("## 2.")]
    bound = set(re.findall(r"pub fn (gbp_[a-z0-9_]+)\s*\(", listing))
    assert set(declared_symbols()) <= bound, sorted(set(declared_symbols()) - bound)
    assert bound <= set(declared_symbols()), sorted(bound - set(declared_symbols()))


def test_config_struct_layout_matches_header():
    # 11 x 4-byte scalars, 4 x u8, 3 x i32, 2 x f64, 1 x i32 (+ tail padding) with natural alignment
    assert ctypes.sizeof(CConfig) == 88
    assert CConfig.strict_reference_quirks.offset == 80 and CConfig.world_width.offset == 64
    c = GbpConfig().to_c()
    assert c.num_variables == 10 and abs(c.sigma_factor_interrobot - 0.01) < 1e-9 and c.world_width == 100.0


def test_host_schedule_matches_reference_tests(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "schedules.json")))
    for c in cases:
        oi, oe = magics_b200.gbp_schedule(c["kind"], c["internal"], c["external"])
        got = [[bool(a), bool(b)] for a, b in zip(oi, oe)]
        assert got[: len(c["sequence"])] == c["sequence"], (c["schedule"], c["test"])
        assert len(got) == max(c["internal"], c["external"])


def test_host_schedule_equals_oracle_for_all_small_counts():
    from oracle import oracle

    for kind in range(5):
        for i in range(0, 13):
            for e in range(0, 13):
                a = magics_b200.gbp_schedule(kind, i, e)
                b = oracle.schedule(kind, i, e)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (kind, i, e)
        a = magics_b200.gbp_schedule(kind, 50, 10)
        assert a[0].sum() == 50 and a[1].sum() == 10 and len(a[0]) == 50


def test_host_timesteps_match_reference_tests(golden_dir):
    for c in json.load(open(os.path.join(golden_dir, "variable_timesteps.json"))):
        got = magics_b200.get_variable_timesteps(c["lookahead_horizon"], c["lookahead_multiple"])
        assert got.tolist() == c["expected"]
    assert magics_b200.get_variable_timesteps(18, 3).tolist() == [0, 1, 2, 3, 5, 7, 9, 12, 15, 18]


def test_world_refuses_to_run_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        magics_b200.World(GbpConfig())


def test_scenarios_are_deterministic_and_shaped():
    from magics_b200 import scenarios

    a, b = scenarios.rings(500), scenarios.rings(500)
    assert np.array_equal(a.init_means, b.init_means) and a.cfg.num_variables == 10
    lat = scenarios.lattice(6, 4)
    assert lat.n == 24 and lat.init_means.shape == (24, 10, 4)
    j = scenarios.junction_twoway(per_lane=1)
    assert j.cfg.num_variables == 12 and j.wp_offsets[-1] == 3 * j.n
    # initial means lie on the straight line start -> horizon (robot.rs:1187-1192)
    d = np.diff(lat.init_means[0, :, 0])
    assert (d > 0).all() and np.allclose(lat.init_means[0, :, 1], lat.init_means[0, 0, 1])


def test_world_create_rejects_unsupported_variable_counts():
    """V outside [2, 32] is refused with an error string before any device is touched."""
    import ctypes as C

    from magics_b200 import GbpConfig, load_library

    lib = load_library()
    for v in (0, 1, 33, 64):
        c = GbpConfig(num_variables=v).to_c()
        assert not lib.gbp_world_create(C.byref(c), 0)
        assert b"num_variables" in lib.gbp_last_error()
    out = (C.c_void_p * 2)()
    assert lib.gbp_world_create_local_shards(C.byref(GbpConfig(num_variables=10).to_c()), 0, 0, out) < 0
    assert lib.gbp_world_create_local_shards(C.byref(GbpConfig(num_variables=10).to_c()), 0, 17, out) < 0
    assert not lib.gbp_world_create_shard(C.byref(GbpConfig(num_variables=10).to_c()), 0, 2, 2, None)
