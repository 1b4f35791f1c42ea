"""What `__graft_entry__.build()` produced, inspected without a GPU: the product library carries sm_100a machine code
for every kernel DESIGN section 4 names, with the register budgets DESIGN sections 4-5 quote (a silent change of a
kernel's occupancy would otherwise only show up as a slower bench on the GPU box)."""
import os
import re
import shutil
import subprocess

import pytest

from magics_b200.world import library_path

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP) or not os.path.exists(library_path()),
                                reason="cuobjdump or the built library missing")


def _resources():
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", library_path()], capture_output=True, text=True,
                         check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    return res


def test_library_holds_sm_100a_code_only():
    out = subprocess.run([CUOBJDUMP, "--list-elf", library_path()], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, out


def test_every_kernel_of_the_design_is_in_the_library_with_its_register_budget():
    res = _resources()
    by_name = lambda key: {k: v for k, v in res.items() if key in k}
    axis = by_name("k_iterate_axis")
    assert len(axis) == 9, sorted(axis)  # EXT / INT / EXT+INT  x  PART 0, 1, 2
    assert all(v["reg"] <= 96 and v["stack"] <= 32 and v["local"] == 0 for v in axis.values()), axis  # 20 warps / SM
    general = {k: v for k, v in res.items() if re.search(r"9k_iterateILb", k)}
    assert len(general) == 3 and all(v["reg"] <= 168 for v in general.values()), general
    assert all(v["reg"] <= 128 for v in by_name("k_edge_messages").values())
    for name in ["k_tick_fused", "k_edge_messages", "k_prior_both", "k_prior_horizon", "k_prior_current", "k_cell_keys",
                 "k_neighbours", "k_edge_diff", "k_robot_collisions", "k_env_collisions", "k_track",
                 "k_reached_waypoint", "k_halo_pack", "k_halo_unpack", "k_env_raster", "k_blur_rows", "k_blur_cols",
                 "k_reset_variables", "k_remove_robots", "k_sdf_lookup"]:
        assert by_name(name), f"{name} is not in the library"
