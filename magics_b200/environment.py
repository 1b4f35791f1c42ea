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


class CEnvironment(C.Structure):
    """include/gbp_b200.h gbp_environment_t"""
    _fields_ = [("nrows", C.c_int32), ("ncols", C.c_int32), ("tiles", C.POINTER(C.c_uint32)),
                ("tile_size", C.c_float), ("path_width", C.c_float), ("resolution", C.c_uint32),
                ("expansion", C.c_float), ("blur", C.c_float), ("n_obstacles", C.c_int32)]


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
        import yaml

        d = yaml.safe_load(text)
        s = d["tiles"]["settings"]
        sdf = s.get("sdf", {})
        return cls(grid=list(d["tiles"]["grid"]), tile_size=float(s["tile-size"]), path_width=float(s["path-width"]),
                   resolution=int(sdf.get("resolution", 200)), expansion=float(sdf.get("expansion", 0.0)),
                   blur=float(sdf.get("blur", 0.0)), obstacles=list(d.get("obstacles") or []))

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

    def c_struct(self):
        """(CEnvironment, keep-alive array)"""
        codes = self.tile_codes()
        ce = CEnvironment(self.nrows, self.ncols, codes.ctypes.data_as(C.POINTER(C.c_uint32)), self.tile_size,
                          self.path_width, self.resolution, self.expansion, self.blur, len(self.obstacles))
        return ce, codes
