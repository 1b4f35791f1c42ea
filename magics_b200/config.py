"""Host-side mirror of the reference's config scalars that cross the C ABI.

Field names follow gbp_config::Config (crates/gbp_config/src/lib.rs:544-594,
651-680); the ctypes layout is `gbp_config_t` in include/gbp_b200.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

# gbp_config::GbpIterationScheduleKind (gbp_config/src/lib.rs:360-376)
SCHEDULE_CENTERED = 0
SCHEDULE_INTERLEAVE_EVENLY = 1
SCHEDULE_SOON_AS_POSSIBLE = 2
SCHEDULE_LATE_AS_POSSIBLE = 3
SCHEDULE_HALF_BEGINNING_HALF_END = 4

FACTOR_DYNAMIC, FACTOR_INTERROBOT, FACTOR_OBSTACLE, FACTOR_TRACKING = 0, 1, 2, 3


class CConfig(C.Structure):
    _fields_ = [
        ("num_variables", C.c_int32),
        ("sigma_factor_dynamics", C.c_float),
        ("sigma_factor_interrobot", C.c_float),
        ("sigma_factor_obstacle", C.c_float),
        ("sigma_factor_tracking", C.c_float),
        ("safety_distance_multiplier", C.c_float),
        ("comms_radius", C.c_float),
        ("target_speed", C.c_float),
        ("delta_t", C.c_float),
        ("tracking_switch_padding", C.c_float),
        ("tracking_attraction_distance", C.c_float),
        ("enable_dynamic", C.c_uint8),
        ("enable_interrobot", C.c_uint8),
        ("enable_obstacle", C.c_uint8),
        ("enable_tracking", C.c_uint8),
        ("schedule_kind", C.c_int32),
        ("iterations_internal", C.c_int32),
        ("iterations_external", C.c_int32),
        ("world_width", C.c_double),
        ("world_height", C.c_double),
        ("strict_reference_quirks", C.c_int32),
    ]


@dataclass
class GbpConfig:
    """Defaults = GbpSection::default() / RobotSection defaults
    (gbp_config/src/lib.rs:576-594, config/config.toml:39-65)."""

    num_variables: int = 10
    sigma_factor_dynamics: float = 0.1
    sigma_factor_interrobot: float = 0.01
    sigma_factor_obstacle: float = 0.01
    sigma_factor_tracking: float = 0.1
    safety_distance_multiplier: float = 2.2
    comms_radius: float = 20.0
    target_speed: float = 4.0
    delta_t: float = 0.1
    tracking_switch_padding: float = 1.0
    tracking_attraction_distance: float = 2.0
    enable_dynamic: int = 1
    enable_interrobot: int = 1
    enable_obstacle: int = 1
    enable_tracking: int = 0
    schedule_kind: int = SCHEDULE_INTERLEAVE_EVENLY
    iterations_internal: int = 10
    iterations_external: int = 10
    world_width: float = 100.0
    world_height: float = 100.0
    strict_reference_quirks: int = 0  # 1: delete_interrobot_factors as written (lossy HashMap)

    def to_c(self) -> CConfig:
        return CConfig(**{f.name: getattr(self, f.name) for f in fields(self)})
