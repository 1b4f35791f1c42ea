"""magics_b200 — B200-native Gaussian Belief Propagation engine for the GBP
iterate hot path of AU-Master-Thesis/magics.

The product is the CUDA shared library built from magics_b200/csrc (C ABI in
include/gbp_b200.h); this package is the thin host-side mirror of the
reference's `FactorGraph` / `gbp_schedule` interface used by tests and bench.
There is no CPU fallback: constructing a `World` without the CUDA library or
without a GPU raises.
"""
from .config import (  # noqa: F401
    FACTOR_DYNAMIC,
    FACTOR_INTERROBOT,
    FACTOR_OBSTACLE,
    FACTOR_TRACKING,
    SCHEDULE_CENTERED,
    SCHEDULE_HALF_BEGINNING_HALF_END,
    SCHEDULE_INTERLEAVE_EVENLY,
    SCHEDULE_LATE_AS_POSSIBLE,
    SCHEDULE_SOON_AS_POSSIBLE,
    GbpConfig,
)
from .environment import Environment, Obstacle  # noqa: F401
from .world import (World, env_to_sdf_image, gbp_schedule, get_variable_timesteps, library_path,  # noqa: F401
                    load_library, pinned_empty)
