#!/usr/bin/env python
"""Time sim ticks of one engine build (GBP_B200_LIB) on one GPU — no torch import, CUDA-event timing
through the C ABI.  Used to compare tuning variants of k_iterate in one short GPU visit; bench.py
stays the judged benchmark.

  python scripts/variant_bench.py --workload lattice --steps 5 [--check]
--check also runs 3 ticks of a 2000-robot rings swarm against the oracle (bit parity of the build).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from magics_b200 import World, gbp_schedule, scenarios  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lattice", choices=["lattice", "rings", "dense"])
    ap.add_argument("--robots", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    out = {"lib": os.environ.get("GBP_B200_LIB", "default"), "workload": a.workload}
    if a.check:
        from oracle.oracle import OracleWorld
        from tests.parity import assert_beliefs_match

        sw = scenarios.rings(2000)
        g, o = World(sw.cfg, device=0), OracleWorld(sw.cfg, threads=os.cpu_count() or 1)
        sw.add_to(g)
        sw.add_to(o)
        for _ in range(3):
            g.step()
            o.step()
        errs = assert_beliefs_match(g.read_beliefs(), o.read_beliefs(), what="variant check")
        out["check_max_err"] = max(float(v) for v in errs.values())
        g.close()
        o.close()
    if a.workload == "lattice":
        n = a.robots or 1_000_000
        side = int(round(n ** 0.5))
        sw = scenarios.lattice(side, side)
    elif a.workload == "dense":
        side = int(round((a.robots or 250_000) ** 0.5))
        sw = scenarios.dense_lattice(side, side)
    else:
        sw = scenarios.rings(a.robots or 100_000)
    g = World(sw.cfg, device=0)
    sw.add_to(g)
    oi, oe = gbp_schedule(sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external)
    substeps = int(np.sum(oi & oe))
    for _ in range(a.warmup):
        g.step()
    g.sync()
    g.timer_start()
    for _ in range(a.steps):
        g.step()
    ms = g.timer_stop_ms()
    g.set_profiling(True)  # per-kernel times from a separate pass (events between kernels defeat launch overlap)
    for _ in range(a.steps):
        g.step()
    prof = g.read_profile()
    out.update(robots=sw.n, ms_per_tick=ms / a.steps, M_per_s=sw.n * substeps * a.steps / ms / 1e3,
               profile={k: [v["count"], round(v["ms"], 2)] for k, v in prof.items()} if isinstance(prof, dict) else None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
