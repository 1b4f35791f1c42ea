import os, sys, time, json
import numpy as np
sys.path.insert(0, "/root/repo")
from magics_b200 import World, scenarios, pinned_empty
sw = scenarios.lattice(1000, 1000)
g = World(sw.cfg, device=0); sw.add_to(g)
n = sw.n
ant = pinned_empty((n,), np.uint8); wpi = pinned_empty((n,), np.int32); ant[:] = 1; wpi[:] = 1
means = [pinned_empty((n, sw.cfg.num_variables, 4), np.float64) for _ in range(2)]
for k in range(2): g.read_means_into(means[k])
for _ in range(3): g.step()
def run(mode, steps=8):
    g.sync(); t0 = time.perf_counter()
    for k in range(steps):
        if "up" in mode:
            g.set_comms(ant, None); g.set_waypoint_index(wpi)
        g.step()
        if "rb" in mode:
            g.read_means_into_async(means[k % 2])
        if "rbsync" in mode:
            g.readback_wait()
    g.readback_wait(); g.sync()
    return (time.perf_counter() - t0) * 1e3 / steps
out = {}
for mode in ["plain", "up", "rb", "up+rb", "rbsync", "plain"]:
    out[mode] = round(run(mode), 2)
# raw copy speed
t0 = time.perf_counter(); g.read_means_into(means[0]); out["sync_readback_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
print(json.dumps(out))
