import numpy as np, sys
sys.path.insert(0, '.')
from magics_b200 import World, scenarios
from oracle.oracle import OracleWorld
sw = scenarios.complex_environment(2)
g, o = World(sw.cfg), OracleWorld(sw.cfg)
sw.add_to(g); sw.add_to(o)
rng = np.random.default_rng(0)
xy = np.concatenate([
    rng.uniform([-130.0, -90.0], [130.0, 90.0], size=(20000, 2)),
    np.array([[0.0, 0.0], [-125.0, 87.5], [125.0, -87.5], [124.99999999, -87.49999999], [1e30, 0.0],
              [-1e30, 0.0], [np.nan, 0.0], [0.0, np.inf], [-125.0 - 1e-13, 0.0]]),
    (np.arange(-1000, 1001)[:, None] * np.array([[0.125, 0.0]])),
])
pg = g.sdf_lookup(xy); po = o.sdf_lookup(xy)
for a, b, name in zip(pg, po, ("px", "py", "value")):
    bad = np.nonzero(~((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))))[0]
    print(name, "mismatches", bad.size)
    for k in bad[:10]:
        print("  idx", k, "xy", xy[k], "gpu", a[k], "oracle", b[k])
