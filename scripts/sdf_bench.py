#!/usr/bin/env python
"""Time SDF generation for the `Collaborative Complex` environment (2000 x 1400 pixels):
the device path through the C ABI (host tile grid in, RGB8 image out) and the oracle on one host core."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magics_b200 import Environment, env_to_sdf_image  # noqa: E402
from oracle import oracle as oo  # noqa: E402

gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "env_to_png.json")))
env = Environment(**gold["environments"]["Collaborative Complex"])
env_to_sdf_image(env)
ts = []
for _ in range(5):
    t = time.perf_counter()
    img = env_to_sdf_image(env)
    ts.append(time.perf_counter() - t)
t = time.perf_counter()
ref = oo.env_to_sdf_image(env)
cpu = time.perf_counter() - t
h, w = env.image_shape
print(json.dumps({"image": [h, w], "device_e2e_ms_best": min(ts) * 1e3, "oracle_1core_ms": cpu * 1e3,
                  "equal": bool((img == ref).all()), "Mpixel_per_s_device_e2e": h * w / min(ts) / 1e6}))
