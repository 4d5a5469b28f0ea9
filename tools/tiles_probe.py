#!/usr/bin/env python3
"""upg_tiles on 2^20 points of the sphere in random order (a warp's 32 tiles share no noise cell) and in spatially sorted order:
    [UPG_LIBRARY=...] python tools/tiles_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet  # noqa: E402

p = Planet(shape="sphere", elevation="lakes", coloring="default", grid=100)
rng = np.random.default_rng(1)
n = 1 << 20
d = rng.normal(size=(n, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
pos = (d * 1100).astype(np.float32)
order = np.lexsort((pos[:, 0] // 8, pos[:, 1] // 8, pos[:, 2] // 8))
out = {}
for label, perm in (("random", np.arange(n)), ("sorted", order)):
    P, N = np.ascontiguousarray(pos[perm]), np.ascontiguousarray(d[perm])
    t = p.tiles("land", P, N)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        t = p.tiles("land", P, N)
        ts.append((time.perf_counter() - t0) * 1e3)
    out[label + "_ms"] = round(min(ts), 2)
    out[label + "_sum"] = float(np.nan_to_num(t.astype(np.float64)).sum())
print(os.path.basename(os.environ.get("UPG_LIBRARY", "libupg.so")), out)
