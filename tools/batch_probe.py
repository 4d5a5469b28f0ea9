#!/usr/bin/env python3
"""Runs single planets of bench.py's batch leg (BASELINE configs[3]) with UPG_TRACE phases: python tools/batch_probe.py 15 21 55"""
import os
import sys
import time

os.environ.setdefault("UPG_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from unnatural_planets_b200 import ELEVATIONS, SHAPES, Planet  # noqa: E402

for rep in range(2):
    for i in [int(a) for a in sys.argv[1:]]:
        seed = bench.seed_for(1000 + i)
        shape, elev = SHAPES[seed[0] % len(SHAPES)], ELEVATIONS[i % len(ELEVATIONS)]
        p = Planet(shape=shape, elevation=elev, grid=500, seed=seed)
        t = time.perf_counter()
        m = p.field_meshes_clean(("land", "water", "navigation"))
        print("planet %d %s/%s: %.2f ms (C call %.2f ms), land %d triangles" % (i, shape, elev, (time.perf_counter() - t) * 1e3, p.last_call_s * 1e3, len(m["land"]["indices"]) // 3), file=sys.stderr, flush=True)
        p.close()
