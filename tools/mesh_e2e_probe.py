#!/usr/bin/env python3
"""Base meshes of one planet through the host-buffer C call (upg_field_meshes_clean: the bench's planet leg), median of N calls:
    [UPG_LIBRARY=...] python tools/mesh_e2e_probe.py [--grid 1000] [--repeat 7]"""
import argparse
import hashlib
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1000)
ap.add_argument("--repeat", type=int, default=7)
ap.add_argument("--shape", default="sphere")
ap.add_argument("--elevation", default="lakes")
a = ap.parse_args()
F = ("land", "water", "navigation")
p = Planet(shape=a.shape, elevation=a.elevation, grid=a.grid)
out = {"lib": os.path.basename(os.environ.get("UPG_LIBRARY", "libupg.so")), "grid": a.grid}
for label, fn in (("clean_ms", lambda: p.field_meshes_clean(F)), ("raw_ms", lambda: p.field_meshes(F))):
    m = fn()
    del m
    fn()
    times = []
    for _ in range(a.repeat):
        m = fn()
        times.append(p.last_call_s * 1e3)
        if len(times) < a.repeat:
            del m
    out[label] = round(statistics.median(times), 2)
    out[label + "_min"] = round(min(times), 2)
    h = hashlib.sha256()
    for f in F:
        for k in ("positions", "normals", "indices"):
            h.update(m[f][k].tobytes())
    out[label.replace("_ms", "_sha")] = h.hexdigest()[:16]
    del m
print(out)
