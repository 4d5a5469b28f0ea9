#!/usr/bin/env python3
"""One pass of the hot path of one planet, for ncu (launch list / --set full captures); not a benchmark.
    python tools/profile_step.py [--grid 500] [--texture 2048] [--repeat 2]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=500)
ap.add_argument("--texture", type=int, default=2048)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--flowers", type=int, default=0)
ap.add_argument("--poles", type=int, default=0)
a = ap.parse_args()
p = Planet(shape="sphere", elevation="lakes", coloring="default", grid=a.grid, poles=bool(a.poles), flowers=bool(a.flowers))
for r in range(a.repeat):
    meshes = {}
    final = p.field_meshes_clean(("land", "water", "navigation"))  # K1 fused + K2..K4 + K9: what the generator keeps
    for f, m in p.field_meshes(("land", "water", "navigation")).items():  # raw meshes feed the synthetic UV chart below
        if f != "navigation":
            uv = synth.octahedral_uv(m["positions"])
            meshes[f] = (m["positions"], -m["normals"], uv, synth.drop_folded_triangles(uv, m["indices"]))
    for f, water in (("land", 0), ("water", 1)):
        pos, nrm, uv, idx = meshes[f]
        out = p.bake_export(water, pos, nrm, uv, idx, a.texture, a.texture)
    print("pass", r, "covered", p.last_covered_texels, "albedo mean", float(out[0].mean()))
