#!/usr/bin/env python3
"""Host-buffer export bake of the land and water surfaces of one planet from two host threads (the bench's planet leg), median of
N calls: python tools/bake_e2e_probe.py [--grid 500] [--texture 4096] [--repeat 7]"""
import argparse
import ctypes as C
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet, capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=500)
ap.add_argument("--texture", type=int, default=4096)
ap.add_argument("--repeat", type=int, default=7)
a = ap.parse_args()
L = capi.load_library()
p = Planet(shape="sphere", elevation="lakes", coloring="default", grid=a.grid)
T = a.texture
meshes, himg = {}, {}
for name, water in (("land", 0), ("water", 1)):
    m = p.field_mesh(name)
    uv = synth.octahedral_uv(m["positions"])
    meshes[name] = (m["positions"], np.ascontiguousarray(-m["normals"]), uv, synth.drop_folded_triangles(uv, m["indices"]))
    himg[name] = [torch.empty(T * T * c, dtype=torch.uint8, pin_memory=True) for c in ((4 if water else 3), 3, 3)]


def one(name, water):
    pos, nrm, uv, idx = meshes[name]
    o = himg[name]
    capi.check(L.upg_bake_export(p._h, water, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), T, T, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))


def concurrent():
    ts = [threading.Thread(target=one, args=x) for x in (("land", 0), ("water", 1))]
    [t.start() for t in ts]
    [t.join() for t in ts]


def sequential():
    one("land", 0)
    one("water", 1)


out = {}
for label, fn in (("concurrent_ms", concurrent), ("sequential_ms", sequential)):
    fn()
    times = []
    for _ in range(a.repeat):
        t = time.perf_counter()
        fn()
        times.append((time.perf_counter() - t) * 1e3)
    out[label] = round(statistics.median(times), 2)
    out[label + "_min"] = round(min(times), 2)
out["digest"] = [int(x.to(torch.int64).sum().item()) for x in himg["land"] + himg["water"]]
print(out)
