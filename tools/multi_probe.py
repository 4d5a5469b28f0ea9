#!/usr/bin/env python3
"""ONE planet over the GPUs of one node from ONE process (Planet(devices=[...]) -> upg_planet_desc::device_count): base meshes and
export bake, median wall clock of the C calls, for 1 .. N devices.   python tools/multi_probe.py [--grid 1024] [--texture 8192]"""
import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1024)
ap.add_argument("--texture", type=int, default=8192)
ap.add_argument("--repeat", type=int, default=5)
ap.add_argument("--counts", default="")
a = ap.parse_args()
F = ("land", "water", "navigation")
ngpu = torch.cuda.device_count()
counts = [int(x) for x in a.counts.split(",")] if a.counts else [n for n in (1, 2, 4, 8) if n <= ngpu]
out = {"grid": a.grid, "texture": a.texture, "gpus_visible": ngpu, "runs": []}
ref_digest = None
T = a.texture
himg = {"land": [np.zeros((T, T, c), np.uint8) for c in (3, 3, 3)], "water": [np.zeros((T, T, c), np.uint8) for c in (4, 3, 3)]}
for n in counts:
    p = Planet(shape="sphere", elevation="lakes", grid=a.grid, devices=list(range(n)))
    m = p.field_meshes_clean(F)
    del m
    p.field_meshes_clean(F)
    times = []
    for _ in range(a.repeat):
        m = p.field_meshes_clean(F)
        times.append(p.last_call_s * 1e3)
    h = hashlib.sha1()
    for f in F:
        for k in ("positions", "normals", "indices"):
            h.update(m[f][k].tobytes())
    digest = h.hexdigest()[:16]
    ref_digest = ref_digest or digest
    meshes = {}
    for name in ("land", "water"):
        uv = synth.octahedral_uv(m[name]["positions"])
        meshes[name] = (m[name]["positions"], m[name]["normals"], uv, synth.drop_folded_triangles(uv, m[name]["indices"]))
    import ctypes as C
    from unnatural_planets_b200 import capi

    L = capi.load_library()

    def one(name, water):
        pos, nrm, uv, idx = meshes[name]
        o = himg[name]
        capi.check(L.upg_bake_export(p._h, water, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), T, T, o[0].ctypes.data_as(C.c_void_p), o[1].ctypes.data_as(C.c_void_p), o[2].ctypes.data_as(C.c_void_p)))

    def bake_both(concurrent):
        if concurrent:
            ts = [threading.Thread(target=one, args=x) for x in (("land", 0), ("water", 1))]
            [t.start() for t in ts]
            [t.join() for t in ts]
        else:
            one("land", 0)
            one("water", 1)

    res = {"devices": n, "base_meshes_ms": round(statistics.median(times), 2), "base_meshes_ms_min": round(min(times), 2), "mesh_digest_equal_to_first": digest == ref_digest}
    for label, conc in (("bake_sequential_ms", False), ("bake_two_threads_ms", True)):
        bake_both(conc)
        ts = []
        for _ in range(a.repeat):
            t = time.perf_counter()
            bake_both(conc)
            ts.append((time.perf_counter() - t) * 1e3)
        res[label] = round(statistics.median(ts), 2)
    hb = hashlib.sha1()
    for name in ("land", "water"):
        for x in himg[name]:
            hb.update(x.tobytes())
    res["bake_digest"] = hb.hexdigest()[:16]
    res["planet_ms"] = round(res["base_meshes_ms"] + min(res["bake_sequential_ms"], res["bake_two_threads_ms"]), 2)
    out["runs"].append(res)
    print(json.dumps(res), flush=True)
    p.close()
print(json.dumps(out))
