#!/usr/bin/env python3
"""Measures the rows of BASELINE.md section 3 on the box it runs on (1 GPU needed; more GPUs fill the multi-device columns):
for BASELINE.json configs 1, 2, 3 (and 5 through the multi-device planet) the reference's own CPU loops (oracle/_ref: as shipped = one
thread per field / per chunk, and all cores) next to the GPU path, plus the roofline fractions from the instrumented work counts.

    python tools/baseline_table.py [--out profiles/r02_baseline_table.json] [--configs 1,2,3,5]

TEST / MEASUREMENT INFRASTRUCTURE: the CPU legs run the oracle, the GPU legs the product through its C ABI."""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402
from unnatural_planets_b200 import Planet, capi, synth  # noqa: E402

FIELDS = ("land", "water", "navigation")
CONFIGS = {
    1: dict(name="1. sphere (lakes), 500^3 x 3 fields, 2048^2 land + water bake (synthetic chart)", shape="sphere", elevation="lakes", grid=500, texture=2048, counts="r02_work_counts_config1.json"),
    2: dict(name="2. torus + craters (Voronoi), 500^3 x 3 fields, 2048^2 land bake", shape="torus", elevation="craters", grid=500, texture=2048, counts="r02_work_counts_config2.json"),
    3: dict(name="3. sphere (lakes), 1000^3 x 3 fields, 4096^2 land + water bake", shape="sphere", elevation="lakes", grid=1000, texture=4096, counts="r02_work_counts_config3.json"),
    5: dict(name="5. sphere (lakes), 1024^3 z-slab sharded, 8192^2 bake (all visible GPUs, one process)", shape="sphere", elevation="lakes", grid=1024, texture=8192, counts="r02_work_counts_config3.json"),
}


def cpu_sampling(shape, elev, res, threads_total):
    """(as shipped: one thread per field, samples/s; all cores, samples/s) on a bounded slab around the equator"""
    out = {}
    for label, per_field_threads, budget in (("as_shipped_3_threads", 1, 2.5e6), ("all_cores", threads_total, 2.0e6 * threads_total)):
        planet = oref.RefPlanet(shape=shape, elevation=elev)
        slices = max(1, int(budget // (res * res)))
        z0 = res // 2 - slices // 2
        planet.grid("water", res, z0=z0, z1=z0 + 1, threads=1)  # touch the code once
        if per_field_threads == 1:
            ts = [threading.Thread(target=planet.grid, args=(f, res), kwargs=dict(z0=z0, z1=z0 + slices, threads=1)) for f in FIELDS]
            t = time.perf_counter()
            [x.start() for x in ts]
            [x.join() for x in ts]
        else:
            t = time.perf_counter()
            for f in FIELDS:
                planet.grid(f, res, z0=z0, z1=z0 + slices, threads=per_field_threads)
        dt = time.perf_counter() - t
        out[label] = {"samples_per_s": 3.0 * slices * res * res / dt, "slices": slices, "seconds": dt}
    return out


def cpu_cells(planet_gpu, shape, elev, res):
    """dualmc + makeMesh tail + meshRemoveDisconnected + meshFlipNormals of the reference over the GPU's own land samples"""
    ref = oref.RefPlanet(shape=shape, elevation=elev)
    if res <= 512:
        dens = planet_gpu.field_samples("land")
        t = time.perf_counter()
        p, n, i = ref.make_mesh(dens, remove_disconnected=True, flip=True)
        dt = time.perf_counter() - t
        return {"cells_per_s": float(res - 1) ** 3 / dt, "seconds": dt, "what": "whole grid: makeMesh + meshRemoveDisconnected + meshFlipNormals, 1 thread", "triangles": int(len(i) // 3)}
    z0 = int(res * 0.23)
    dens = planet_gpu.field_samples("land", z0, z0 + 48)
    t = time.perf_counter()
    v, q = ref.dualmc(dens)
    dt = time.perf_counter() - t
    return {"cells_per_s": 47.0 * (res - 1) ** 2 / dt, "seconds": dt, "what": "dualmc over a 48-slice slab, 1 thread (extrapolated: no clean-up)", "quads": int(len(q))}


def cpu_texels(mesh, texture, threads):
    ref = oref.RefPlanet(shape=mesh["shape"], elevation=mesh["elevation"], coloring="default")
    window = 256
    size = window / texture
    pos, nrm, uv, idx = mesh["arrays"]
    jobs = []
    for k in range(threads):
        u0 = 0.27 + (k % 4) * size * 1.01
        v0 = 0.27 + (k // 4) * size * 1.01
        jobs.append(synth.uv_window_submesh(pos, nrm, uv, idx, u0, v0, size))
    covered = [0] * threads

    def work(k):
        p, n, u, i = jobs[k]
        a, s, h = ref.bake(0, p, n, u, i, window, window)
        covered[k] = int((a != 0).any(axis=2).sum())

    t = time.perf_counter()
    work(0)
    one = time.perf_counter() - t
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t = time.perf_counter()
    [x.start() for x in ts]
    [x.join() for x in ts]
    dt = time.perf_counter() - t
    return {"texels_per_s_one_thread": covered[0] / one, "texels_per_s_all_cores": sum(covered) / dt, "what": "%d UV windows of %d^2 of the land bake, generateTexturesLand per window" % (threads, window)}


def gpu_legs(cfg, devices):
    res, tex = cfg["grid"], cfg["texture"]
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream()
    out = {}
    bufs = [torch.empty(res**3, dtype=torch.float32, device=dev) for _ in range(3)]
    for mode in ("exact", "contract"):
        p = Planet(shape=cfg["shape"], elevation=cfg["elevation"], grid=res, arithmetic=mode)

        def step():
            p.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.cuda_stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            step()
        e1.record(s)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out["k1_%s_ms" % mode] = ms
        out["k1_%s_samples_per_s" % mode] = 3.0 * res**3 / (ms * 1e-3)
        p.close()
    del bufs
    torch.cuda.empty_cache()
    p = Planet(shape=cfg["shape"], elevation=cfg["elevation"], grid=res, devices=devices)
    for _ in range(2):
        m = p.field_meshes_clean(FIELDS)
    ts = []
    for _ in range(5):
        m = p.field_meshes_clean(FIELDS)
        ts.append(p.last_call_s * 1e3)
    out["base_meshes_ms"] = statistics.median(ts)
    out["land_triangles"] = int(len(m["land"]["indices"]) // 3)
    out["cells_per_s"] = 3.0 * float(res - 1) ** 3 / (out["base_meshes_ms"] * 1e-3)
    L = capi.load_library()
    surfaces = {}
    for name in ("land", "water"):
        if len(m[name]["indices"]) == 0:
            continue
        uv = synth.octahedral_uv(m[name]["positions"])
        surfaces[name] = (m[name]["positions"], m[name]["normals"], uv, synth.drop_folded_triangles(uv, m[name]["indices"]))
    himg = {n: [torch.empty(tex * tex * c, dtype=torch.uint8, pin_memory=True) for c in ((4 if n == "water" else 3), 3, 3)] for n in surfaces}
    texels = [0]

    def one(name):
        pos, nrm, uv, idx = surfaces[name]
        o = himg[name]
        capi.check(L.upg_bake_export(p._h, int(name == "water"), pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), tex, tex, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))
        texels[0] += int(L.upg_bake_last_covered_texels())

    def both():
        texels[0] = 0
        th = [threading.Thread(target=one, args=(n,)) for n in surfaces]
        [x.start() for x in th]
        [x.join() for x in th]

    both()
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        both()
        ts.append((time.perf_counter() - t) * 1e3)
    out["bake_ms"] = statistics.median(ts)
    out["bake_texels"] = texels[0]
    out["bake_texels_per_s"] = texels[0] / (out["bake_ms"] * 1e-3)
    out["planet_ms"] = out["base_meshes_ms"] + out["bake_ms"]
    out["devices"] = len(devices)
    land = surfaces.get("land")
    p.close()
    return out, (p, land)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_baseline_table.json"))
    ap.add_argument("--configs", default="1,2,3,5")
    a = ap.parse_args()
    threads = os.cpu_count() or 1
    ngpu = torch.cuda.device_count()
    peak = C.c_float(0)
    capi.check(capi.load_library().upg_measure_fp32_peak(0, C.byref(peak)))
    table = {"host_threads": threads, "gpus_visible": ngpu, "gpu": torch.cuda.get_device_name(0), "fp32_peak_tflops": float(peak.value), "rows": {}}
    for c in [int(x) for x in a.configs.split(",")]:
        cfg = CONFIGS[c]
        row = {"name": cfg["name"]}
        counts = json.load(open(os.path.join(ROOT, "profiles", cfg["counts"])))
        one_eval = counts["fields"]["land"]["flops"] + 3.0
        row["flops_per_grid_node_reference"] = counts["flops_per_grid_node"]
        row["flops_one_evaluation_per_node"] = one_eval
        devices = list(range(ngpu)) if c == 5 else [0]
        gpu, (planet, land) = gpu_legs(cfg, devices)
        row["gpu"] = gpu
        n = float(cfg["grid"]) ** 3
        for mode in ("exact", "contract"):
            row["k1_frac_%s" % mode] = one_eval * n / (gpu["k1_%s_ms" % mode] * 1e-3) / 1e12 / table["fp32_peak_tflops"]
        if "texels" in counts:
            # texels of the land + water bake with the per-texel counts of each surface (covered texels split evenly is not assumed: use land's)
            row["flops_per_land_texel"] = counts["texels"]["land"]["flops"]
        row["cpu_sampling"] = cpu_sampling(cfg["shape"], cfg["elevation"], cfg["grid"], threads)
        probe = Planet(shape=cfg["shape"], elevation=cfg["elevation"], grid=cfg["grid"])
        row["cpu_isosurface"] = cpu_cells(probe, cfg["shape"], cfg["elevation"], cfg["grid"])
        probe.close()
        if land is not None:
            row["cpu_bake"] = cpu_texels(dict(shape=cfg["shape"], elevation=cfg["elevation"], arrays=land), cfg["texture"], threads)
        row["speedup_sampling_vs_all_cores"] = gpu["k1_exact_samples_per_s"] / row["cpu_sampling"]["all_cores"]["samples_per_s"]
        table["rows"][str(c)] = row
        print(json.dumps({str(c): row}), flush=True)
    json.dump(table, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
