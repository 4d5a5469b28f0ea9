#!/usr/bin/env python3
"""Exact arithmetic work per SDF sample, from the COUNTED build of the CPU restatement (oracle/up_count.cpp: up_oracle.cpp compiled
with `float` replaced by an instrumented type; results bit-identical to the plain build, asserted here).

    python tools/count_ops.py [--grid 1000] [--stride 2] [--shape sphere] [--elevation lakes] [--json out.json]

Counts every scalar operation the reference's sampling loop (MarchingCubes::updateByPosition x terrainSdfLand / Water / Navigation)
performs at every `stride`-th node per axis of the grid and prints the averages per sample.  flops = fadd + fmul + fdiv + fsqrt
(the reference has no FMA; an FMA counts 2 under that convention); comparisons, negations, conversions and libm calls are listed
separately.  TEST INFRASTRUCTURE: runs the oracle, never the product."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port as oport  # noqa: E402


def counted_lib():
    path = os.path.join(ROOT, "oracle", "libup_count.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "count"])
    L = C.CDLL(path)
    L.upo_planet_create.restype = C.c_void_p
    L.upo_planet_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64]
    L.upo_planet_destroy.argtypes = [C.c_void_p]
    L.upc_op_name.restype = C.c_char_p
    L.upc_grid_counts.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def count_field(L, shape, elevation, field, res, stride, threads, box=4000.0, check=True):
    h = L.upo_planet_create(oport.SHAPES.index(shape), oport.ELEVATIONS.index(elevation), *oport.DEFAULT_SEED)
    nops = L.upc_op_count()
    names = [L.upc_op_name(i).decode() for i in range(nops)]
    counts = np.zeros(nops, dtype=np.uint64)
    nodes = C.c_uint64(0)
    if check:
        # bit-identity of the counted build against the plain one on two full slices
        z0 = res // 2
        got = np.zeros((2, res, res), dtype=np.float32)
        c2, n2 = np.zeros(nops, dtype=np.uint64), C.c_uint64(0)
        L.upc_grid_counts(C.c_void_p(h), oport.FIELDS[field], box, res, z0, z0 + 2, 1, threads, c2.ctypes.data_as(C.c_void_p), C.byref(n2), got.ctypes.data_as(C.c_void_p))
        want = oport.OraclePlanet(shape, elevation).grid(field, res, z0=z0, z1=z0 + 2, threads=threads)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), "counted build differs from the plain oracle"
    L.upc_grid_counts(C.c_void_p(h), oport.FIELDS[field], box, res, 0, res, stride, threads, counts.ctypes.data_as(C.c_void_p), C.byref(nodes), None)
    L.upo_planet_destroy(C.c_void_p(h))
    per = {n: float(c) / nodes.value for n, c in zip(names, counts)}
    per["flops"] = per["fadd"] + per["fmul"] + per["fdiv"] + per["fsqrt"]
    per["nodes"] = int(nodes.value)
    return per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1000)
    ap.add_argument("--stride", type=int, default=2)
    ap.add_argument("--shape", default="sphere")
    ap.add_argument("--elevation", default="lakes")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    L = counted_lib()
    out = {"shape": a.shape, "elevation": a.elevation, "grid": a.grid, "stride": a.stride, "fields": {}}
    for field in ("land", "water", "navigation"):
        t = time.time()
        per = count_field(L, a.shape, a.elevation, field, a.grid, a.stride, a.threads)
        out["fields"][field] = per
        print("%-10s %8.1f flops/sample (fadd %.1f fmul %.1f fdiv %.2f fsqrt %.2f) + fcmp %.1f fneg %.1f i2f %.1f f2i %.1f special %.2f   [%d nodes, %.0f s]" % (
            field, per["flops"], per["fadd"], per["fmul"], per["fdiv"], per["fsqrt"], per["fcmp"], per["fneg"], per["i2f"], per["f2i"], per["special"], per["nodes"], time.time() - t), flush=True)
    out["flops_per_grid_node"] = sum(out["fields"][f]["flops"] for f in out["fields"])
    out["special_per_grid_node"] = sum(out["fields"][f]["special"] for f in out["fields"])
    print("per grid node (land + water + navigation, each evaluated on its own like the reference's three tasks): %.1f flops, %.2f libm calls" % (out["flops_per_grid_node"], out["special_per_grid_node"]))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
