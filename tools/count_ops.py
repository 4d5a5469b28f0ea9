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


def count_texels(shape, elevation, coloring, mesh_res, purposes=("land", "water")):
    """operation counts of terrainTile per surface point: the points are the vertices (positions + normals) of the oracle's own
    isosurface of the land / water field at `mesh_res`^3 -- spread over the whole planet like the covered texels of a bake"""
    d = os.path.join(ROOT, "tests", "hostlib")
    subprocess.check_call(["make", "-s", "-C", d, "libupg_hostcount.so", "libupg_hosttest.so"])
    out = {}
    libs = {}
    for name in ("libupg_hostcount.so", "libupg_hosttest.so"):
        H = C.CDLL(os.path.join(d, name))
        H.upgh_create.restype = C.c_void_p
        H.upgh_create.argtypes = [C.c_int] * 5 + [C.c_uint64] * 2 + [C.c_float, C.c_int]
        H.upgh_tiles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        libs[name] = H
    HC, HP = libs["libupg_hostcount.so"], libs["libupg_hosttest.so"]
    HC.upghc_get.argtypes = [C.c_void_p]
    nops = HC.upghc_op_count()
    L = counted_lib()
    names = [L.upc_op_name(i).decode() for i in range(nops)]
    colorings = ("default", "barren", "debug")
    args = (oport.SHAPES.index(shape), oport.ELEVATIONS.index(elevation), colorings.index(coloring), 0, 0, oport.DEFAULT_SEED[0], oport.DEFAULT_SEED[1], 4000.0, mesh_res)
    hc, hp = HC.upgh_create(*args), HP.upgh_create(*args)
    planet = oport.OraclePlanet(shape, elevation)
    for purpose in purposes:
        grid = planet.grid("land" if purpose == "land" else "water", mesh_res, threads=os.cpu_count() or 1)
        m = oport.isosurface(grid)
        pos = np.ascontiguousarray(m["positions"], dtype=np.float32)
        nrm = np.ascontiguousarray(-m["normals"], dtype=np.float32)  # meshFlipNormals: the generator bakes the flipped mesh
        keep = np.isfinite(nrm).all(axis=1) & (np.abs((nrm * nrm).sum(axis=1) - 1) < 1e-3)
        pos, nrm = pos[keep], nrm[keep]
        code = {"land": 1, "water": 2}[purpose]
        got = np.zeros((len(pos), 13), dtype=np.float32)
        want = np.zeros((len(pos), 13), dtype=np.float32)
        HC.upghc_reset()
        HC.upgh_tiles(C.c_void_p(hc), code, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), len(pos), got.ctypes.data_as(C.c_void_p))
        counts = np.zeros(nops, dtype=np.uint64)
        HC.upghc_get(counts.ctypes.data_as(C.c_void_p))
        HP.upgh_tiles(C.c_void_p(hp), code, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), len(pos), want.ctypes.data_as(C.c_void_p))
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), "counted colouring build differs from the plain host build"
        per = {n: float(c) / len(pos) for n, c in zip(names, counts)}
        per["flops"] = per["fadd"] + per["fmul"] + per["fdiv"] + per["fsqrt"]
        per["flops_fp64"] = per["dadd"] + per["dmul"] + per["ddiv"] + per["dsqrt"]
        per["points"] = int(len(pos))
        out[purpose] = per
        print("%-6s texel %8.1f FP32 flops (fadd %.1f fmul %.1f fdiv %.2f fsqrt %.2f) + %.1f FP64 flops + fcmp %.1f i2f %.1f f2i %.1f special %.2f dspecial %.2f   [%d surface points]" % (
            purpose, per["flops"], per["fadd"], per["fmul"], per["fdiv"], per["fsqrt"], per["flops_fp64"], per["fcmp"], per["i2f"], per["f2i"], per["special"], per["dspecial"], len(pos)), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1000)
    ap.add_argument("--stride", type=int, default=2)
    ap.add_argument("--shape", default="sphere")
    ap.add_argument("--elevation", default="lakes")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--json", default=None)
    ap.add_argument("--texels", action="store_true", help="also count terrainTile per surface point (colouring default)")
    ap.add_argument("--coloring", default="default")
    ap.add_argument("--mesh-res", type=int, default=200)
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
    if a.texels:
        out["texels"] = count_texels(a.shape, a.elevation, a.coloring, a.mesh_res)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
