#!/usr/bin/env python3
"""How many texels of a bake would an FP32 HSLuv change?  ("HSLuv: evaluate in FP32 ... and measure flips against the
FP64 path".)  CPU experiment, no GPU: the product's colouring headers are compiled for the host twice (tests/hostlib/upg_host_eval.cpp) --
as shipped (HSLuv in double like hsluv.c) and with -DUPG_HSLUV_REAL=float (csrc/upg_color.cuh: the same expressions with sinf / cosf /
powf / cbrtf / atan2f) -- and terrainTile is evaluated for the same surface points by both; the albedo is quantised like the bake does
(numeric_cast<uint8>(saturate(v) * 255), truncating).  Host float arithmetic is the reference's, so the double column is the reference.

    python tools/hsluv_precision.py [--points 200000] [--out profiles/r02_hsluv_fp32.json]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(real, out):
    csrc = os.path.join(ROOT, "unnatural_planets_b200", "csrc")
    src = os.path.join(ROOT, "tests", "hostlib", "upg_host_eval.cpp")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-gnu-unique", "-w", "-x", "c++", "-I" + csrc, "-DUPG_HSLUV_REAL=" + real, "-shared", "-Wl,-Bsymbolic", "-o", out, src]
    subprocess.check_call(cmd)
    H = C.CDLL(out)
    H.upgh_create.restype = C.c_void_p
    H.upgh_create.argtypes = [C.c_int] * 5 + [C.c_uint64] * 2 + [C.c_float, C.c_int]
    H.upgh_destroy.argtypes = [C.c_void_p]
    H.upgh_tiles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    return H


def tiles(H, cfg, purpose, pos, nrm):
    from unnatural_planets_b200.planet import COLORINGS, ELEVATIONS, SHAPES

    s, e, c, p, f = cfg
    h = H.upgh_create(SHAPES.index(s), ELEVATIONS.index(e), COLORINGS.index(c), int(p), int(f), 1346519564496, 42245614964156, 4000.0, 500)
    out = np.zeros((len(pos), 13), dtype=np.float32)
    H.upgh_tiles(C.c_void_p(h), purpose, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), len(pos), out.ctypes.data_as(C.c_void_p))
    H.upgh_destroy(C.c_void_p(h))
    return out


def u8(v):
    return (np.clip(np.nan_to_num(v, nan=0.0), 0.0, 1.0).astype(np.float32) * np.float32(255.0)).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=200000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_hsluv_fp32.json"))
    a = ap.parse_args()
    import bake_common as bc

    tmp = tempfile.mkdtemp(prefix="upg_hsluv_")
    Hd, Hf = build("double", os.path.join(tmp, "d.so")), build("float", os.path.join(tmp, "f.so"))
    result = {"what": "terrainTile albedo, HSLuv evaluated in double (as shipped = the reference) against the same expressions in single precision; u8 = numeric_cast<uint8>(saturate(v) * 255)", "cases": []}
    for cfg, purpose, name in ((("sphere", "lakes", "default", 1, 1), 1, "land"), (("sphere", "lakes", "default", 1, 1), 2, "water"), (("torus", "islands", "default", 0, 0), 1, "land"), (("sphere", "lakes", "barren", 0, 0), 1, "land")):
        pos, nrm = bc.tile_probes(cfg[0], n=a.points, seed=23)
        d, f = tiles(Hd, cfg, purpose, pos, nrm), tiles(Hf, cfg, purpose, pos, nrm)
        qa, qb = u8(d[:, 0:3]), u8(f[:, 0:3])
        diff = np.abs(qa.astype(np.int16) - qb.astype(np.int16)).max(axis=1)
        fd = np.abs(d[:, 0:3] - f[:, 0:3]).max(axis=1)
        other = int((d[:, 3:] != f[:, 3:]).any(axis=1).sum())
        case = {
            "config": "%s/%s/%s poles=%d flowers=%d" % cfg, "surface": name, "texels": int(len(pos)),
            "texels_with_a_different_u8_channel": int((diff > 0).sum()), "share": float((diff > 0).mean()), "max_u8_difference": int(diff.max()),
            "texels_off_by_more_than_one": int((diff > 1).sum()), "max_float_difference": float(fd.max()), "mean_float_difference": float(fd.mean()),
            "texels_with_another_field_changed": other,
        }
        print(json.dumps(case), flush=True)
        result["cases"].append(case)
    json.dump(result, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
