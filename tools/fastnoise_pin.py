#!/usr/bin/env python3
"""The FastNoiseLite pin hook (VERDICT r1 missing #7).  The coherent-noise arithmetic of the reference lives in Auburn/FastNoiseLite
(Cpp/FastNoiseLite.h @ 785f37a9), an un-vendored submodule that is absent here; oracle/up_noise.h restates it.  The day the header
is available:

    python tools/fastnoise_pin.py --header /path/to/FastNoiseLite.h [--golden tests/golden/randvecs3d_upstream.npy]

  1. builds the reference's translation units against the REAL header (make -C oracle ref FASTNOISE=...  ->  oracle/_ref_upstream),
  2. evaluates every generator x fractal combination the generator uses (and the rest of the grid) through BOTH libraries at the
     same points and reports, per combination, the fraction of bit-identical values and the largest difference,
  3. extracts upstream's RandVecs3D (the one constant the restatement could not reconstruct) and, with --golden, saves it: load it
     through upg_planet_set_cellular_table / Planet.set_cellular_table and the Cellular layers follow upstream as well.
Exit status 1 when any combination differs.  TEST INFRASTRUCTURE."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref import NoiseCfg  # noqa: E402

TYPES = ("simplex", "simplex_reduced", "cellular", "perlin", "cubic", "value")
FRACTALS = ("none", "fbm", "ridged", "pingpong")


def load(path):
    L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    L.upref_fastnoise_eval.argtypes = [C.POINTER(NoiseCfg), C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.upref_cellular_table.argtypes = [C.c_void_p]
    return L


def bounding(octaves, gain):
    g = np.float32(abs(gain))
    amp, total = g, np.float32(1)
    for _ in range(1, octaves):
        total = np.float32(total + amp)
        amp = np.float32(amp * g)
    return float(np.float32(1) / total)


def compare(restated, upstream, points=4096, seed=5):
    rng = np.random.default_rng(seed)
    pos = (rng.random((points, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(4000)
    rows = []
    for t, tname in enumerate(TYPES):
        for f, fname in enumerate(FRACTALS):
            for octaves, lac, gain, freq in ((4, 2.0, 0.5, 0.0013), (3, 1.5, -0.4, 0.01), (2, 2.0, 0.3, 0.7)):
                for dist, op in (((1, 1),) if tname != "cellular" else ((0, 0), (0, 1), (1, 2), (2, 3), (3, 4), (0, 5), (1, 6))):
                    cfg = NoiseCfg(t, f, dist, op, 1234 + 7 * t + f, octaves, lac, gain, freq, bounding(octaves, gain))
                    a, b = np.zeros(points, np.float32), np.zeros(points, np.float32)
                    restated.upref_fastnoise_eval(C.byref(cfg), pos.ctypes.data_as(C.c_void_p), points, 3, a.ctypes.data_as(C.c_void_p))
                    upstream.upref_fastnoise_eval(C.byref(cfg), pos.ctypes.data_as(C.c_void_p), points, 3, b.ctypes.data_as(C.c_void_p))
                    rows.append({"type": tname, "fractal": fname, "octaves": octaves, "frequency": freq, "distance": dist, "operation": op,
                                 "bit_identical": float((a.view(np.uint32) == b.view(np.uint32)).mean()), "max_abs_diff": float(np.abs(a.astype(np.float64) - b).max())})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--header", required=True, help="path of the real FastNoiseLite.h")
    ap.add_argument("--golden", default=None, help="where to save upstream's RandVecs3D (npy, 256 x 4 float32)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    subprocess.check_call(["make", "-s", "-j", str(os.cpu_count() or 4), "-C", os.path.join(ROOT, "oracle"), "ref"])
    subprocess.check_call(["make", "-s", "-j", str(os.cpu_count() or 4), "-C", os.path.join(ROOT, "oracle"), "ref", "FASTNOISE=" + os.path.abspath(a.header)])
    restated = load(os.path.join(ROOT, "oracle", "_ref", "libup_ref.so"))
    upstream = load(os.path.join(ROOT, "oracle", "_ref_upstream", "libup_ref.so"))
    assert restated.upref_fastnoise_is_upstream() == 0 and upstream.upref_fastnoise_is_upstream() == 1
    rows = compare(restated, upstream)
    worst = {}
    for r in rows:
        k = (r["type"], r["fractal"])
        w = worst.setdefault(k, {"bit_identical": 1.0, "max_abs_diff": 0.0})
        w["bit_identical"] = min(w["bit_identical"], r["bit_identical"])
        w["max_abs_diff"] = max(w["max_abs_diff"], r["max_abs_diff"])
    print("%-16s %-9s %14s %14s" % ("generator", "fractal", "bit-identical", "max |diff|"))
    for (t, f), w in worst.items():
        print("%-16s %-9s %13.4f%% %14.3g" % (t, f, 100 * w["bit_identical"], w["max_abs_diff"]))
    table_up, table_re = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
    upstream.upref_cellular_table(table_up.ctypes.data_as(C.c_void_p))
    restated.upref_cellular_table(table_re.ctypes.data_as(C.c_void_p))
    same_table = bool((table_up.view(np.uint32) == table_re.view(np.uint32)).all())
    print("RandVecs3D: upstream table %s the declared substitute" % ("EQUALS" if same_table else "differs from"))
    if a.golden:
        np.save(a.golden, table_up.reshape(256, 4))
        print("saved upstream's table to", a.golden, "-- Planet.set_cellular_table(np.load(...)) makes the Cellular layers follow upstream")
    differing = [k for k, w in worst.items() if w["bit_identical"] < 1.0 and k[0] != "cellular"]
    out = {"rows": rows, "randvecs3d_equal": same_table, "differing_generators": ["%s/%s" % k for k in differing]}
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    sys.exit(1 if differing else 0)


if __name__ == "__main__":
    main()
