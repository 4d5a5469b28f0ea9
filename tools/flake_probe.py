#!/usr/bin/env python3
"""Runs the mesh entry points (band sampling, three concurrent extractions, K9 clean-up) on a few planets and prints a digest of every
mesh: under compute-sanitizer --tool initcheck this shows reads of scratch memory nobody wrote; under two different UPG_POISON bytes the
digests must agree.   python tools/flake_probe.py [--res 130] [--repeat 1]"""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=130)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--dump", default=None, help="directory: on a mismatch, save the reference mesh and the deviating one")
ap.add_argument("--planets", default="sphere:lakes,torus:craters,doubletorus:islands,mandelbulb:simple")
a = ap.parse_args()
digests = {}
first = {}
for it in range(a.repeat):
    for spec in a.planets.split(","):
        shape, elev = spec.split(":")
        for sampling in ("band", "dense"):
            p = Planet(shape=shape, elevation=elev, grid=a.res, mesh_sampling=sampling)
            m = p.field_meshes_clean(("land", "water", "navigation"))
            h = hashlib.sha1()
            for f in ("land", "water", "navigation"):
                for k in ("positions", "normals", "indices"):
                    h.update(m[f][k].tobytes())
            key = "%s:%s" % (spec, sampling)
            d = h.hexdigest()[:16]
            if key in digests and digests[key] != d:
                print("MISMATCH", key, "iteration", it, digests[key], d)
                import numpy as np

                if a.dump:
                    os.makedirs(a.dump, exist_ok=True)
                    tag = key.replace(":", "_") + "_it%d" % it
                    np.savez_compressed(os.path.join(a.dump, tag + ".npz"), **{"%s_%s_%s" % (w, f, k): (first[key] if w == "ref" else m)[f][k] for w in ("ref", "bad") for f in ("land", "water", "navigation") for k in ("positions", "normals", "indices")})
                for f in ("land", "water", "navigation"):
                    for k in ("positions", "normals", "indices"):
                        x, y = first[key][f][k], m[f][k]
                        if x.shape != y.shape:
                            print("   ", f, k, "shape", x.shape, "->", y.shape)
                        else:
                            ne = np.flatnonzero(x.reshape(-1).view(np.uint32) != y.reshape(-1).view(np.uint32))
                            if len(ne):
                                print("   ", f, k, len(ne), "elements differ, first at", ne[:6].tolist(), "values", x.reshape(-1)[ne[:3]].tolist(), "->", y.reshape(-1)[ne[:3]].tolist())
            else:
                digests[key] = d
                first.setdefault(key, m)
            p.close()
print(json.dumps(digests, sort_keys=True))
