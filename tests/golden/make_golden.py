#!/usr/bin/env python3
"""Regenerates the golden fixtures in tests/golden/ from the REFERENCE ITSELF (oracle/_ref/libup_ref.so: the
reference's translation units compiled verbatim, see oracle/Makefile).  Runs only where /root/reference exists;
the fixtures it writes are committed so that the test-suite (CPU and GPU box) never needs the reference.

    python tests/golden/make_golden.py

Fixtures
  probe_points.npy                    (N,3) float32 probe positions (seeded; half of them within +-60 of some surface)
  probe_shapes.npz                    one array per shape: terrainSdfWater at the probe points (elevation 'none')
  probe_fields_<shape>_<elev>.npz     water / elevraw / land / navigation / elevation at the probe points + noise registry
  grid_<shape>_<elev>_<res>.npz       full res^3 land grid, raw dualmc output, makeMesh output, final mesh
  misc.npz                            RNG known answers, grid coordinates, colour conversions, dilation
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import DEFAULT_SEED, RefPlanet  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SHAPES = ("asteroid", "belt", "bowl", "box", "bunny", "capsule", "cube", "disk", "doubletorus", "fibers", "gear", "h2o", "h3o", "h4o", "hemispheres", "hexagon", "hexagonalprism", "insidecube", "knot", "mandelbulb", "mobiusstrip", "monkeyhead", "octahedron", "pipe", "sphere", "square", "tetrahedron", "torus", "toruscross", "triangularprism", "tube", "twistedhexagonalprism", "twistedplane", "wormhole")
ELEVATIONS = ("none", "simple", "legacy", "lakes", "islands", "craters")
N_PROBE = 3000


def probe_points():
    rng = np.random.default_rng(20261017)
    pts = ((rng.random((N_PROBE, 3), dtype=np.float32) * 2 - 1) * 2100).astype(np.float32)
    pts[:64] *= np.float32(0.01)  # near the origin (asteroid / mandelbulb / neural early-outs)
    pts[64:72] = 0
    pts[72] = (1100, 0, 0)
    pts[73] = (0, -1100, 0)
    # exact grid nodes of the default 500^3 grid
    g = RefPlanet(init=False).grid_positions(500)
    idx = rng.integers(0, 500, size=(256, 3))
    pts[80:336] = np.stack([g[idx[:, 0]], g[idx[:, 1]], g[idx[:, 2]]], axis=1)
    return pts


def near_surface(ref, pts, rng, count):
    """replace the tail of the probe set by points close to this planet's land surface (sign-critical samples)"""
    cand = ((rng.random((count * 60, 3), dtype=np.float32) * 2 - 1) * 2000).astype(np.float32)
    v = ref.eval("land", cand)
    order = np.argsort(np.abs(v))[:count]
    out = pts.copy()
    out[-count:] = cand[order]
    return out


def cleanup_volume(seed=5, res=22):
    """a dense random volume (every dualmc configuration, many components, non-manifold edges, a surface that leaves the clip
    box everywhere) smoothed once so that components are more than single cells"""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((res + 2, res + 2, res + 2)).astype(np.float32)
    v = (v[1:-1, 1:-1, 1:-1] + v[2:, 1:-1, 1:-1] + v[:-2, 1:-1, 1:-1] + v[1:-1, 2:, 1:-1] + v[1:-1, :-2, 1:-1] + v[1:-1, 1:-1, 2:] + v[1:-1, 1:-1, :-2]).astype(np.float32)
    return np.ascontiguousarray(v)


def cleanup_fixture():
    """mesh_cleanup.npz: the clean-up chain (makeMesh tail, meshRemoveDisconnected, meshFlipNormals, water culling) on inputs the
    planet grids do not reach; python tests/golden/make_golden.py cleanup = only this"""
    d = {}
    g = cleanup_volume()
    ref = RefPlanet(init=False)
    d["volume"] = g
    d["mesh_positions"], d["mesh_normals"], d["mesh_indices"] = ref.make_mesh(g)
    d["final_positions"], d["final_normals"], d["final_indices"] = ref.make_mesh(g, remove_disconnected=True, flip=True)
    # the water mesh of a small planet through meshGenerateBaseWater's culling
    ref = RefPlanet(shape="sphere", elevation="lakes")
    w = ref.grid("water", 44)
    fp, fn, fi = ref.make_mesh(w, remove_disconnected=True, flip=True)
    d["water_grid"] = w
    d["water_positions"], d["water_normals"], d["water_indices"] = ref.water_cull(fp, fn, fi)
    d["water_elevation_raw"] = ref.eval("elevraw", fp)
    d["water_before_cull_vertices"] = np.array([len(fp)], dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "mesh_cleanup.npz"), **d)
    print("cleanup: volume mesh tris", len(d["mesh_indices"]) // 3, "final", len(d["final_indices"]) // 3, "water", len(fi) // 3, "->", len(d["water_indices"]) // 3)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "cleanup":
        return cleanup_fixture()
    cleanup_fixture()
    pts = probe_points()
    np.save(os.path.join(OUT, "probe_points.npy"), pts)

    # 1. every shape, elevation none: the base-shape SDF alone
    shapes = {}
    for s in SHAPES:
        ref = RefPlanet(shape=s, elevation="none")
        shapes[s] = ref.eval("water", pts)
        print("shape", s, float(np.nanmin(shapes[s])), float(np.nanmax(shapes[s])))
    np.savez_compressed(os.path.join(OUT, "probe_shapes.npz"), **shapes)

    # 2. field compositions for the BASELINE configs + one planet per elevation mode + the RNG-dependent shapes
    combos = [("sphere", e) for e in ELEVATIONS] + [("torus", "lakes"), ("torus", "craters"), ("asteroid", "islands"), ("fibers", "simple"), ("bunny", "legacy"), ("knot", "craters")]
    rng = np.random.default_rng(7)
    for s, e in combos:
        ref = RefPlanet(shape=s, elevation=e)
        p2 = near_surface(ref, pts, rng, 1000)
        d = {"points": p2}
        for f in ("water", "elevraw", "land", "navigation", "elevation"):
            d[f] = ref.eval(f, p2)
        reg = ref.noise_configs()
        d["registry"] = np.frombuffer(json.dumps(reg).encode(), dtype=np.uint8)
        d["rng_state_after_preseed"] = np.array(ref.rng_state(), dtype=np.uint64)
        np.savez_compressed(os.path.join(OUT, "probe_fields_%s_%s.npz" % (s, e)), **d)
        print("fields", s, e, "noise objects:", len(reg))

    # 3. whole small grids through the reference's MarchingCubes + dualmc + makeMesh
    for s, e, res in (("sphere", "lakes", 40), ("torus", "craters", 40), ("mandelbulb", "none", 36)):
        ref = RefPlanet(shape=s, elevation=e)
        g = ref.grid("land", res)
        v, q = ref.dualmc(g, 0.0, True)
        v0, q0 = ref.dualmc(g, 0.0, False)
        mp, mn, mi = ref.make_mesh(g)
        fp, fn, fi = ref.make_mesh(g, remove_disconnected=True, flip=True)
        np.savez_compressed(os.path.join(OUT, "grid_%s_%s_%d.npz" % (s, e, res)), land=g, dmc_vertices=v, dmc_quads=q, dmc_vertices_nomanifold=v0, dmc_quads_nomanifold=q0, mesh_positions=mp, mesh_normals=mn, mesh_indices=mi, final_positions=fp, final_normals=fn, final_indices=fi)
        print("grid", s, e, res, "verts", len(v), "quads", len(q), "mesh tris", len(mi) // 3, "final tris", len(fi) // 3)

    # 4. miscellaneous known answers
    ref = RefPlanet(init=False)
    misc = {}
    misc["rng_raw"] = ref.rng_raw(DEFAULT_SEED[0], DEFAULT_SEED[1], 64)
    f3, i3 = ref.rng_range3(DEFAULT_SEED[0], DEFAULT_SEED[1], 10, 1000, 10, 1000)
    misc["rng_range3_after_range3i"] = f3
    misc["rng_range3i"] = i3
    for r in (6, 7, 8, 9, 110, 500, 1000, 1024):
        misc["grid_positions_%d" % r] = ref.grid_positions(r)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **misc)
    print("done")


if __name__ == "__main__":
    main()
