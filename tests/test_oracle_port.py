"""Pins the oracle restatement (oracle/up_oracle.cpp + up_noise.h): bit-for-bit against the golden vectors that the
reference's own translation units produced (tests/golden/make_golden.py), and -- where oracle/_ref is present --
against the reference library directly on fresh random points."""
import numpy as np
import pytest

from conftest import golden, golden_registry, same_bits
from oracle import port, ref

COMBOS = [("sphere", e) for e in port.ELEVATIONS] + [("torus", "lakes"), ("torus", "craters"), ("asteroid", "islands"), ("fibers", "simple"), ("bunny", "legacy"), ("knot", "craters")]


@pytest.mark.parametrize("shape", port.SHAPES)
def test_shapes_against_golden(shape):
    pts = golden("probe_points.npy")
    want = golden("probe_shapes.npz")[shape]
    got = port.OraclePlanet(shape, "none").points("water", pts)
    assert same_bits(got, want).all()


@pytest.mark.parametrize("shape,elev", COMBOS)
def test_fields_against_golden(shape, elev):
    g = golden("probe_fields_%s_%s.npz" % (shape, elev))
    p = port.OraclePlanet(shape, elev)
    for f in ("water", "elevraw", "land", "navigation", "elevation"):
        assert same_bits(p.points(f, g["points"]), g[f]).all(), f
    reg = golden_registry(g)
    log = p.seed_log()
    noise_seeds = [r["seed"] & 0xFFFFFFFF for r in reg]
    mine = [int(x) for x in log]
    if elev == "craters":
        mine = mine[:-2] + mine[-1:]  # the Voronoi seed is not a NoiseFunction
    assert mine == noise_seeds[: len(mine)]


@pytest.mark.parametrize("name,shape,elev,res", [("grid_sphere_lakes_40.npz", "sphere", "lakes", 40), ("grid_torus_craters_40.npz", "torus", "craters", 40), ("grid_mandelbulb_none_36.npz", "mandelbulb", "none", 36)])
def test_grid_against_golden(name, shape, elev, res):
    g = golden(name)
    got = port.OraclePlanet(shape, elev).grid("land", res, threads=2)
    assert same_bits(got, g["land"]).all()


@pytest.mark.parametrize("res", [6, 7, 8, 9, 110, 500, 1000, 1024])
def test_grid_coordinates(res):
    assert same_bits(port.grid_coords(res), golden("misc.npz")["grid_positions_%d" % res]).all()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref needs the reference checkout")
@pytest.mark.parametrize("shape,elev", [("sphere", "lakes"), ("torus", "craters"), ("hemispheres", "simple"), ("twistedplane", "legacy"), ("monkeyhead", "islands")])
def test_against_reference_library(shape, elev):
    rng = np.random.default_rng(99)
    pts = ((rng.random((4000, 3), dtype=np.float32) * 2 - 1) * 2000).astype(np.float32)
    r = ref.RefPlanet(shape=shape, elevation=elev)
    o = port.OraclePlanet(shape, elev)
    for f in ("water", "elevraw", "land", "navigation"):
        assert same_bits(o.points(f, pts), r.eval(f, pts)).all(), f


def test_noise_structure():
    """FastNoiseLite restatement: range and continuity of every generator (upstream source is absent -> values unpinned)"""
    import ctypes as C

    L = port.lib()
    L.upo_noise_eval.restype = C.c_float

    class Cfg(C.Structure):
        _fields_ = [("type", C.c_int32), ("fractal", C.c_int32), ("distance", C.c_int32), ("operation", C.c_int32), ("seed", C.c_int32), ("octaves", C.c_int32), ("lacunarity", C.c_float), ("gain", C.c_float), ("frequency", C.c_float), ("bounding", C.c_float)]

    L.upo_noise_eval.argtypes = [C.POINTER(Cfg), C.c_float, C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    for ntype in (0, 1, 3, 4, 5):  # simplex, simplex-reduced, perlin, cubic, value
        cfg = Cfg(ntype, 0, 0, 0, 1337, 3, 2.0, 0.5, 1.0, 1.0)
        base = rng.random((3000, 3)) * 40 - 20
        step = (rng.random((3000, 3)) - 0.5) * 2e-3
        a = np.array([L.upo_noise_eval(C.byref(cfg), *map(float, p)) for p in base])
        b = np.array([L.upo_noise_eval(C.byref(cfg), *map(float, p)) for p in base + step])
        assert np.isfinite(a).all() and np.abs(a).max() <= 1.0 + 1e-3, (ntype, np.abs(a).max())
        assert a.std() > 0.1, ntype  # not degenerate
        # continuity across lattice cells: |f(p + d) - f(p)| bounded by a Lipschitz constant times |d|
        assert np.abs(a - b).max() < 0.02, (ntype, np.abs(a - b).max())


def test_texture_tail_restatement_matches_the_reference_fixture():
    # the numpy restatement of imageConvertSpecialToGltfPbr / imageConvertHeigthToNormal is pinned on the reference's own outputs
    import bake_common as bc
    from conftest import golden

    g = golden("texture_tail.npz")
    for name, sp, hm in bc.tail_images():
        pbr, nm = port.texture_tail(sp, hm, 1.0)
        assert (pbr == g[name + "_pbr"]).all(), name
        assert (nm == g[name + "_normal"]).all(), name
    name, sp, hm = bc.tail_images()[-1]
    assert (port.texture_tail(None, hm, 4.0)[1] == g["strength4_normal"]).all()
