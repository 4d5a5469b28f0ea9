"""CPU tests of the product's device headers, compiled for the host (tests/hostlib), against golden vectors
generated from the reference itself (tests/golden/make_golden.py).  Everything that does not call a libm
transcendental with a run-time argument must be BIT-EXACT: same source order, no FMA contraction."""
import numpy as np
import pytest

from conftest import HostPlanet, golden, golden_registry, same_bits
from unnatural_planets_b200.planet import ELEVATIONS, SHAPES

COMBOS = [("sphere", e) for e in ELEVATIONS] + [("torus", "lakes"), ("torus", "craters"), ("asteroid", "islands"), ("fibers", "simple"), ("bunny", "legacy"), ("knot", "craters")]


def test_grid_coordinates_reference_kat(hostlib):
    # externals/cage/sources/test-core/marchingCubes.cpp:27-90: box [0,1], resolution 9 -> -0.5, -0.25, 0 ... 1.5
    for res, first, step in ((9, -0.5, 0.25), (8, -2.0 / 3, 1.0 / 3), (7, -1.0, 0.5), (6, -2.0, 1.0)):
        # box [0,1] is box size 1 centred at 0.5: evaluate through a box of size 1 and shift
        hp = HostPlanet(hostlib, "sphere", "none", res=res, box=1.0)
        c = hp.coords() + np.float32(0.5)
        want = first + step * np.arange(res)
        assert np.allclose(c, want, atol=1e-6), (res, c, want)


@pytest.mark.parametrize("res", [6, 7, 8, 9, 110, 500, 1000, 1024])
def test_grid_coordinates_bit_exact(hostlib, res):
    want = golden("misc.npz")["grid_positions_%d" % res]
    got = HostPlanet(hostlib, "sphere", "none", res=res).coords()
    assert same_bits(got, want).all()


@pytest.mark.parametrize("shape", SHAPES)
def test_shape_sdf_bit_exact(hostlib, shape):
    pts = golden("probe_points.npy")
    want = golden("probe_shapes.npz")[shape]
    got = HostPlanet(hostlib, shape, "none").points("water", pts)
    ok = same_bits(got, want)
    if shape == "mandelbulb":
        # powf(r, 2) is folded to r*r by the host compiler in this build only; <= 1 ulp of the result scale
        assert np.abs(got - want).max() < 1e-4
        assert ok.mean() > 0.99
    else:
        assert ok.all(), (shape, int((~ok).sum()), float(np.abs(got - want)[~ok].max()))


@pytest.mark.parametrize("shape,elev", COMBOS)
def test_fields_bit_exact(hostlib, shape, elev):
    g = golden("probe_fields_%s_%s.npz" % (shape, elev))
    hp = HostPlanet(hostlib, shape, elev)
    for f in ("water", "elevraw", "land", "navigation", "elevation"):
        got = hp.points(f, g["points"])
        ok = same_bits(got, g[f])
        assert ok.all(), (shape, elev, f, int((~ok).sum()))


@pytest.mark.parametrize("shape,elev", COMBOS)
def test_seed_order_matches_reference_registry(hostlib, shape, elev):
    # the noise objects the reference constructed during terrainPreseed, in construction order
    g = golden("probe_fields_%s_%s.npz" % (shape, elev))
    reg = golden_registry(g)
    log = HostPlanet(hostlib, shape, elev).seed_log()
    n_shape = 1 if shape == "asteroid" else 0
    n_elev = {"none": 0, "simple": 1, "legacy": 2, "lakes": 4, "islands": 4, "craters": 2}[elev]
    # ... followed by the 40 draws of coloringDefault (sources/coloringModes.cpp:905-930 through terrainPreseed)
    assert len(log) == n_shape + n_elev + 40
    ref_seeds = [np.uint32(r["seed"] & 0xFFFFFFFF) for r in reg]
    b = n_shape + n_elev
    if elev == "craters":
        # the first elevation seed goes to the Voronoi object (not a NoiseFunction): registry holds only the second
        assert log[n_shape + 1] == ref_seeds[n_shape]
        if n_shape:
            assert log[0] == ref_seeds[0]
        rb = b - 1
    else:
        assert list(log[:b]) == ref_seeds[:b]
        rb = b
    # colouring: draws that are not NoiseFunctions (three Voronoi seeds) are absent from the registry, the three `static
    # const uint32 seed` draws (bedrock, grass, ice) show up later, twice each, on the Cellular objects that share them
    c = [int(x) for x in log[b:]]
    expect = c[0:3] + [c[4], c[5], c[3], c[3], c[6]] + c[7:17] + [c[18], c[17], c[17], c[19], c[20], c[21]] + c[23:27] + c[29:33] + [c[34], c[35], c[33], c[33]] + c[36:40]
    assert [int(x) for x in ref_seeds[rb:]] == expect
    # the thread generator itself is only advanced by the fibers (3 draws) and asteroid (1 draw) statics
    assert tuple(int(x) for x in g["rng_state_after_preseed"]) != (0, 0)


def test_full_grid_matches_reference(hostlib):
    for name, shape, elev, res in (("grid_sphere_lakes_40.npz", "sphere", "lakes", 40), ("grid_torus_craters_40.npz", "torus", "craters", 40)):
        g = golden(name)
        got = HostPlanet(hostlib, shape, elev, res=res).grid("land")
        assert same_bits(got, g["land"]).all()


@pytest.mark.parametrize("elev", ELEVATIONS)
def test_elevation_interval_bounds_every_mode(hostlib, elev):
    """elevation_interval10 (csrc/upg_planet_host.h) is what the band sampler trusts: elevation / 10 of every mode must lie inside
    it, for any seed and anywhere in (and well beyond) the meshing box; the hull contains 0 so it also bounds max(elevation, 0) / 10."""
    import ctypes as C

    lo, hi = C.c_double(), C.c_double()
    hostlib.upgh_elevation_interval10(ELEVATIONS.index(elev), C.byref(lo), C.byref(hi))
    lo, hi = lo.value, hi.value
    assert lo <= 0 <= hi
    rng = np.random.default_rng(12345)
    seen_lo, seen_hi = np.inf, -np.inf
    for k in range(6):
        seed = (int(rng.integers(1, 2**62)), int(rng.integers(1, 2**62)))
        hp = HostPlanet(hostlib, "sphere", elev, seed=seed)
        pts = rng.uniform(-2600, 2600, size=(60000, 3)).astype(np.float32)
        e10 = hp.points("elevraw", pts).astype(np.float64) / 10
        assert np.isfinite(e10).all()
        seen_lo, seen_hi = min(seen_lo, e10.min()), max(seen_hi, e10.max())
    assert lo <= seen_lo and seen_hi <= hi, (elev, lo, seen_lo, seen_hi, hi)
    # the interval is a bound, not a guess: it must not sit orders of magnitude away from what the mode produces
    if elev != "none":
        assert hi - lo < 12 * (seen_hi - seen_lo), (elev, lo, seen_lo, seen_hi, hi)


@pytest.mark.parametrize("elev,res", [("lakes", 40), ("islands", 33), ("legacy", 28), ("craters", 12), ("lakes", 1000)])
def test_column_evaluation_is_bit_identical(hostlib, elev, res):
    """the dense kernel walks the grid in z columns and keeps the first Value fBm of lakes / islands / legacy between samples
    (ValueFbmColumn, csrc/upg_noise.cuh): same bits as the sample-by-sample evaluation, in any z order"""
    import ctypes as C

    hp = HostPlanet(hostlib, "sphere", elev, res=res)
    if res > 100:
        # a slab of the bench grid, where an octave stays up to 192 samples between two lattice planes; two grid rows of every slice
        z0, z1, rows = 300, 700, 2
        c = hp.coords()
        zz, yy, xx = np.meshgrid(c[z0:z1], c[:rows], c, indexing="ij")
        want = hp.points("elevraw", np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)).reshape(z1 - z0, rows, res)
        got = np.zeros((z1 - z0, res, res), dtype=np.float32)
        hostlib.upgh_grid_elevation_columns_rows(C.c_void_p(hp.h), z0, z1, rows, got.ctypes.data_as(C.c_void_p))
        assert same_bits(got[:, :rows, :], want).all()
        return
    want = hp.grid("elevraw")
    for order in (0, 1, 2):
        got = np.zeros((res, res, res), dtype=np.float32)
        hostlib.upgh_grid_elevation_columns(C.c_void_p(hp.h), 0, res, order, got.ctypes.data_as(C.c_void_p))
        assert same_bits(got, want).all(), (elev, order)


def test_rng_known_answers(hostlib):
    # externals/cage/sources/test-core/random.cpp:37-79 pins randomRange3i(10, 1000) == (730, 955, 866) for this state;
    # the raw stream below was produced by cage's RandomGenerator itself
    m = golden("misc.npz")
    assert list(m["rng_range3i"]) == [730, 955, 866]
    s = [1346519564496, 42245614964156]
    mask = (1 << 64) - 1

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & mask

    out = []
    for _ in range(64):
        s0, s1 = s
        out.append((rotl((s0 * 5) & mask, 7) * 9) & mask)
        s1 ^= s0
        s = [rotl(s0, 24) ^ s1 ^ ((s1 << 16) & mask), rotl(s1, 37)]
    assert out == [int(x) for x in m["rng_raw"]]
    # and the product's own stream (seed log of a lakes planet with the default colouring: 4 + 40 draws) is the low 32 bits
    log = HostPlanet(hostlib, "sphere", "lakes").seed_log()
    assert [int(x) for x in log] == [x & 0xFFFFFFFF for x in out[:44]]
