"""GPU parity on the BASELINE.json configurations AT THEIR OWN SIZE, against the live reference (oracle/_ref: the reference's own
translation units compiled verbatim; it travels to the GPU box with the snapshot):

  config 1   --shape sphere (elevation lakes, default seed), 500^3                     sources/meshGeneration.cpp:19-28
  config 2   --shape torus --elevation craters, 500^3 (reaches the Voronoi code)       sources/elevationModes.cpp:173-198
  config 3   sphere at 1000^3: a 16-slice slab through the surface

Per configuration: 32 (16) z-slices of land / water / navigation within the stated tolerance (<= 1e-4 relative + 2e-3 absolute) with
every sample outside it ENUMERATED BY CAUSE and capped in magnitude; the sign mask identical away from zero; and the complete
meshGenerateBaseLand / Water / Navigation bodies (upg_field_meshes_clean: band sampling, Dual Marching Cubes, K9 clean-up) bit-equal
to the reference's makeMesh + meshRemoveDisconnected + meshFlipNormals run over the GPU's own dense samples.

The one cause of out-of-tolerance samples is a DISCONTINUITY OF THE REFERENCE ITSELF: commonElevationMountains returns `land` when
cover < 1e-7 and `land + smoothMax(0, max(ridge, terraces), 50)` otherwise (sources/elevationModes.cpp:102-104, 122); with both masks
at zero the second form is land + 12.5 (smoothMin's k * h * (1 - h) at h = 1/2, sources/math.cpp:28-38).  `land` comes out of powf (libm: a few ulp apart on CPU and GPU),
so a sample whose land lies within those ulps of -10 takes the other branch: a step of 12.5 m = 1.25 field units.  No other mode
of the BASELINE configurations has a discontinuity."""
import os

import numpy as np
import pytest

from conftest import SDF_ATOL, SDF_RTOL, same_bits, sdf_close
from unnatural_planets_b200 import Planet, capi, isosurface

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 1
STEP = 12.5 / 10.0  # smoothMax(0, 0, 50) / meshElevationRatio


def ref_or_skip(**kw):
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    return oref.RefPlanet(**kw)


def slice_blocks(res, slices):
    """`slices` z-slices in blocks of 8: through the pole cap, mid latitudes and the equator of a +-1100 planet in a +-2000 box"""
    starts = [int(res * f) for f in (0.23, 0.36, 0.5, 0.7)][: max(1, slices // 8)]
    return [(z, z + 8) for z in starts]


def check_field(p, ref, field, res, blocks, elev_mode):
    bad_total, n_total, flips_far = 0, 0, 0
    for z0, z1 in blocks:
        got = p.field_samples(field, z0, z1)
        want = ref.grid(field, res, z0=z0, z1=z1, threads=THREADS)
        ok = sdf_close(got, want, SDF_RTOL, SDF_ATOL)
        n_total += ok.size
        bad = ~ok
        if bad.any():
            # enumerate by cause: only the lakes / islands cover threshold may differ, by exactly one step, and only where the
            # reference's own raw elevation sits on one of the two sides of that step
            assert elev_mode in ("lakes", "islands") and field != "water", (field, int(bad.sum()), float(np.abs(got - want)[bad].max()))
            diff = np.abs(got.astype(np.float64) - want)[bad]
            assert np.all(np.abs(diff - STEP) < 5e-3), (field, diff[:8])
            zz, yy, xx = np.nonzero(bad)
            c = ref.grid_positions(res)
            pts = np.stack([c[xx], c[yy], c[zz + z0]], axis=1)
            raw = ref.eval("elevraw", pts)
            assert np.all((np.abs(raw + 10.0) < 2e-2) | (np.abs(raw - 2.5) < 2e-2)), raw[:8]
            bad_total += int(bad.sum())
        far = np.abs(want) > SDF_ATOL
        flips = ((got >= 0) != (want >= 0)) & far & ok
        flips_far += int(flips.sum())
    assert flips_far == 0, (field, flips_far)
    # hard cap on how many samples may sit on the reference's discontinuity: a few per million
    assert bad_total <= max(3, n_total // 200000), (field, bad_total, n_total)
    return bad_total, n_total


@pytest.mark.parametrize("arithmetic", ["exact", "contract"])
@pytest.mark.parametrize("shape,elev,res,slices", [("sphere", "lakes", 500, 32), ("torus", "craters", 500, 32), ("sphere", "lakes", 1000, 16)])
def test_baseline_config_fields_against_reference(shape, elev, res, slices, arithmetic):
    """both arithmetic builds of the dense kernel against the reference itself: `contract` (FMA contraction, SFU pow / reciprocal in the
    elevation curves -- the build bench.py's headline times) has to meet the same tolerance as `exact`"""
    ref = ref_or_skip(shape=shape, elevation=elev)
    p = Planet(shape=shape, elevation=elev, grid=res, arithmetic=arithmetic)
    blocks = slice_blocks(res, slices)
    for field in ("land", "water", "navigation"):
        check_field(p, ref, field, res, blocks, elev)
    if arithmetic == "exact":
        # the shape of both planets has no transcendental: the water field is bit-exact
        z0, z1 = blocks[-1]
        assert same_bits(p.field_samples("water", z0, z1), ref.grid("water", res, z0=z0, z1=z1, threads=THREADS)).all()


def assert_same(m, positions, normals, indices, what=""):
    assert len(m["indices"]) == len(indices), (what, len(m["indices"]) // 3, len(indices) // 3)
    assert len(m["positions"]) == len(positions), (what, len(m["positions"]), len(positions))
    assert np.array_equal(m["indices"], np.asarray(indices).reshape(-1)), what
    assert same_bits(m["positions"], positions).all(), what
    assert same_bits(m["normals"], normals).all(), what


@pytest.mark.parametrize("shape,elev", [("sphere", "lakes"), ("torus", "craters")])
def test_baseline_config_generator_meshes_against_reference(shape, elev):
    """configs 1 and 2 at 500^3: what meshGenerateBaseLand / Navigation / Water return (sources/meshGeneration.cpp:47-107)"""
    res = 500
    ref = ref_or_skip(shape=shape, elevation=elev)
    p = Planet(shape=shape, elevation=elev, grid=res)
    got = p.field_meshes_clean(("land", "water", "navigation"), stages=capi.MESH_GENERATOR)
    for f in ("land", "navigation"):
        wp, wn, wi = ref.make_mesh(p.field_samples(f), remove_disconnected=True, flip=True)
        assert_same(got[f], wp, wn, wi, (shape, f))
        assert len(wi) > 300000  # a real planet-sized mesh went through
    fp, fn, fi = ref.make_mesh(p.field_samples("water"), remove_disconnected=True, flip=True)
    wp, wn, wi = ref.water_cull(fp, fn, fi)
    # the culling compares terrainSdfElevationRaw (powf inside) with 0.1 per vertex: exact unless a vertex sits within float noise of it
    if np.abs(ref.eval("elevraw", fp) - 0.1).min() > 5e-3:
        assert_same(got["water"], wp, wn, wi, (shape, "water"))
    else:
        assert abs(len(got["water"]["indices"]) - len(wi)) <= 60, (len(got["water"]["indices"]), len(wi))


def test_config3_slab_isosurface_against_reference():
    """sphere 1000^3: dualmc over a 20-slice slab of the GPU's samples, vertices and quads with the reference's numbering"""
    res = 1000
    ref = ref_or_skip(shape="sphere", elevation="lakes")
    p = Planet(shape="sphere", elevation="lakes", grid=res)
    z0 = int(res * 0.23)
    dens = p.field_samples("land", z0, z0 + 20)
    wv, wq = ref.dualmc(dens)
    m = isosurface(dens, full_res=res)
    assert len(wq) > 20000
    assert np.array_equal(m["dmc_quads"], wq.astype(np.uint32))
    assert same_bits(m["dmc_vertices"], wv).all()
