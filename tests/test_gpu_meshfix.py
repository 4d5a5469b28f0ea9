"""GPU parity of K9 (the mesh clean-up after the isosurface: removeNonManifoldTriangles, meshClip incl. its merge / invalid tail,
meshRemoveDisconnected, meshFlipNormals, the water-vertex culling) THROUGH THE C ABI.  Everything here is integer work or
+,-,*,/,sqrt in the reference's order: vertex order, triangle order, positions and normals are asserted BIT-EXACT against
fixtures the reference itself produced (tests/golden/make_golden.py) and, where oracle/_ref travelled, against live runs."""
import numpy as np
import pytest

from conftest import golden, same_bits
from unnatural_planets_b200 import Planet, UpgError, capi, isosurface, mesh_clean

pytestmark = pytest.mark.gpu

GRIDS = ["grid_sphere_lakes_40.npz", "grid_torus_craters_40.npz", "grid_mandelbulb_none_36.npz"]


def assert_same(m, positions, normals, indices, what=""):
    assert len(m["indices"]) == len(indices), (what, len(m["indices"]) // 3, len(indices) // 3)
    assert len(m["positions"]) == len(positions), (what, len(m["positions"]), len(positions))
    assert np.array_equal(m["indices"], np.asarray(indices).reshape(-1)), what
    assert same_bits(m["positions"], positions).all(), what
    assert same_bits(m["normals"], normals).all(), what


def ref_or_skip(**kw):
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    return oref.RefPlanet(**kw)


@pytest.mark.parametrize("name", GRIDS)
def test_clean_against_reference_fixtures(name):
    # mandelbulb leaves the +-2000 box: its fixture exercises the real clipping (new vertices, cut triangles)
    g = golden(name)
    raw = isosurface(g["land"])
    m = mesh_clean(raw["positions"], raw["normals"], raw["indices"], stages=capi.MESH_MAKEMESH_TAIL)
    assert_same(m, g["mesh_positions"], g["mesh_normals"], g["mesh_indices"], "makeMesh")
    m = mesh_clean(raw["positions"], raw["normals"], raw["indices"], stages=7)
    assert_same(m, g["final_positions"], g["final_normals"], g["final_indices"], "final")
    # the later stages alone, applied to the reference's own makeMesh output
    m = mesh_clean(g["mesh_positions"], g["mesh_normals"], g["mesh_indices"], stages=capi.MESH_REMOVE_DISCONNECTED | capi.MESH_FLIP_NORMALS)
    assert_same(m, g["final_positions"], g["final_normals"], g["final_indices"], "tail only")


def test_clean_random_volume_fixture():
    # every dualmc configuration, hundreds of components, non-manifold edges, a surface cut by all six box planes
    g = golden("mesh_cleanup.npz")
    raw = isosurface(g["volume"])
    m = mesh_clean(raw["positions"], raw["normals"], raw["indices"], stages=1)
    assert_same(m, g["mesh_positions"], g["mesh_normals"], g["mesh_indices"], "makeMesh")
    m = mesh_clean(raw["positions"], raw["normals"], raw["indices"], stages=7)
    assert_same(m, g["final_positions"], g["final_normals"], g["final_indices"], "final")


@pytest.mark.parametrize("seed,res", [(1, 12), (2, 19), (3, 30)])
def test_clean_random_volumes_live(seed, res):
    import sys, os

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import cleanup_volume

    ref = ref_or_skip(init=False)
    v = cleanup_volume(seed, res)
    raw = isosurface(v)
    for stages, kw in ((1, {}), (3, dict(remove_disconnected=True)), (7, dict(remove_disconnected=True, flip=True))):
        wp, wn, wi = ref.make_mesh(v, **kw)
        assert_same(mesh_clean(raw["positions"], raw["normals"], raw["indices"], stages=stages), wp, wn, wi, (seed, stages))


@pytest.mark.parametrize("shape,elev,res", [("belt", "lakes", 56), ("bowl", "simple", 48), ("hemispheres", "craters", 50), ("hexagon", "lakes", 44), ("square", "legacy", 40), ("tube", "islands", 52), ("twistedplane", "lakes", 48), ("wormhole", "none", 46), ("mandelbulb", "lakes", 40), ("sphere", "lakes", 64), ("fibers", "lakes", 60)])
def test_generator_meshes_live(shape, elev, res):
    # the complete meshGenerateBaseLand / Navigation bodies for planets that cross the clip box (and two that do not): the GPU's own
    # samples go through the reference's makeMesh + meshRemoveDisconnected + meshFlipNormals
    ref = ref_or_skip(shape=shape, elevation=elev)
    p = Planet(shape=shape, elevation=elev, grid=res)
    got = p.field_meshes_clean(("land", "navigation"), stages=7)
    for f in ("land", "navigation"):
        wp, wn, wi = ref.make_mesh(p.field_samples(f), remove_disconnected=True, flip=True)
        assert_same(got[f], wp, wn, wi, (shape, f))
        assert got[f]["dmc_quads"].size == 0 and got[f]["dmc_vertices"].size == 0


def test_water_culling_against_fixture():
    g = golden("mesh_cleanup.npz")
    p = Planet(shape="sphere", elevation="lakes", grid=44)
    assert same_bits(p.field_samples("water"), g["water_grid"]).all()
    got = p.field_meshes_clean(("water",), stages=capi.MESH_GENERATOR)["water"]
    # the culling threshold (elevation < 0.1) sees powf: exact unless a vertex sits within float noise of it
    raw = p.field_meshes_clean(("water",), stages=7)["water"]
    assert len(raw["positions"]) == int(g["water_before_cull_vertices"][0])
    e = g["water_elevation_raw"]
    assert np.abs(p.field_points("elevation_raw", raw["positions"]) - e).max() < 2e-3
    if np.abs(e - 0.1).min() > 5e-3:
        assert_same(got, g["water_positions"], g["water_normals"], g["water_indices"], "water")
    else:
        assert abs(len(got["indices"]) - len(g["water_indices"])) <= 30
    assert 0 < len(got["indices"]) < len(raw["indices"])


def test_water_culling_live():
    ref = ref_or_skip(shape="torus", elevation="islands")
    p = Planet(shape="torus", elevation="islands", grid=60)
    got = p.field_meshes_clean(("water",), stages=capi.MESH_GENERATOR)["water"]
    fp, fn, fi = ref.make_mesh(p.field_samples("water"), remove_disconnected=True, flip=True)
    wp, wn, wi = ref.water_cull(fp, fn, fi)
    if np.abs(ref.eval("elevraw", fp) - 0.1).min() > 5e-3:
        assert_same(got, wp, wn, wi, "water live")
    else:
        assert abs(len(got["indices"]) - len(wi)) <= 30


def test_clean_edge_cases_and_errors():
    z3, zi = np.zeros((0, 3), np.float32), np.zeros(0, np.uint32)
    m = mesh_clean(z3, z3, zi, stages=7)
    assert len(m["positions"]) == 0 and len(m["indices"]) == 0
    # one triangle inside the box survives all stages; flipped winding and normals
    pos = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [5, 5, 5]], np.float32)  # vertex 3 is unused
    nrm = np.array([[0, 0, 1]] * 4, np.float32)
    m = mesh_clean(pos, nrm, np.array([0, 1, 2], np.uint32), stages=7)
    assert np.array_equal(m["indices"], [0, 2, 1]) and len(m["positions"]) == 3 and (m["normals"][:, 2] == -1).all()
    # the same directed edge twice: both faces are singular and go (marchingCubes.cpp:66-76)
    pos = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [50, 50, 50], [60, 50, 50], [50, 60, 50]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (7, 1))
    m = mesh_clean(pos, nrm, np.array([0, 1, 2, 0, 1, 3, 4, 5, 6], np.uint32), stages=1)
    assert np.array_equal(m["indices"], [0, 1, 2]) and same_bits(m["positions"], pos[4:]).all()
    # coincident vertices merge onto the smaller id; a triangle fully outside the box disappears
    pos = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [10, 0, 0], [10, 10, 0], [0, 10, 0], [3000, 0, 0], [3000, 10, 0], [3000, 0, 10]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (9, 1))
    m = mesh_clean(pos, nrm, np.array([0, 1, 2, 3, 4, 5, 6, 7, 8], np.uint32), stages=1)
    assert np.array_equal(m["indices"], [0, 1, 2, 1, 3, 2]) and len(m["positions"]) == 4
    with pytest.raises(UpgError):
        mesh_clean(pos, nrm, np.array([0, 1], np.uint32), stages=1)
    with pytest.raises(UpgError):
        mesh_clean(pos, nrm, np.array([0, 1, 99], np.uint32), stages=1)
    with pytest.raises(UpgError):
        mesh_clean(pos, nrm, np.array([0, 1, 2], np.uint32), stages=8)
    # an empty land mesh is the reference's exception, an empty water mesh is a result
    p = Planet(shape="sphere", elevation="none", grid=16, box=100.0)
    with pytest.raises(UpgError) as e:
        p.field_meshes_clean(("land",), stages=7)
    assert e.value.code == 4
    assert len(p.field_meshes_clean(("water",), stages=15)["water"]["indices"]) == 0


def test_default_size_planet_generator_meshes():
    # 500^3 sphere: the clean-up leaves a closed planet untouched but for the flip, and costs little
    p = Planet(shape="sphere", elevation="lakes", grid=500)
    raw = p.field_meshes(("land",))["land"]
    got = p.field_meshes_clean(("land",), stages=7)["land"]
    assert len(got["positions"]) <= len(raw["positions"]) and len(got["indices"]) <= len(raw["indices"])
    if len(got["positions"]) == len(raw["positions"]):
        assert same_bits(got["positions"], raw["positions"]).all() and same_bits(got["normals"], -raw["normals"]).all()
        assert np.array_equal(got["indices"].reshape(-1, 3)[:, [0, 2, 1]], raw["indices"].reshape(-1, 3))
    # one connected component: every vertex reachable (checked through the edge count of a closed triangle mesh: E = 3F/2)
    tri = got["indices"].reshape(-1, 3).astype(np.int64)
    edges = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    und = np.unique(np.sort(edges, axis=1), axis=0)
    assert len(und) * 2 == len(tri) * 3


def test_connected_components_flatten_is_race_free():
    """Regression test of the round-1 flake ("one run lost a land vertex after the clean-up", profiles/r02_flake.md): the pass that turns
    the union-find forest into representatives used path halving, whose stores could overwrite a vertex's final root with an inner node;
    the faces whose FIRST vertex that was were then dropped as "disconnected".  A long strip whose vertex ids are scrambled builds deep
    trees with many walkers per path: every one of its faces must survive meshRemoveDisconnected, in every repetition, also when
    three clean-ups run concurrently like the generator's three field tasks."""
    import threading

    n = 400000
    rng = np.random.default_rng(11)
    perm = rng.permutation(n).astype(np.uint32)  # vertex k of the strip gets id perm[k]
    k = np.arange(n - 2, dtype=np.uint32)
    tris = np.stack([perm[k], perm[k + 1 + (k & 1)], perm[k + 2 - (k & 1)]], axis=1)  # consistent winding along the strip
    tris = tris[rng.permutation(len(tris))]
    pos = np.zeros((n, 3), np.float32)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    pos[:, 0] = (inv // 2) * 0.004 - 800.0  # inside the +-2000 box, far apart compared with the 1e-5 merge threshold
    pos[:, 1] = (inv % 2) * 0.5
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (n, 1))
    idx = np.ascontiguousarray(tris.reshape(-1))
    failures = []

    def work(rep):
        m = mesh_clean(pos, nrm, idx, stages=capi.MESH_REMOVE_DISCONNECTED)
        if len(m["indices"]) != len(idx) or len(m["positions"]) != n:
            failures.append((rep, len(m["indices"]) // 3, len(m["positions"])))

    for rep in range(6):
        ts = [threading.Thread(target=work, args=(rep,)) for _ in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert not failures, failures[:3]
