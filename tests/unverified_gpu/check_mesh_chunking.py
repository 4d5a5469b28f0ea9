"""upg_mesh_chunking (cage's meshChunking as the generator's meshSplit calls it) against the reference compiled from the
reference tree (oracle/_ref): the same chunks in the same order, positions / normals / indices bit for bit, empty parts included.
The decisions (axis, candidate planes, best split, enlarged boxes) are host functions already checked on the CPU against the reference
(tests/test_oracle_chunks.py); what runs here for the first time are the three small kernels (bounding box, areas + plane bits, sums in
triangle order) and K9's clip against boxes that are not cubes.  (First run on hardware is the driver's: tests/test_gpu_unverified.py.)"""
import numpy as np
import pytest

from conftest import golden
from unnatural_planets_b200 import Planet, mesh_chunking

pytestmark = pytest.mark.gpu


def ref_or_skip():
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    return oref.RefPlanet(init=False)


def same_chunks(got, want, what):
    assert len(got) == len(want), (what, len(got), len(want))
    for k, (a, b) in enumerate(zip(got, want)):
        for name in ("positions", "normals", "indices"):
            assert a[name].shape == b[name].shape, (what, k, name, a[name].shape, b[name].shape)
            assert np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), (what, k, name)


@pytest.mark.parametrize("max_area", [800000.0, 150000.0, 1e9])
def test_chunks_of_a_fixture_mesh(max_area):
    ref = ref_or_skip()
    g = golden("grid_sphere_lakes_40.npz")
    pos, nrm, idx = g["final_positions"], g["final_normals"], g["final_indices"]
    want = ref.chunking(pos, nrm, idx, max_surface_area=max_area)
    got = mesh_chunking(pos, nrm, idx, max_surface_area=max_area)
    assert len(want) == 1 if max_area == 1e9 else len(want) >= 16
    same_chunks(got, want, max_area)


@pytest.mark.parametrize("shape,elev", [("sphere", "lakes"), ("torus", "islands"), ("knot", "simple")])
def test_chunks_of_the_debug_planet(shape, elev):
    """the generator's debug configuration (110^3 grid, sources/meshGeneration.cpp:20-23): the GPU's own cleaned land mesh, split like meshSplit does"""
    ref = ref_or_skip()
    m = Planet(shape=shape, elevation=elev, grid=110).field_meshes_clean(("land",))["land"]
    want = ref.chunking(m["positions"], m["normals"], m["indices"], max_surface_area=800000.0)
    got = mesh_chunking(m["positions"], m["normals"], m["indices"], max_surface_area=800000.0)
    assert len(want) >= 8
    same_chunks(got, want, (shape, elev))


def test_empty_mesh_and_arguments():
    from unnatural_planets_b200 import UpgError

    assert mesh_chunking(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.uint32)) == []
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    with pytest.raises(UpgError):
        mesh_chunking(tri, tri, np.array([0, 1, 2], np.uint32), max_surface_area=0.0)
    one = mesh_chunking(tri, tri, np.array([0, 1, 2], np.uint32), max_surface_area=10.0)
    assert len(one) == 1 and np.array_equal(one[0]["positions"], tri) and list(one[0]["indices"]) == [0, 1, 2]


@pytest.mark.parametrize("name", ["flat", "strip", "tilted", "equal_extents"])
def test_chunks_of_degenerate_boxes(name):
    """the meshes of tests/test_oracle_chunks.py::test_chunking_decisions_on_degenerate_boxes through the device path"""
    from test_oracle_chunks import _grid_mesh

    ref = ref_or_skip()
    if name == "flat":
        pos, nrm, idx = _grid_mesh(40, 30, 2000.0, 1500.0, lambda x, y: 0 * x + 7.0)
    elif name == "strip":
        pos, nrm, idx = _grid_mesh(200, 3, 9000.0, 90.0, lambda x, y: np.sin(x * 0.01) * 20)
    elif name == "tilted":
        pos, nrm, idx = _grid_mesh(30, 30, 1000.0, 1400.0, lambda x, y: y * 1.0001)
    else:
        pos, nrm, idx = _grid_mesh(32, 32, 1600.0, 1600.0, lambda x, y: 0.25 * x)
    for max_area in (300000.0, 60000.0):
        same_chunks(mesh_chunking(pos, nrm, idx, max_surface_area=max_area), ref.chunking(pos, nrm, idx, max_surface_area=max_area), (name, max_area))
