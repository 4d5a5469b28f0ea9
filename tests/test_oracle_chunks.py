"""oracle/_ref also carries the serial geometry between stage 2 and stage 3 (cage meshChunking + xatlas meshUnwrap, compiled from the
reference tree) so that the GPU bake can be checked on real chunk atlases (tests/test_gpu_chunks.py).  CPU-side sanity of that harness."""
import numpy as np
import pytest

from conftest import golden


def test_chunk_and_unwrap_of_a_reference_mesh():
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    g = golden("grid_sphere_lakes_40.npz")
    ref = oref.RefPlanet(init=False)
    pos, nrm, idx = g["final_positions"], g["final_normals"], g["final_indices"]
    chunks = ref.chunk_and_unwrap(pos, nrm, idx, max_surface_area=800000.0, texels_per_unit=0.3, padding=6)
    assert len(chunks) >= 16  # a r = 1100 sphere has 1.5e7 units^2 of surface: at least 19 chunks of <= 8e5

    def area(c):
        p, i = c["positions"].astype(np.float64), c["indices"].reshape(-1, 3)
        return 0.5 * np.linalg.norm(np.cross(p[i[:, 1]] - p[i[:, 0]], p[i[:, 2]] - p[i[:, 0]]), axis=1).sum()

    total = sum(area(c) for c in chunks)
    whole = area(dict(positions=pos, indices=idx))
    assert abs(total - whole) / whole < 1e-3  # clipping at the split planes keeps the surface
    for c in chunks:
        assert area(c) < 800000.0 * 1.001
        assert c["uvs"].min() >= 0.0 and c["uvs"].max() <= 1.0 and c["resolution"] >= 32
        assert len(c["uvs"]) == len(c["positions"]) == len(c["normals"]) and c["indices"].max() < len(c["positions"])


# ---- the product's meshChunking (upg_mesh_chunking) on the CPU: its DECISIONS are plain host functions (csrc/upg_chunking_host.h, compiled
# here into tests/hostlib/libupg_hostchunking.so) fed by a bounding box and eight serial sums.  This emulation feeds them with numpy
# restatements of what the three small kernels deliver (bounding box, per-triangle area + below-plane bits, sums in triangle order) and
# uses the REFERENCE's meshClip for the clip (the device clip is K9's, checked bit for bit elsewhere): the recursion, the order of the
# leaves, the empty parts, the enlarged boxes and every float expression of the decisions must then reproduce cage's meshChunking exactly.
def _hostchunking():
    import ctypes as C
    import os
    import subprocess

    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostlib")
    subprocess.run(["make", "-s", "-C", d, "libupg_hostchunking.so"], check=True)
    L = C.CDLL(os.path.join(d, "libupg_hostchunking.so"))
    fp = C.POINTER(C.c_float)
    L.upgh_chunk_longest_axis.argtypes = [fp, fp]
    L.upgh_chunk_candidates.argtypes = [fp, fp, C.c_int, fp]
    L.upgh_chunk_wants_split.argtypes = [fp, C.c_float, fp]
    L.upgh_chunk_clipping_box.argtypes = [fp, fp, C.c_int, C.c_float, C.c_int, fp]
    return L


def _emulated_chunking(L, ref, pos, nrm, idx, max_area, depth=0):
    import ctypes as C

    f32 = np.float32
    fp = C.POINTER(C.c_float)
    if len(idx) == 0:
        return [dict(positions=pos, normals=nrm, indices=idx)]
    assert depth < 48
    # chunk_bbox_kernel: minimum / maximum over all positions
    a, b = pos.min(axis=0).astype(f32), pos.max(axis=0).astype(f32)
    axis = L.upgh_chunk_longest_axis(a.ctypes.data_as(fp), b.ctypes.data_as(fp))
    cand = np.zeros(7, f32)
    L.upgh_chunk_candidates(a.ctypes.data_as(fp), b.ctypes.data_as(fp), axis, cand.ctypes.data_as(fp))
    # chunk_areas_kernel: Triangle::area() and center()[axis] in single precision, cage's operation order
    t = idx.reshape(-1, 3)
    va, vb, vc = pos[t[:, 0]], pos[t[:, 1]], pos[t[:, 2]]
    l, r = vb - va, vc - va
    cr = np.stack([l[:, 1] * r[:, 2] - l[:, 2] * r[:, 1], l[:, 2] * r[:, 0] - l[:, 0] * r[:, 2], l[:, 0] * r[:, 1] - l[:, 1] * r[:, 0]], axis=1).astype(f32)
    area = (np.sqrt(((f32(0) + cr[:, 0] * cr[:, 0]) + cr[:, 1] * cr[:, 1]) + cr[:, 2] * cr[:, 2]).astype(f32) * f32(0.5)).astype(f32)
    centre = (((va + vb) + vc) / f32(3)).astype(f32)[:, axis]
    # chunk_serial_sums_kernel: running single-precision sums in triangle order (np.cumsum accumulates sequentially)
    sums = np.zeros(8, f32)
    sums[0] = np.cumsum(area, dtype=f32)[-1]
    for i in range(7):
        part = np.where(centre < cand[i], area, f32(0)).astype(f32)  # adding +0 leaves a running sum unchanged
        sums[1 + i] = np.cumsum(part, dtype=f32)[-1]
    fraction = C.c_float(0)
    if not L.upgh_chunk_wants_split(sums.ctypes.data_as(fp), f32(max_area), C.byref(fraction)):
        return [dict(positions=pos, normals=nrm, indices=idx)]
    out = []
    for second in (0, 1):
        box = np.zeros(6, f32)
        L.upgh_chunk_clipping_box(a.ctypes.data_as(fp), b.ctypes.data_as(fp), axis, fraction, second, box.ctypes.data_as(fp))
        c = ref.clip_box(pos, nrm, idx, box)
        out += _emulated_chunking(L, ref, c["positions"], c["normals"], c["indices"], max_area, depth + 1)
    return out


@pytest.mark.parametrize("fixture,max_area", [("grid_sphere_lakes_40.npz", 800000.0), ("grid_sphere_lakes_40.npz", 150000.0)])
def test_chunking_decisions_reproduce_the_reference(fixture, max_area):
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    g = golden(fixture)
    ref = oref.RefPlanet(init=False)
    pos, nrm, idx = (np.ascontiguousarray(g[k]) for k in ("final_positions", "final_normals", "final_indices"))
    want = ref.chunking(pos, nrm, idx, max_surface_area=max_area)
    got = _emulated_chunking(_hostchunking(), ref, pos.astype(np.float32).reshape(-1, 3), nrm.astype(np.float32).reshape(-1, 3), idx.astype(np.uint32).reshape(-1), max_area)
    assert len(got) == len(want) and len(want) >= 16
    for k, (a, b) in enumerate(zip(got, want)):
        for name in ("positions", "normals", "indices"):
            assert a[name].shape == b[name].shape and np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), (k, name)


def _grid_mesh(nx, ny, sx, sy, height):
    """a (nx x ny)-quad height field over [0, sx] x [0, sy]: positions, unit normals (0, 0, 1), triangle indices"""
    xs, ys = np.meshgrid(np.linspace(0, sx, nx + 1, dtype=np.float32), np.linspace(0, sy, ny + 1, dtype=np.float32))
    pos = np.stack([xs.ravel(), ys.ravel(), height(xs, ys).astype(np.float32).ravel()], axis=1).astype(np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (len(pos), 1))
    i = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).ravel()
    idx = np.stack([i, i + 1, i + nx + 2, i, i + nx + 2, i + nx + 1], axis=1).astype(np.uint32).ravel()
    return pos, nrm, idx


@pytest.mark.parametrize("name", ["flat", "strip", "tilted", "equal_extents"])
def test_chunking_decisions_on_degenerate_boxes(name):
    """bounding boxes with a zero extent, two equal extents (dominantAxis' strict comparisons), a long strip"""
    from oracle import ref as oref

    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    ref = oref.RefPlanet(init=False)
    if name == "flat":  # zero extent in z
        pos, nrm, idx = _grid_mesh(40, 30, 2000.0, 1500.0, lambda x, y: 0 * x + 7.0)
    elif name == "strip":  # x dominates
        pos, nrm, idx = _grid_mesh(200, 3, 9000.0, 90.0, lambda x, y: np.sin(x * 0.01) * 20)
    elif name == "tilted":  # z extent close to y extent
        pos, nrm, idx = _grid_mesh(30, 30, 1000.0, 1400.0, lambda x, y: y * 1.0001)
    else:  # x and y extents equal: neither `>` holds first, the third test decides
        pos, nrm, idx = _grid_mesh(32, 32, 1600.0, 1600.0, lambda x, y: 0.25 * x)
    for max_area in (300000.0, 60000.0):
        want = ref.chunking(pos, nrm, idx, max_surface_area=max_area)
        got = _emulated_chunking(_hostchunking(), ref, pos, nrm, idx, max_area)
        assert len(got) == len(want) and len(want) >= 4, (name, len(got), len(want))
        for k, (a, b) in enumerate(zip(got, want)):
            for key in ("positions", "normals", "indices"):
                assert a[key].shape == b[key].shape and np.array_equal(a[key].view(np.uint32), b[key].view(np.uint32)), (name, max_area, k, key)
