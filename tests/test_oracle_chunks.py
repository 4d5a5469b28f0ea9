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
