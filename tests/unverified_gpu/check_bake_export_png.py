"""upg_bake_export_png = bake + texture tail + three PNG files in one call: decoding the files gives upg_bake_export's images.
(First run on hardware is the driver's: see tests/test_gpu_unverified.py.)"""
import io

import numpy as np
import pytest

from unnatural_planets_b200 import Planet, synth

pytestmark = pytest.mark.gpu


def pil_decode(data):
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(data)))


def test_bake_export_png_equals_bake_export_then_decode():
    p = Planet(shape="sphere", elevation="lakes", grid=160)
    for name, water in (("land", 0), ("water", 1)):
        m = p.field_meshes_clean((name,))[name]
        uv = synth.octahedral_uv(m["positions"])
        idx = synth.drop_folded_triangles(uv, m["indices"])
        want = p.bake_export(water, m["positions"], m["normals"], uv, idx, 1024, 768)
        covered = p.last_covered_texels
        files = p.bake_export_png(water, m["positions"], m["normals"], uv, idx, 1024, 768)
        assert p.last_covered_texels == covered
        for f, img in zip(files, want):
            got = pil_decode(f.tobytes())
            assert np.array_equal(got.reshape(img.shape), img), name
            assert len(f) < img.nbytes
