"""Multi-GPU inside ONE process through the C ABI (upg_planet_desc::device_count > 1, csrc/upg_multi.cu): z-slab sharded sampling +
extraction, field-per-device clean-up, texel-row sharded baking -- everything bit-identical to one device.  Needs >= 2 GPUs
(`gpurun --gpus 2`); on a 1-GPU box the tests skip, and the multi-device planet of ONE device is covered instead.

These tests pin the PEER exchange (NVLink copies between peer-mapped pools): that is the variant the 8-GPU run of round 2 passed at these
sizes on 2 / 3 / 4 / 8 devices.  The library's default beyond two devices is the host-staged exchange written after that run
(csrc/upg_multi.cu); it has not run on hardware and is checked by tests/unverified_gpu/check_multi_host_exchange.py,
which runs these same tests with UPG_MULTI_EXCHANGE=host in its own process."""
import os

os.environ.setdefault("UPG_MULTI_EXCHANGE", "peer")  # read once by the library, before the first multi-device planet

import numpy as np  # noqa: E402
import pytest  # noqa: E402

from unnatural_planets_b200 import Planet, capi, synth  # noqa: E402

pytestmark = pytest.mark.gpu


def device_count():
    import torch

    return torch.cuda.device_count()


def devices_or_skip(n):
    if device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    return list(range(n))


def same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert (a.view(np.uint32) == b.view(np.uint32)).all(), what


def test_single_entry_device_list_is_the_single_device_planet():
    a = Planet(shape="sphere", elevation="lakes", grid=64)
    b = Planet(shape="sphere", elevation="lakes", grid=64, devices=[0])
    ma, mb = a.field_meshes_clean(), b.field_meshes_clean()
    for f in ma:
        for k in ("positions", "normals", "indices"):
            same(ma[f][k], mb[f][k], (f, k))


@pytest.mark.parametrize("n", [2, 3, 4, 8])
@pytest.mark.parametrize("shape,elev,res", [("sphere", "lakes", 160), ("torus", "craters", 130), ("mandelbulb", "simple", 96)])
def test_multi_device_meshes_equal_one_device(n, shape, elev, res):
    devs = devices_or_skip(n)
    one = Planet(shape=shape, elevation=elev, grid=res)
    many = Planet(shape=shape, elevation=elev, grid=res, devices=devs)
    a, b = one.field_meshes(("land", "water", "navigation")), many.field_meshes(("land", "water", "navigation"))
    for f in a:
        for k in ("positions", "normals", "indices", "dmc_vertices", "dmc_quads"):
            same(a[f][k], b[f][k], (n, f, k))
        # the statistic counts the cell origins of every slab INCLUDING its recomputed halo slices: never less than one device's
        assert a[f]["near_zero_samples"] <= b[f]["near_zero_samples"] <= a[f]["near_zero_samples"] * (1 + n) + 8
    a, b = one.field_meshes_clean(), many.field_meshes_clean()
    for f in a:
        for k in ("positions", "normals", "indices"):
            same(a[f][k], b[f][k], (n, "clean", f, k))
    # a subset of the fields, and the tail stages alone
    a, b = one.field_meshes_clean(("navigation",), stages=capi.MESH_MAKEMESH_TAIL), many.field_meshes_clean(("navigation",), stages=capi.MESH_MAKEMESH_TAIL)
    same(a["navigation"]["indices"], b["navigation"]["indices"], "tail only")


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_device_bake_equals_one_device(n):
    devs = devices_or_skip(n)
    one = Planet(shape="sphere", elevation="lakes", grid=200)
    many = Planet(shape="sphere", elevation="lakes", grid=200, devices=devs)
    for name, water in (("land", 0), ("water", 1)):
        m = one.field_meshes_clean((name,))[name]
        uv = synth.octahedral_uv(m["positions"])
        idx = synth.drop_folded_triangles(uv, m["indices"])
        for size in (1024, 777):
            a = one.bake_export(water, m["positions"], m["normals"], uv, idx, size, size)
            ca = one.last_covered_texels
            b = many.bake_export(water, m["positions"], m["normals"], uv, idx, size, size)
            assert many.last_covered_texels == ca
            for x, y, what in zip(a, b, ("albedo", "pbr", "normal")):
                assert np.array_equal(x, y), (n, name, size, what)
        a = one.bake(water, m["positions"], m["normals"], uv, idx, 640, 512)
        b = many.bake(water, m["positions"], m["normals"], uv, idx, 640, 512)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_multi_device_errors():
    devs = devices_or_skip(2)
    from unnatural_planets_b200 import UpgError

    with pytest.raises(UpgError):  # devices[0] must be the home device; no duplicates
        Planet(shape="sphere", elevation="lakes", grid=64, devices=[0, 0])
    p = Planet(shape="sphere", elevation="none", grid=16, box=100.0, devices=devs)
    with pytest.raises(UpgError) as e:  # the reference throws "generated empty base land mesh"
        p.field_meshes_clean(("land",))
    assert e.value.code == 4
    assert len(p.field_meshes_clean(("water",))["water"]["indices"]) == 0
