"""The integration patch COMPILED (tests/shim/gpu.cpp: the translation unit INTEGRATION.md tells a maintainer to add) against the
reference's own headers, linked with the reference's objects (oracle/_ref) and libupg.so, and called through the reference's own types:
meshGenerateBaseLand / Water / Navigation() -> Holder<Mesh>, generateTextures{Land,Water}(const Holder<Mesh> &, w, h, Holder<Image> & x 3).
Built by tests/shim/Makefile where /root/reference exists; the library travels to the GPU box with the snapshot."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "shim", "libupg_shim.so")
SEED = "1346519564496,42245614964156"  # the --seed option of the patch: the two words of cage's own RNG test (test-core/random.cpp:38)


def shim_or_skip():
    if not os.path.exists(LIB):
        if os.path.isdir("/root/reference/sources"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "shim")])
        else:
            pytest.skip("tests/shim/libupg_shim.so not built (needs the reference checkout)")
    L = C.CDLL(LIB)
    L.upgshim_last_error.restype = C.c_char_p
    L.upgshim_setup.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.upgshim_base_mesh.argtypes = [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.upgshim_base_mesh_copy.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.upgshim_textures.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def test_shim_links_against_the_reference_and_fails_loudly_without_a_gpu():
    """compiles + links (every cage / unnatural symbol the patch uses resolves against the reference's own objects); without a CUDA
    device planet creation fails and the patch turns the C error into cage's exception -- there is no CPU fallback behind the shim"""
    L = shim_or_skip()
    for name in ("upgshim_setup", "upgshim_base_mesh", "upgshim_base_mesh_copy", "upgshim_textures", "upgshim_rng_state"):
        assert hasattr(L, name)
    try:
        import torch

        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    if gpu:
        pytest.skip("covered by the GPU test below")
    # own process: the reference keeps its noise objects in function-local statics (one planet per process image)
    code = "import ctypes as C, sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests'); from test_shim import shim_or_skip; L = shim_or_skip(); rc = L.upgshim_setup(b'sphere', b'lakes', b'default', 0, 0, %r.encode(), 110); print(rc, L.upgshim_last_error().decode())" % (ROOT, ROOT, SEED)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    rc, msg = r.stdout.strip().split(" ", 1)
    assert rc == "1" and "gpu generation core failed" in msg, r.stdout


RUNNER = r"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + '/tests')
from test_shim import shim_or_skip, SEED
from unnatural_planets_b200 import Planet, synth
L = shim_or_skip()
shape, elev, grid, tex = %(shape)r, %(elev)r, %(grid)d, %(tex)d
assert L.upgshim_setup(shape.encode(), elev.encode(), b"default", 0, 0, SEED.encode(), grid) == 0, L.upgshim_last_error()
p = Planet(shape=shape, elevation=elev, grid=grid)     # the library on its own, same seed
want = p.field_meshes_clean(("land", "water", "navigation"))
out = {"meshes": {}}
shim_land = None
for f, name in enumerate(("land", "water", "navigation")):
    nv, ni = C.c_uint32(0), C.c_uint32(0)
    assert L.upgshim_base_mesh(f, C.byref(nv), C.byref(ni)) == 0, L.upgshim_last_error()
    pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nv.value, 3), np.float32); idx = np.zeros(ni.value, np.uint32)
    L.upgshim_base_mesh_copy(f, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
    w = want[name]
    out["meshes"][name] = {"vertices": int(nv.value), "triangles": int(ni.value // 3),
        "equal": bool(pos.shape == w["positions"].shape and (pos.view(np.uint32) == w["positions"].view(np.uint32)).all() and (nrm.view(np.uint32) == w["normals"].view(np.uint32)).all() and np.array_equal(idx, w["indices"]))}
    if name == "land":
        shim_land = (pos, nrm, idx)
pos, nrm, idx = shim_land
uv = synth.octahedral_uv(pos)
idx = synth.drop_folded_triangles(uv, idx)
res = {}
for which in (0, 1):   # 0: the patch (GPU), 1: the reference's own generateTexturesLand (CPU), both through Holder<Mesh> / Holder<Image>
    a = np.zeros((tex, tex, 3), np.uint8); s = np.zeros((tex, tex, 2), np.uint8); h = np.zeros((tex, tex, 1), np.uint8)
    rc = L.upgshim_textures(which, 0, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), tex, tex, a.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p))
    assert rc == 0, L.upgshim_last_error()
    res[which] = (a, s, h)
# the reference's own body a second time: sand texels draw their colour deviation from the thread RNG (SURVEY.md F7), so only the
# texels it reproduces itself can be compared
a2 = np.zeros((tex, tex, 3), np.uint8); s2 = np.zeros((tex, tex, 2), np.uint8); h2 = np.zeros((tex, tex, 1), np.uint8)
assert L.upgshim_textures(1, 0, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), tex, tex, a2.ctypes.data_as(C.c_void_p), s2.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p)) == 0
stable = (a2 == res[1][0]).all(axis=2)
direct = p.bake(0, pos, nrm, uv, idx, tex, tex)
out["textures"] = {"equal_to_direct_call": bool(all(np.array_equal(x, y) for x, y in zip(res[0], direct)))}
for k, name in enumerate(("albedo", "special", "height")):
    d = np.abs(res[0][k].astype(np.int32) - res[1][k].astype(np.int32))
    out["textures"][name] = {"within_1": float((d <= 1).mean()), "max": int(d.max()), "covered": float((res[1][k] != 0).any(axis=2).mean())}
    if name == "albedo":
        out["textures"][name]["stable_fraction"] = float(stable.mean())
        out["textures"][name]["within_1_on_stable"] = float((d.max(axis=2) <= 1)[stable].mean())
print(json.dumps(out))
"""


@pytest.mark.gpu
@pytest.mark.parametrize("shape,elev", [("sphere", "lakes"), ("torus", "craters")])
def test_patched_entry_points_through_cage_types(shape, elev):
    shim_or_skip()
    code = RUNNER % dict(root=ROOT, shape=shape, elev=elev, grid=110, tex=256)  # 110^3 = the generator's debug grid (sources/meshGeneration.cpp:20-23)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for name, m in out["meshes"].items():
        # Holder<Mesh> of the patched meshGenerateBase* == the library's own result, bit for bit (which tests/test_gpu_fullsize.py and
        # tests/test_gpu_meshfix.py compare with the reference's makeMesh + meshRemoveDisconnected + meshFlipNormals)
        assert m["equal"], (name, m)
        assert m["triangles"] > 1000 or name == "water", (name, m)  # the water culling may leave nothing (torus + craters: every vertex is above 0.1)
    t = out["textures"]
    assert t["equal_to_direct_call"]
    # against the reference's own generateTexturesLand in the same process: <= 1/255 per channel except the sand texels (the reference
    # draws their colour deviation from the thread RNG) and the libm threshold flips tests/test_gpu_bake.py enumerates
    assert t["height"]["within_1"] > 0.995 and t["special"]["within_1"] > 0.99, t
    assert t["albedo"]["stable_fraction"] > 0.3 and t["albedo"]["within_1_on_stable"] > 0.99, t
    assert t["albedo"]["covered"] > 0.5
