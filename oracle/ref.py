"""ctypes access to oracle/_ref/libup_ref.so -- TEST INFRASTRUCTURE.

The library is the reference's own translation units compiled verbatim (oracle/Makefile `ref`).
Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package never does.

One `RefPlanet` == one private copy of the shared object, because the reference keeps its noise
objects in function-local statics that are initialised once per process image.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libup_ref.so")

DEFAULT_SEED = (1346519564496, 42245614964156)  # the pair cage's own RNG test uses (test-core/random.cpp:38)


def available() -> bool:
    return os.path.exists(LIB_PATH)


class NoiseCfg(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("fractal", C.c_int32),
        ("distance", C.c_int32),
        ("operation", C.c_int32),
        ("seed", C.c_int32),
        ("octaves", C.c_int32),
        ("lacunarity", C.c_float),
        ("gain", C.c_float),
        ("frequency", C.c_float),
        ("bounding", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _Buf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("count", C.c_uint64)]


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class RefPlanet:
    """The reference generator configured like `unnatural-planets -s SHAPE -e ELEV -c COLORS` with an explicit RNG state."""

    FIELDS = {"shape": 0, "water": 0, "elevraw": 1, "land": 2, "navigation": 3, "elevation": 4}

    def __init__(self, shape="sphere", elevation="lakes", coloring="default", poles=False, flowers=False, seed=DEFAULT_SEED, init=True):
        if not available():
            raise RuntimeError("oracle/_ref/libup_ref.so is not built (needs /root/reference; run `make -C oracle ref`)")
        self._tmpdir = tempfile.mkdtemp(prefix="upref_")
        private = os.path.join(self._tmpdir, "libup_ref_%d.so" % id(self))
        shutil.copy(LIB_PATH, private)
        self.lib = C.CDLL(private, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        os.unlink(private)
        L = self.lib
        L.upref_last_error.restype = C.c_char_p
        L.upref_init.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        self.config = dict(shape=shape, elevation=elevation, coloring=coloring, poles=bool(poles), flowers=bool(flowers), seed=tuple(seed))
        if init:
            self._check(L.upref_init(shape.encode(), elevation.encode(), coloring.encode(), int(poles), int(flowers), seed[0], seed[1]))

    def __del__(self):
        try:
            shutil.rmtree(self._tmpdir, ignore_errors=True)
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.upref_last_error().decode())

    # --- RNG -----------------------------------------------------------------------------------
    def rng_state(self):
        s = (C.c_uint64 * 2)()
        self.lib.upref_rng_state(s)
        return int(s[0]), int(s[1])

    def rng_raw(self, s0, s1, n):
        out = np.zeros(n, dtype=np.uint64)
        self.lib.upref_rng_kat(C.c_uint64(s0), C.c_uint64(s1), C.c_int(n), _fp(out))
        return out

    def rng_range3(self, s0, s1, a, b, ia, ib):
        f = np.zeros(3, dtype=np.float32)
        i = np.zeros(3, dtype=np.int32)
        self.lib.upref_rng_range3(C.c_uint64(s0), C.c_uint64(s1), C.c_float(a), C.c_float(b), _fp(f), _fp(i), C.c_int(ia), C.c_int(ib))
        return f, i

    # --- noise registry ------------------------------------------------------------------------
    def noise_configs(self):
        n = self.lib.upref_noise_count()
        out = []
        for i in range(n):
            c = NoiseCfg()
            self.lib.upref_noise_get(C.c_int(i), C.byref(c))
            out.append(c.as_dict())
        return out

    def noise_eval(self, index, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(pos), dtype=np.float32)
        self.lib.upref_noise_eval(C.c_int(index), _fp(pos), C.c_int(len(pos)), _fp(out))
        return out

    # --- fields --------------------------------------------------------------------------------
    def eval(self, what, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(pos), dtype=np.float32)
        self._check(self.lib.upref_eval(C.c_int(self.FIELDS[what]), _fp(pos), C.c_int(len(pos)), _fp(out)))
        return out

    def grid_positions(self, res, box=4000.0):
        out = np.zeros(res, dtype=np.float32)
        self.lib.upref_grid_positions(C.c_float(box), C.c_int(res), _fp(out))
        return out

    def grid(self, what, res, box=4000.0, z0=0, z1=None, threads=1):
        z1 = res if z1 is None else z1
        out = np.zeros((z1 - z0, res, res), dtype=np.float32)
        self._check(self.lib.upref_grid(C.c_int(self.FIELDS[what]), C.c_float(box), C.c_int(res), C.c_int(z0), C.c_int(z1), _fp(out), C.c_int(threads)))
        return out

    # --- isosurface ----------------------------------------------------------------------------
    def _take(self, buf, dtype, width):
        n = int(buf.count)
        if n == 0 or not buf.ptr:
            arr = np.zeros((0, width) if width > 1 else (0,), dtype=dtype)
        else:
            raw = (C.c_char * (n * width * np.dtype(dtype).itemsize)).from_address(buf.ptr)
            arr = np.frombuffer(raw, dtype=dtype).copy()
            if width > 1:
                arr = arr.reshape(n, width)
        if buf.ptr:
            self.lib.upref_free(C.c_void_p(buf.ptr))
        return arr

    def dualmc(self, dens, iso=0.0, manifold=True):
        dens = np.ascontiguousarray(dens, dtype=np.float32)
        rz, ry, rx = dens.shape
        v, q = _Buf(), _Buf()
        self._check(self.lib.upref_dualmc(_fp(dens), C.c_int(rx), C.c_int(ry), C.c_int(rz), C.c_float(iso), C.c_int(int(manifold)), C.byref(v), C.byref(q)))
        return self._take(v, np.float32, 3), self._take(q, np.int32, 4)

    def make_mesh(self, dens, box=4000.0, clip=True, remove_disconnected=False, flip=False):
        dens = np.ascontiguousarray(dens, dtype=np.float32)
        res = dens.shape[0]
        assert dens.shape == (res, res, res)
        p, n, i = _Buf(), _Buf(), _Buf()
        stages = 1 | (2 if remove_disconnected else 0) | (4 if flip else 0)
        self._check(self.lib.upref_make_mesh(_fp(dens), C.c_float(box), C.c_int(res), C.c_int(int(clip)), C.c_int(stages), C.byref(p), C.byref(n), C.byref(i)))
        return self._take(p, np.float32, 3), self._take(n, np.float32, 3), self._take(i, np.uint32, 1)

    # --- tiles / textures ----------------------------------------------------------------------
    TILE_COLUMNS = ("albedo_r", "albedo_g", "albedo_b", "roughness", "metallic", "height", "elevation", "slope", "temperature", "precipitation", "opacity", "biome", "type")

    def tiles(self, purpose, pos, nrm):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(pos), 13), dtype=np.float32)
        code = {"land": 1, "water": 2, "navigation": 3}[purpose]
        self._check(self.lib.upref_tiles(C.c_int(code), _fp(pos), _fp(nrm), C.c_int(len(pos)), _fp(out)))
        return out

    def tiles_seeded(self, purpose, pos, nrm):
        """as tiles(), with the thread RNG re-seeded from each tile position (pins generateSand's colorDeviation)"""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(pos), 13), dtype=np.float32)
        code = {"land": 1, "water": 2, "navigation": 3}[purpose]
        self._check(self.lib.upref_tiles_seeded(C.c_int(code), _fp(pos), _fp(nrm), C.c_int(len(pos)), _fp(out)))
        return out

    def bake_seeded(self, water, pos, nrm, uv, idx, w, h):
        """generateTextures* re-assembled from the reference's building blocks with the position-keyed RNG (ref_harness.cpp)"""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        ch = 4 if water else 3
        albedo = np.zeros((h, w, ch), dtype=np.uint8)
        special = np.zeros((h, w, 2), dtype=np.uint8)
        height = np.zeros((h, w, 1), dtype=np.uint8)
        covered = C.c_uint64(0)
        self._check(self.lib.upref_bake_seeded(C.c_int(int(water)), _fp(pos), _fp(nrm), _fp(uv), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.c_uint32(w), C.c_uint32(h), _fp(albedo), _fp(special), _fp(height), C.byref(covered)))
        self.last_covered_texels = int(covered.value)
        return albedo, special, height

    def bake(self, water, pos, nrm, uv, idx, w, h):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        ch = 4 if water else 3
        albedo = np.zeros((h, w, ch), dtype=np.uint8)
        special = np.zeros((h, w, 2), dtype=np.uint8)
        height = np.zeros((h, w, 1), dtype=np.uint8)
        self._check(self.lib.upref_bake(C.c_int(int(water)), _fp(pos), _fp(nrm), _fp(uv), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.c_uint32(w), C.c_uint32(h), _fp(albedo), _fp(special), _fp(height)))
        return albedo, special, height

    def raster(self, pos, uv, idx, w, h):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        owner = np.zeros((h, w), dtype=np.uint32)
        weights = np.zeros((h, w, 3), dtype=np.float32)
        calls = C.c_uint64(0)
        self._check(self.lib.upref_raster(_fp(pos), _fp(uv), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.c_uint32(w), C.c_uint32(h), _fp(owner), _fp(weights), C.byref(calls)))
        return owner, weights, int(calls.value)

    def color(self, what, v):
        code = {"rgb2hsv": 0, "hsv2rgb": 1, "rgb2hsluv": 2, "hsluv2rgb": 3}[what]
        i = np.ascontiguousarray(v, dtype=np.float32)
        o = np.zeros(3, dtype=np.float32)
        self.lib.upref_color(C.c_int(code), _fp(i), _fp(o))
        return o

    def color_interpolate(self, a, b, f):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        o = np.zeros(3, dtype=np.float32)
        self.lib.upref_color_interpolate(_fp(a), _fp(b), C.c_float(f), _fp(o))
        return o

    def dilate(self, img, rounds, use_nan=True, to_u8=False):
        img = np.array(img, dtype=np.float32, order="C", copy=True)
        h, w, ch = img.shape
        u8 = np.zeros((h, w, ch), dtype=np.uint8) if to_u8 else None
        self._check(self.lib.upref_dilate(_fp(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(ch), C.c_uint32(rounds), C.c_int(int(use_nan)), _fp(u8) if to_u8 else None))
        return (img, u8) if to_u8 else img

    def texture_tail(self, special, height, strength=1.0):
        """imageConvertSpecialToGltfPbr + imageConvertHeigthToNormal on U8 images (sources/generator.cpp:135-139)."""
        special = np.ascontiguousarray(special, dtype=np.uint8)
        height = np.ascontiguousarray(height, dtype=np.uint8)
        h, w = height.shape[:2]
        pbr = np.zeros((h, w, 3), dtype=np.uint8)
        normal = np.zeros((h, w, 3), dtype=np.uint8)
        self._check(self.lib.upref_texture_tail(_fp(special), _fp(height), C.c_uint32(w), C.c_uint32(h), C.c_float(strength), _fp(pbr), _fp(normal)))
        return pbr, normal

    def water_cull(self, positions, normals, indices):
        """the vertex culling of meshGenerateBaseWater (sources/meshGeneration.cpp:62-95) over a given mesh."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        p, n, i = _Buf(), _Buf(), _Buf()
        self._check(self.lib.upref_water_cull(_fp(pos), _fp(nrm), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.byref(p), C.byref(n), C.byref(i)))
        return self._take(p, np.float32, 3), self._take(n, np.float32, 3), self._take(i, np.uint32, 1)

    # --- chunking + unwrap (the serial geometry between stage 2 and stage 3; out of the product's scope) ---------------------------
    def chunking(self, positions, normals, indices, max_surface_area=800000.0):
        """meshSplit of sources/meshGeneration.cpp:173-182 (cage meshChunking) alone: list of dict(positions, normals, indices)"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        n = C.c_uint32(0)
        self._check(self.lib.upref_chunking(_fp(pos), _fp(nrm), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.c_float(max_surface_area), C.byref(n)))
        return [self._held_chunk(i) for i in range(n.value)]

    def _held_chunk(self, i):
        nv, ni = C.c_uint32(0), C.c_uint32(0)
        self.lib.upref_chunk_size(C.c_uint32(i), C.byref(nv), C.byref(ni))
        c = dict(positions=np.zeros((nv.value, 3), np.float32), normals=np.zeros((nv.value, 3), np.float32), indices=np.zeros(ni.value, np.uint32))
        if nv.value or ni.value:
            self.lib.upref_chunk_copy(C.c_uint32(i), _fp(c["positions"]), _fp(c["normals"]), None, _fp(c["indices"]))
        return c

    def clip_box(self, positions, normals, indices, box6):
        """cage meshClip(mesh, Aabb) (meshAlgorithms.cpp:1056-1114) of a copy of the mesh: dict(positions, normals, indices)"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        box = np.ascontiguousarray(box6, dtype=np.float32).reshape(6)
        nv, ni = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.upref_mesh_clip_box(_fp(pos), _fp(nrm), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), _fp(box), C.byref(nv), C.byref(ni)))
        return self._held_chunk(0)

    def chunk_and_unwrap(self, positions, normals, indices, max_surface_area=800000.0, texels_per_unit=1.35, padding=6):
        """meshSplit + meshUnwrap of sources/meshGeneration.cpp:173-197 (cage meshChunking, xatlas): a list of chunks, each a dict with
        positions, normals, uvs, indices and the atlas `resolution` xatlas chose for it"""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        n = C.c_uint32(0)
        self._check(self.lib.upref_chunking(_fp(pos), _fp(nrm), C.c_uint32(len(pos)), _fp(idx), C.c_uint32(len(idx)), C.c_float(max_surface_area), C.byref(n)))
        out = []
        for i in range(n.value):
            res, nv, ni = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
            self._check(self.lib.upref_chunk_unwrap(C.c_uint32(i), C.c_float(texels_per_unit), C.c_uint32(padding), C.byref(res), C.byref(nv), C.byref(ni)))
            c = dict(positions=np.zeros((nv.value, 3), np.float32), normals=np.zeros((nv.value, 3), np.float32), uvs=np.zeros((nv.value, 2), np.float32), indices=np.zeros(ni.value, np.uint32), resolution=int(res.value))
            self.lib.upref_chunk_copy(C.c_uint32(i), _fp(c["positions"]), _fp(c["normals"]), _fp(c["uvs"]), _fp(c["indices"]))
            out.append(c)
        return out

