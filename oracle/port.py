"""ctypes access to oracle/libup_oracle.so -- TEST INFRASTRUCTURE: the standalone CPU restatement of the hot path
(oracle/up_oracle*.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libup_oracle.so")

SHAPES = ("asteroid", "belt", "bowl", "box", "bunny", "capsule", "cube", "disk", "doubletorus", "fibers", "gear", "h2o", "h3o", "h4o", "hemispheres", "hexagon", "hexagonalprism", "insidecube", "knot", "mandelbulb", "mobiusstrip", "monkeyhead", "octahedron", "pipe", "sphere", "square", "tetrahedron", "torus", "toruscross", "triangularprism", "tube", "twistedhexagonalprism", "twistedplane", "wormhole")
ELEVATIONS = ("none", "simple", "legacy", "lakes", "islands", "craters")
FIELDS = {"land": 0, "water": 1, "navigation": 2, "elevraw": 3, "elevation_raw": 3, "elevation": 4}
DEFAULT_SEED = (1346519564496, 42245614964156)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        L = C.CDLL(LIB_PATH)
        L.upo_planet_create.restype = C.c_void_p
        L.upo_planet_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        L.upo_planet_destroy.argtypes = [C.c_void_p]
        L.upo_seed_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.upo_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.upo_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.upo_grid_coords.argtypes = [C.c_float, C.c_int, C.c_void_p]
        L.upo_cellular_table.restype = C.POINTER(C.c_float)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OraclePlanet:
    def __init__(self, shape="sphere", elevation="lakes", seed=DEFAULT_SEED):
        self.L = lib()
        self.h = self.L.upo_planet_create(SHAPES.index(shape), ELEVATIONS.index(elevation), seed[0], seed[1])
        if not self.h:
            raise ValueError("unknown mode")

    def __del__(self):
        try:
            self.L.upo_planet_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    def seed_log(self):
        out = np.zeros(128, dtype=np.uint32)
        n = self.L.upo_seed_log(C.c_void_p(self.h), _p(out), 128)
        return out[:n].copy()

    def points(self, field, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(pos), dtype=np.float32)
        self.L.upo_points(C.c_void_p(self.h), FIELDS[field], _p(pos), len(pos), _p(out))
        return out

    def grid(self, field, res, box=4000.0, z0=0, z1=None, threads=1):
        z1 = res if z1 is None else z1
        out = np.zeros((z1 - z0, res, res), dtype=np.float32)
        self.L.upo_grid(C.c_void_p(self.h), FIELDS[field], box, res, z0, z1, _p(out), threads)
        return out


def grid_coords(res, box=4000.0):
    out = np.zeros(res, dtype=np.float32)
    lib().upo_grid_coords(box, res, _p(out))
    return out


def sample_rate(shape, elevation, res, slices, threads):
    """(samples/s, seconds) of the restated sampling loop on `slices` equatorial z-slices of the land field"""
    p = OraclePlanet(shape, elevation)
    z0 = res // 2 - slices // 2
    t = time.perf_counter()
    p.grid("land", res, z0=z0, z1=z0 + slices, threads=threads)
    dt = time.perf_counter() - t
    return slices * res * res / dt, dt


def isosurface(dens, box=4000.0, full_res=None, iso=0.0, manifold=True):
    """dualmc (shared vertices) + the triangle tail of MarchingCubes::makeMesh, serial restatement (up_oracle_mesh.cpp)"""
    L = lib()
    L.upo_isosurface.restype = C.c_void_p
    L.upo_isosurface.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
    L.upo_isosurface_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3
    L.upo_isosurface_copy.argtypes = [C.c_void_p] * 6
    L.upo_isosurface_free.argtypes = [C.c_void_p]
    dens = np.ascontiguousarray(dens, dtype=np.float32)
    rz, ry, rx = dens.shape
    h = L.upo_isosurface(_p(dens), rx, ry, rz, iso, int(manifold), box, full_res or rx)
    nv, nq, ni = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.upo_isosurface_sizes(C.c_void_p(h), C.byref(nv), C.byref(nq), C.byref(ni))
    out = dict(
        dmc_vertices=np.zeros((nv.value, 3), np.float32),
        dmc_quads=np.zeros((nq.value, 4), np.int32),
        positions=np.zeros((nv.value, 3), np.float32),
        normals=np.zeros((nv.value, 3), np.float32),
        indices=np.zeros(ni.value, np.uint32),
    )
    L.upo_isosurface_copy(C.c_void_p(h), _p(out["dmc_vertices"]), _p(out["dmc_quads"]), _p(out["positions"]), _p(out["normals"]), _p(out["indices"]))
    L.upo_isosurface_free(C.c_void_p(h))
    return out


def texture_tail(special=None, height=None, strength=1.0):
    """numpy restatement (float32, operation order of the C++) of the generator's texture tail (sources/generator.cpp:135-139):
    imageConvertSpecialToGltfPbr (externals/cage/sources/libcore/image/imageAlgorithms.cpp:296-302): channels (0, special0, special1);
    imageConvertHeigthToNormal (:202-246) over a U8 image: clamped 3x3 Sobel of value = u8 / 255.f (image.cpp:220-236),
    v = normalize(-dX, -dY, 1 / strength), (v + 1) * 0.5, Image::set = uint8(saturate(v) * 255) truncating (image.cpp:468-481).
    Pinned against the reference's own output by tests/golden/texture_tail.npz (tests/test_oracle_port.py)."""
    f32 = np.float32
    pbr = nm = None
    if special is not None:
        sp = np.asarray(special, dtype=np.uint8)
        pbr = np.zeros(sp.shape[:2] + (3,), dtype=np.uint8)
        pbr[..., 1:] = sp
    if height is not None:
        hm = np.asarray(height, dtype=np.uint8)
        hm = hm.reshape(hm.shape[0], hm.shape[1])
        v = hm.astype(f32) / f32(255)
        p = np.pad(v, 1, mode="edge")
        H, W = v.shape
        tl, tc, tr = p[0:H, 0:W], p[0:H, 1 : W + 1], p[0:H, 2 : W + 2]
        lc, rc = p[1 : H + 1, 0:W], p[1 : H + 1, 2 : W + 2]
        bl, bc, br = p[2 : H + 2, 0:W], p[2 : H + 2, 1 : W + 1], p[2 : H + 2, 2 : W + 2]
        two = f32(2)
        dx = ((tr + two * rc) + br) - ((tl + two * lc) + bl)
        dy = ((bl + two * bc) + br) - ((tl + two * tc) + tr)
        x, y, z = -dx, -dy, np.full_like(dx, f32(1) / f32(strength))
        length = np.sqrt(((f32(0) + x * x) + y * y) + z * z)
        out = np.stack([x / length, y / length, z / length], axis=-1)
        out = (out + f32(1)) * f32(0.5)
        out = np.minimum(np.maximum(out, f32(0)), f32(1)) * f32(255)
        nm = out.astype(np.uint8)
    return pbr, nm
