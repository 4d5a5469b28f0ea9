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
