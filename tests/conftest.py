import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_registry(npz):
    return json.loads(bytes(npz["registry"]).decode())


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def hostlib():
    """TEST-ONLY host compilation of the product's device headers (tests/hostlib)."""
    d = os.path.join(ROOT, "tests", "hostlib")
    subprocess.check_call(["make", "-s", "-C", d])
    H = C.CDLL(os.path.join(d, "libupg_hosttest.so"))
    H.upgh_create.restype = C.c_void_p
    H.upgh_create.argtypes = [C.c_int] * 5 + [C.c_uint64] * 2 + [C.c_float, C.c_int]
    H.upgh_destroy.argtypes = [C.c_void_p]
    H.upgh_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    H.upgh_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    H.upgh_grid_coords.argtypes = [C.c_void_p, C.c_void_p]
    H.upgh_seed_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    H.upgh_noise3.restype = C.c_float
    H.upgh_noise3.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    return H


class HostPlanet:
    """drives tests/hostlib like unnatural_planets_b200.Planet drives libupg.so"""

    FIELDS = {"land": 0, "water": 1, "navigation": 2, "elevraw": 3, "elevation": 4}

    def __init__(self, H, shape, elevation, seed=(1346519564496, 42245614964156), res=500, box=4000.0):
        from unnatural_planets_b200.planet import ELEVATIONS, SHAPES

        self.H = H
        self.res = res
        self.h = H.upgh_create(SHAPES.index(shape), ELEVATIONS.index(elevation), 0, 0, 0, seed[0], seed[1], box, res)

    def points(self, field, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        out = np.zeros(len(pts), dtype=np.float32)
        self.H.upgh_points(C.c_void_p(self.h), self.FIELDS[field], pts.ctypes.data_as(C.c_void_p), len(pts), out.ctypes.data_as(C.c_void_p))
        return out

    def grid(self, field, z0=0, z1=None):
        z1 = self.res if z1 is None else z1
        out = np.zeros((z1 - z0, self.res, self.res), dtype=np.float32)
        self.H.upgh_grid(C.c_void_p(self.h), self.FIELDS[field], z0, z1, out.ctypes.data_as(C.c_void_p))
        return out

    def coords(self):
        out = np.zeros(self.res, dtype=np.float32)
        self.H.upgh_grid_coords(C.c_void_p(self.h), out.ctypes.data_as(C.c_void_p))
        return out

    def seed_log(self):
        out = np.zeros(128, dtype=np.uint32)
        n = self.H.upgh_seed_log(C.c_void_p(self.h), out.ctypes.data_as(C.c_void_p), 128)
        return out[:n]

    def __del__(self):
        try:
            self.H.upgh_destroy(C.c_void_p(self.h))
        except Exception:
            pass


def same_bits(a, b):
    """bit-identical floats, treating +0 == -0 and NaN == NaN"""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))


# FP32 tolerance of GPU results against the reference (north_star: "<=1e-4 relative on SDF"): the only
# operations that are not bit-identical are libm transcendentals (powf, sinf, cosf, atan2f, acosf, logf);
# values near zero come from cancellation of O(1000) magnitudes, hence the absolute term (world units; one
# grid cell of the default grid is 8.08 units).
SDF_RTOL = 1e-4
SDF_ATOL = 2e-3


def sdf_close(a, b, rtol=SDF_RTOL, atol=SDF_ATOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol
