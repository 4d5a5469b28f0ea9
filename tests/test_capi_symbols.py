"""The C-ABI library loads without a GPU and exports every symbol include/upg.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "upg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(upg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("upg_planet_create", "upg_planet_destroy", "upg_field_samples", "upg_field_points", "upg_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from unnatural_planets_b200 import capi

    path = capi.library_path()
    if not os.path.exists(path):
        pytest.fail("libupg.so missing: run `python __graft_entry__.py build`")
    L = C.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_mode_tables_follow_reference_order():
    from unnatural_planets_b200 import COLORINGS, ELEVATIONS, SHAPES, capi

    L = capi.load_library()
    assert L.upg_shape_count() == 34 == len(SHAPES)
    for i, n in enumerate(SHAPES):
        assert L.upg_shape_id(n.encode()) == i and L.upg_shape_name(i).decode() == n
    for i, n in enumerate(ELEVATIONS):
        assert L.upg_elevation_id(n.encode()) == i and L.upg_elevation_name(i).decode() == n
    for i, n in enumerate(COLORINGS):
        assert L.upg_coloring_id(n.encode()) == i
    assert L.upg_shape_id(b"cubee") == -1  # the reference throws "unknown shape mode configuration"
    assert SHAPES.index("sphere") == 24 and SHAPES.index("torus") == 27


def test_argument_errors_need_no_gpu():
    from unnatural_planets_b200 import capi

    L = capi.load_library()
    d = capi.PlanetDesc()
    L.upg_planet_desc_default(C.byref(d))
    assert d.grid_resolution == 500 and d.box_size == 4000.0 and d.elevation == 3
    h = C.c_void_p()
    d.grid_resolution = 5  # cage asserts resolution > 5 (marchingCubes.cpp:21-23)
    assert L.upg_planet_create(C.byref(d), C.byref(h)) == 1
    assert b"resolution" in L.upg_last_error()
    d.grid_resolution = 500
    d.shape = 99
    assert L.upg_planet_create(C.byref(d), C.byref(h)) == 1
    assert b"unknown shape" in L.upg_last_error()
