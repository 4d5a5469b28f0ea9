"""ctypes declarations for include/upg.h.  Loading is lazy and LOUD: a missing library is an error, never a fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

FIELDS = {"land": 0, "water": 1, "navigation": 2, "elevation_raw": 3, "elevation": 4}


class UpgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("upg error %d: %s" % (code, message))
        self.code = code


class PlanetDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("shape", C.c_int32),
        ("elevation", C.c_int32),
        ("coloring", C.c_int32),
        ("poles", C.c_int32),
        ("flowers", C.c_int32),
        ("seed", C.c_uint64 * 2),
        ("box_size", C.c_float),
        ("grid_resolution", C.c_int32),
        ("device", C.c_int32),
        ("mesh_sampling", C.c_int32),
        ("arithmetic", C.c_int32),
        ("device_count", C.c_int32),
        ("devices", C.c_int32 * 8),
    ]


SAMPLING = {"band": 0, "dense": 1}
ARITHMETIC = {"exact": 0, "contract": 1}


class Mesh(C.Structure):
    _fields_ = [
        ("positions", C.POINTER(C.c_float)),
        ("normals", C.POINTER(C.c_float)),
        ("indices", C.POINTER(C.c_uint32)),
        ("dmc_vertices", C.POINTER(C.c_float)),
        ("dmc_quads", C.POINTER(C.c_uint32)),
        ("vertex_count", C.c_uint32),
        ("index_count", C.c_uint32),
        ("quad_count", C.c_uint32),
        ("reserved", C.c_uint32),
        ("near_zero_samples", C.c_uint64),
        ("owner", C.c_void_p),
        ("owner_bytes", C.c_uint64),
    ]


class DeviceMesh(C.Structure):
    _fields_ = [
        ("positions", C.c_void_p),
        ("normals", C.c_void_p),
        ("indices", C.c_void_p),
        ("dmc_vertices", C.c_void_p),
        ("dmc_quads", C.c_void_p),
        ("vertex_count", C.c_uint32),
        ("index_count", C.c_uint32),
        ("quad_count", C.c_uint32),
        ("device", C.c_int32),
        ("near_zero_samples", C.c_uint64),
        ("owner", C.c_void_p),
    ]


class Tile(C.Structure):
    _fields_ = [
        ("albedo", C.c_float * 3),
        ("roughness", C.c_float),
        ("metallic", C.c_float),
        ("height", C.c_float),
        ("elevation", C.c_float),
        ("slope", C.c_float),
        ("temperature", C.c_float),
        ("precipitation", C.c_float),
        ("opacity", C.c_float),
        ("biome", C.c_float),
        ("type", C.c_float),
    ]


# K9 stages (include/upg.h UPG_MESH_*)
MESH_MAKEMESH_TAIL, MESH_REMOVE_DISCONNECTED, MESH_FLIP_NORMALS, MESH_WATER_CULL, MESH_GENERATOR = 1, 2, 4, 8, 15

PURPOSES = {"land": 1, "water": 2, "navigation": 3}
TILE_COLUMNS = ("albedo_r", "albedo_g", "albedo_b", "roughness", "metallic", "height", "elevation", "slope", "temperature", "precipitation", "opacity", "biome", "type")


def library_path() -> str:
    # UPG_LIBRARY: an alternative build of the same library (A/B measurements of kernel variants); never a fallback
    return os.environ.get("UPG_LIBRARY") or os.path.join(HERE, "libupg.so")


_lib = None


def load_library():
    """Returns the loaded libupg.so.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("libupg.so is not built: run `make -C unnatural_planets_b200/csrc -j8` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, u64, fp = C.c_void_p, C.c_int, C.c_uint64, C.c_void_p
    L.upg_last_error.restype = C.c_char_p
    L.upg_planet_desc_default.argtypes = [C.POINTER(PlanetDesc)]
    L.upg_planet_desc_default.restype = None
    for n in ("upg_shape_id", "upg_elevation_id", "upg_coloring_id"):
        getattr(L, n).argtypes = [C.c_char_p]
    for n in ("upg_shape_name", "upg_elevation_name", "upg_coloring_name"):
        getattr(L, n).argtypes = [i32]
        getattr(L, n).restype = C.c_char_p
    L.upg_planet_create.argtypes = [C.POINTER(PlanetDesc), C.POINTER(vp)]
    L.upg_planet_destroy.argtypes = [vp]
    L.upg_planet_destroy.restype = None
    L.upg_planet_seed_log.argtypes = [vp, fp, i32]
    L.upg_planet_async_status.argtypes = [vp, vp]
    L.upg_kernel_param_bytes.restype = C.c_uint32
    L.upg_reserve.argtypes = [i32, u64, u64, i32]
    L.upg_trim.argtypes = [i32, u64]
    L.upg_memory_stats.argtypes = [i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.upg_field_samples.argtypes = [vp, i32, i32, i32, fp]
    L.upg_field_samples_device.argtypes = [vp, i32, i32, i32, fp, vp]
    L.upg_fields_samples_device.argtypes = [vp, i32, i32, fp, fp, fp, vp]
    L.upg_upload.argtypes = [i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), i32, vp]
    L.upg_fields_samples_band_device.argtypes = [vp, i32, i32, fp, fp, fp, vp, C.POINTER(C.c_uint64)]
    L.upg_field_meshes.argtypes = [vp, C.c_uint, C.POINTER(Mesh)]
    L.upg_field_meshes_slab.argtypes = [vp, C.c_uint, i32, i32, C.POINTER(Mesh)]
    L.upg_field_meshes_slab_device.argtypes = [vp, C.c_uint, i32, i32, C.POINTER(DeviceMesh)]
    L.upg_device_mesh_free.argtypes = [C.POINTER(DeviceMesh)]
    L.upg_device_mesh_free.restype = None
    L.upg_field_meshes_clean.argtypes = [vp, C.c_uint, C.c_uint, C.POINTER(Mesh)]
    L.upg_mesh_clean.argtypes = [i32, fp, fp, C.c_uint32, fp, C.c_uint32, C.c_float, C.c_uint, C.POINTER(Mesh)]
    L.upg_mesh_chunking.argtypes = [i32, fp, fp, C.c_uint32, fp, C.c_uint32, C.c_float, C.POINTER(C.POINTER(Mesh)), C.POINTER(C.c_uint32)]
    L.upg_mesh_chunks_free.argtypes = [C.POINTER(Mesh), C.c_uint32]
    L.upg_mesh_chunks_free.restype = None
    L.upg_field_points.argtypes = [vp, i32, fp, u64, fp]
    L.upg_field_mesh.argtypes = [vp, i32, C.POINTER(Mesh)]
    L.upg_field_mesh_slab.argtypes = [vp, i32, i32, i32, C.POINTER(Mesh)]
    L.upg_slab_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.upg_slab_range.restype = None
    L.upg_isosurface.argtypes = [fp, i32, i32, i32, C.c_float, i32, i32, C.POINTER(Mesh)]
    L.upg_mesh_free.argtypes = [C.POINTER(Mesh)]
    L.upg_mesh_free.restype = None
    L.upg_measure_fp32_peak.argtypes = [i32, C.POINTER(C.c_float)]
    u32 = C.c_uint32
    L.upg_planet_set_cellular_table.argtypes = [vp, fp]
    L.upg_tiles.argtypes = [vp, i32, fp, fp, u64, fp]
    L.upg_bake.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, fp, fp, fp]
    L.upg_bake_last_covered_texels.restype = u64
    L.upg_bake_rows.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, u32, u32, fp, fp, fp]
    L.upg_bake_device.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, u32, u32, fp, fp, fp, vp]
    L.upg_bake_export.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, fp, fp, fp]
    L.upg_bake_export_rows.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, u32, u32, fp, fp, fp]
    L.upg_bake_export_device.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, u32, u32, fp, fp, fp, vp]
    L.upg_texture_tail.argtypes = [i32, fp, fp, u32, u32, C.c_float, fp, fp]
    L.upg_raster.argtypes = [i32, fp, u32, fp, u32, u32, u32, fp, fp, C.POINTER(u64)]
    L.upg_dilate.argtypes = [i32, fp, u32, u32, u32, u32, i32, fp]
    L.upg_png_bound.argtypes = [u32, u32, u32]
    L.upg_png_bound.restype = u64
    L.upg_png_encode.argtypes = [i32, fp, u32, u32, u32, fp, u64, C.POINTER(u64)]
    L.upg_bake_export_png.argtypes = [vp, i32, fp, fp, fp, u32, fp, u32, u32, u32, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise UpgError(rc, load_library().upg_last_error().decode())
