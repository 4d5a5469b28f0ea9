"""`Planet`: Python face of upg_planet (include/upg.h), mirroring how the reference is configured:
`unnatural-planets --shape S --elevation E --colors C [--poles] [--flowers]` (sources/main.cpp:17-54) plus the
explicit seed the reference lacks."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import capi

# the reference's name tables, in order (sources/modesConfiguration.cpp:78-113, 176-183, 207-211)
SHAPES = ("asteroid", "belt", "bowl", "box", "bunny", "capsule", "cube", "disk", "doubletorus", "fibers", "gear", "h2o", "h3o", "h4o", "hemispheres", "hexagon", "hexagonalprism", "insidecube", "knot", "mandelbulb", "mobiusstrip", "monkeyhead", "octahedron", "pipe", "sphere", "square", "tetrahedron", "torus", "toruscross", "triangularprism", "tube", "twistedhexagonalprism", "twistedplane", "wormhole")
ELEVATIONS = ("none", "simple", "legacy", "lakes", "islands", "craters")
COLORINGS = ("default", "barren", "debug")
DEFAULT_SEED = (1346519564496, 42245614964156)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _mesh_to_numpy(lib, m):
    nv, ni, nq = m.vertex_count, m.index_count, m.quad_count

    def take(ptr, count, dtype):
        if count == 0 or not ptr:  # a cleaned mesh (K9) carries no dualmc arrays
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

    out = dict(
        positions=take(m.positions, nv * 3, np.float32).reshape(-1, 3),
        normals=take(m.normals, nv * 3, np.float32).reshape(-1, 3),
        indices=take(m.indices, ni, np.uint32),
        dmc_vertices=take(m.dmc_vertices, nv * 3, np.float32).reshape(-1, 3),
        dmc_quads=take(m.dmc_quads, nq * 4, np.uint32).reshape(-1, 4),
        near_zero_samples=int(m.near_zero_samples),
    )
    lib.upg_mesh_free(C.byref(m))
    return out


def isosurface(samples, box=4000.0, full_res=None, device=0):
    """K2..K4 over caller-provided samples [z][y][x] (upg_isosurface): dualmc + makeMesh's triangle tail on the GPU."""
    L = capi.load_library()
    s = np.ascontiguousarray(samples, dtype=np.float32)
    rz, ry, rx = s.shape
    m = capi.Mesh()
    capi.check(L.upg_isosurface(_ptr(s), rx, ry, rz, box, full_res or rx, device, C.byref(m)))
    return _mesh_to_numpy(L, m)


def slab_range(res, slab, slab_count):
    z0, z1 = C.c_int(0), C.c_int(0)
    capi.load_library().upg_slab_range(res, slab, slab_count, C.byref(z0), C.byref(z1))
    return z0.value, z1.value


def stitch_slabs(slabs):
    """concatenates field_mesh_slab() results of slabs 0..n-1: adds each slab's vertex base to its relative ids (mod 2^32)"""
    base = np.cumsum([0] + [len(s["positions"]) for s in slabs])
    out = {k: np.concatenate([s[k] for s in slabs]) for k in ("positions", "normals", "dmc_vertices")}
    out["indices"] = np.concatenate([s["indices"] + np.uint32(b) for s, b in zip(slabs, base)])
    out["dmc_quads"] = np.concatenate([s["dmc_quads"] + np.uint32(b) for s, b in zip(slabs, base)])
    out["near_zero_samples"] = sum(s["near_zero_samples"] for s in slabs)
    return out


def upload(arrays, device=0, stream=0):
    """host numpy arrays -> CUDA tensors on `device` through the library's staged multi-threaded copy (upg_upload): pageable memory at
    link speed instead of cudaMemcpy's ~11 GB/s.  uint32 arrays arrive as int32 tensors (same bits)."""
    import torch

    L = capi.load_library()
    arrays = [np.ascontiguousarray(a) for a in arrays]
    dev = torch.device("cuda", device)
    outs = [torch.empty(a.shape, dtype=torch.from_numpy(a[:0].view(np.int32) if a.dtype == np.uint32 else a[:0]).dtype, device=dev) for a in arrays]
    n = len(arrays)
    dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    src = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    nbytes = (C.c_uint64 * n)(*[a.nbytes for a in arrays])
    torch.cuda.current_stream(dev).synchronize()  # the tensors above were allocated on torch's stream
    capi.check(L.upg_upload(device, dst, src, nbytes, n, C.c_void_p(stream)))
    return outs


def raster(uv, idx, w, h, device=0):
    """K5 alone (upg_raster): owner image (0xffffffff = uncovered), barycentric weights of the winning visit, visit count."""
    L = capi.load_library()
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
    owner = np.zeros((h, w), dtype=np.uint32)
    weights = np.zeros((h, w, 3), dtype=np.float32)
    calls = C.c_uint64(0)
    capi.check(L.upg_raster(device, _ptr(uv), len(uv), _ptr(idx), len(idx), w, h, _ptr(owner), _ptr(weights), C.byref(calls)))
    return owner, weights, int(calls.value)


def dilate(img, rounds, use_nan=True, to_u8=False, device=0):
    """K7 alone (upg_dilate): imageDilation(img, rounds, use_nan) [+ imageConvert(U8)] on an (h, w, ch) float image."""
    L = capi.load_library()
    img = np.array(img, dtype=np.float32, order="C", copy=True)
    h, w, ch = img.shape
    u8 = np.zeros((h, w, ch), dtype=np.uint8) if to_u8 else None
    capi.check(L.upg_dilate(device, _ptr(img), w, h, ch, rounds, int(use_nan), _ptr(u8) if to_u8 else None))
    return (img, u8) if to_u8 else img


def mesh_clean(positions, normals, indices, box=4000.0, stages=7, device=0):
    """K9 alone (upg_mesh_clean) over a host triangle mesh -> dict(positions, normals, indices); stages = capi.MESH_* bits."""
    L = capi.load_library()
    pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    m = capi.Mesh()
    capi.check(L.upg_mesh_clean(device, _ptr(pos), _ptr(nrm), len(pos), _ptr(idx), len(idx), box, stages, C.byref(m)))
    return _mesh_to_numpy(L, m)


def mesh_chunking(positions, normals, indices, max_surface_area=800000.0, device=0):
    """cage's meshChunking as the generator's meshSplit calls it (upg_mesh_chunking) over a host triangle mesh -> list of
    dict(positions, normals, indices), in the reference's order (empty parts included)."""
    L = capi.load_library()
    pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    chunks, count = C.POINTER(capi.Mesh)(), C.c_uint32(0)
    capi.check(L.upg_mesh_chunking(device, _ptr(pos), _ptr(nrm), len(pos), _ptr(idx), len(idx), max_surface_area, C.byref(chunks), C.byref(count)))
    out = []
    for i in range(count.value):
        m = chunks[i]
        nv, ni = m.vertex_count, m.index_count
        out.append(dict(
            positions=(np.ctypeslib.as_array(m.positions, shape=(nv * 3,)).astype(np.float32, copy=True) if nv else np.zeros(0, np.float32)).reshape(-1, 3),
            normals=(np.ctypeslib.as_array(m.normals, shape=(nv * 3,)).astype(np.float32, copy=True) if nv else np.zeros(0, np.float32)).reshape(-1, 3),
            indices=np.ctypeslib.as_array(m.indices, shape=(ni,)).astype(np.uint32, copy=True) if ni else np.zeros(0, np.uint32),
        ))
    L.upg_mesh_chunks_free(chunks, count)
    return out


def texture_tail(special=None, height=None, strength=1.0, device=0):
    """K8 alone (upg_texture_tail): special (h,w,2) u8 -> pbr (h,w,3); height (h,w[,1]) u8 -> normal map (h,w,3)."""
    L = capi.load_library()
    sp = None if special is None else np.ascontiguousarray(special, dtype=np.uint8)
    hm = None if height is None else np.ascontiguousarray(height, dtype=np.uint8)
    h, w = (sp if sp is not None else hm).shape[:2]
    pbr = np.zeros((h, w, 3), dtype=np.uint8) if sp is not None else None
    nm = np.zeros((h, w, 3), dtype=np.uint8) if hm is not None else None
    capi.check(L.upg_texture_tail(device, _ptr(sp) if sp is not None else None, _ptr(hm) if hm is not None else None, w, h, strength, _ptr(pbr) if pbr is not None else None, _ptr(nm) if nm is not None else None))
    return pbr, nm


def reserve(device=0, device_bytes=0, pinned_block_bytes=0, pinned_blocks=0):
    """grows the library's device pool / pinned-block cache ahead of the first planet (upg_reserve)"""
    capi.check(capi.load_library().upg_reserve(device, device_bytes, pinned_block_bytes, pinned_blocks))


def trim(device=0, keep_bytes=0):
    capi.check(capi.load_library().upg_trim(device, keep_bytes))


def memory_stats(device=0):
    r, u, h = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    capi.check(capi.load_library().upg_memory_stats(device, C.byref(r), C.byref(u), C.byref(h)))
    return {"device_reserved": int(r.value), "device_used_high": int(u.value), "pinned_cached": int(h.value)}


def png_encode(image, device=0):
    """K10 (upg_png_encode): an (h, w[, ch]) uint8 image (host array) or a CUDA uint8 tensor of that shape -> bytes of a PNG file"""
    L = capi.load_library()
    if hasattr(image, "data_ptr"):  # torch tensor on the device
        h, w = int(image.shape[0]), int(image.shape[1])
        ch = int(image.shape[2]) if image.dim() == 3 else 1
        ptr = C.c_void_p(image.data_ptr())
        keep = image
    else:
        keep = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = keep.shape[:2]
        ch = keep.shape[2] if keep.ndim == 3 else 1
        ptr = _ptr(keep)
    cap = int(L.upg_png_bound(w, h, ch))
    out = np.empty(cap, dtype=np.uint8)
    size = C.c_uint64(0)
    capi.check(L.upg_png_encode(device, ptr, w, h, ch, _ptr(out), cap, C.byref(size)))
    return out[: size.value].tobytes()


class Planet:
    def __init__(self, shape="sphere", elevation="lakes", coloring="default", poles=False, flowers=False, seed=DEFAULT_SEED, grid=500, box=4000.0, device=0, mesh_sampling="band", arithmetic="exact", devices=None):
        L = capi.load_library()
        d = capi.PlanetDesc()
        L.upg_planet_desc_default(C.byref(d))
        sid, eid, cid = L.upg_shape_id(shape.encode()), L.upg_elevation_id(elevation.encode()), L.upg_coloring_id(coloring.encode())
        if sid < 0:
            raise ValueError("unknown shape mode configuration: %r" % shape)
        if eid < 0:
            raise ValueError("unknown elevation mode configuration: %r" % elevation)
        if cid < 0:
            raise ValueError("unknown coloring mode configuration: %r" % coloring)
        d.shape, d.elevation, d.coloring = sid, eid, cid
        d.poles, d.flowers = int(poles), int(flowers)
        d.seed[0], d.seed[1] = seed
        d.box_size = box
        d.grid_resolution = grid
        d.device = device
        d.mesh_sampling = capi.SAMPLING[mesh_sampling]
        d.arithmetic = capi.ARITHMETIC[arithmetic]
        if devices:  # multi-GPU inside one process: devices[0] is the home device
            d.device = devices[0]
            d.device_count = len(devices)
            for i, dev in enumerate(devices):
                d.devices[i] = dev
            device = devices[0]
        self.res = grid
        self.box = box
        self.device = device
        self._lib = L
        self._h = C.c_void_p()
        capi.check(L.upg_planet_create(C.byref(d), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.upg_planet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def async_status(self, stream=0):
        """raises UpgError(3) if an asynchronous sampling call produced a NaN since the last query (upg_planet_async_status)"""
        capi.check(self._lib.upg_planet_async_status(self._h, C.c_void_p(stream)))

    def seed_log(self):
        out = np.zeros(128, dtype=np.uint32)
        n = self._lib.upg_planet_seed_log(self._h, _ptr(out), 128)
        return out[:n].copy()

    # K1 ------------------------------------------------------------------------------------------
    def field_samples(self, field="land", z0=0, z1=None, out=None):
        """terrainSdf* over z-slices [z0, z1) of the grid -> host float32 array [z][y][x]."""
        z1 = self.res if z1 is None else z1
        if out is None:
            out = np.empty((z1 - z0, self.res, self.res), dtype=np.float32)
        capi.check(self._lib.upg_field_samples(self._h, capi.FIELDS[field], z0, z1, _ptr(out)))
        return out

    def field_samples_device(self, field, z0, z1, device_ptr, stream=0):
        """asynchronous, into device memory (an integer device pointer), on CUDA stream handle `stream`."""
        capi.check(self._lib.upg_field_samples_device(self._h, capi.FIELDS[field], z0, z1, C.c_void_p(device_ptr), C.c_void_p(stream)))

    def fields_samples_device(self, z0, z1, d_land=0, d_water=0, d_navigation=0, stream=0):
        """land / water / navigation of the same slices from one evaluation per sample (upg_fields_samples_device); device
        pointers as integers, 0 = field not wanted."""
        vp = C.c_void_p
        capi.check(self._lib.upg_fields_samples_device(self._h, z0, z1, vp(d_land or None), vp(d_water or None), vp(d_navigation or None), vp(stream)))

    def fields_samples_band_device(self, z0, z1, d_land=0, d_water=0, d_navigation=0, stream=0, count=False):
        """the samples the mesh entry points extract from (upg_fields_samples_band_device): exact wherever the isosurface can see
        the value, a value of the settled sign elsewhere; with count=True returns the number of nodes whose elevation was evaluated."""
        vp = C.c_void_p
        n = C.c_uint64(0)
        capi.check(self._lib.upg_fields_samples_band_device(self._h, z0, z1, vp(d_land or None), vp(d_water or None), vp(d_navigation or None), vp(stream), C.byref(n) if count else None))
        return int(n.value) if count else None

    # K1 + K2..K4 ---------------------------------------------------------------------------------
    def field_meshes(self, fields=("land", "water", "navigation")):
        """the base meshes of several fields from one fused sampling pass (upg_field_meshes) -> {field: mesh dict}."""
        ms = (capi.Mesh * 3)()
        mask = 0
        for f in fields:
            if capi.FIELDS[f] > 2:
                raise ValueError("field_meshes: land, water and navigation only")
            mask |= 1 << capi.FIELDS[f]
        t = time.perf_counter()
        capi.check(self._lib.upg_field_meshes(self._h, mask, ms))
        self.last_call_s = time.perf_counter() - t
        return {f: _mesh_to_numpy(self._lib, ms[capi.FIELDS[f]]) for f in fields}

    def field_mesh(self, field="land"):
        """meshGenerateGeneric up to makeMesh's normals (upg_field_mesh) -> dict of numpy arrays."""
        m = capi.Mesh()
        capi.check(self._lib.upg_field_mesh(self._h, capi.FIELDS[field], C.byref(m)))
        return _mesh_to_numpy(self._lib, m)

    def field_meshes_clean(self, fields=("land", "water", "navigation"), stages=capi.MESH_GENERATOR):
        """upg_field_meshes_clean: fused sampling + extraction + K9 clean-up; with MESH_GENERATOR the results are what
        meshGenerateBaseLand / Water / Navigation return (sources/meshGeneration.cpp:47-107)."""
        ms = (capi.Mesh * 3)()
        mask = 0
        for f in fields:
            if capi.FIELDS[f] > 2:
                raise ValueError("field_meshes_clean: land, water and navigation only")
            mask |= 1 << capi.FIELDS[f]
        t = time.perf_counter()
        capi.check(self._lib.upg_field_meshes_clean(self._h, mask, stages, ms))
        self.last_call_s = time.perf_counter() - t  # the C call alone (the numpy copies below are the binding's, not the library's)
        return {f: _mesh_to_numpy(self._lib, ms[capi.FIELDS[f]]) for f in fields}

    def _field_mask(self, fields):
        mask = 0
        for f in fields:
            if capi.FIELDS[f] > 2:
                raise ValueError("land, water and navigation only")
            mask |= 1 << capi.FIELDS[f]
        return mask

    def field_meshes_slab(self, fields, slab, slab_count):
        """slab `slab` of `slab_count` of several fields from one fused sampling pass over the slab (upg_field_meshes_slab)."""
        ms = (capi.Mesh * 3)()
        capi.check(self._lib.upg_field_meshes_slab(self._h, self._field_mask(fields), slab, slab_count, ms))
        return {f: _mesh_to_numpy(self._lib, ms[capi.FIELDS[f]]) for f in fields}

    def field_meshes_slab_device(self, fields, slab, slab_count):
        """the same, left in HBM (upg_field_meshes_slab_device): {field: dict of CUDA tensors + near_zero_samples}.  The tensors are
        torch-owned copies (one device-to-device copy of the chunk), so the library's arrays are released before returning."""
        import torch

        ms = (capi.DeviceMesh * 3)()
        capi.check(self._lib.upg_field_meshes_slab_device(self._h, self._field_mask(fields), slab, slab_count, ms))
        dev = torch.device("cuda", self.device)

        class _View:  # zero-copy view of library memory through the CUDA array interface
            def __init__(self, ptr, count, typestr):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}

        def take(ptr, count, typestr, dtype):
            if not ptr or count == 0:
                return torch.zeros(0, dtype=dtype, device=dev)
            return torch.as_tensor(_View(ptr, count, typestr), device=dev).clone()

        out = {}
        with torch.cuda.device(dev):
            for f in fields:
                m = ms[capi.FIELDS[f]]
                nv, ni, nq = m.vertex_count, m.index_count, m.quad_count
                out[f] = dict(
                    positions=take(m.positions, nv * 3, "<f4", torch.float32).view(-1, 3),
                    normals=take(m.normals, nv * 3, "<f4", torch.float32).view(-1, 3),
                    dmc_vertices=take(m.dmc_vertices, nv * 3, "<f4", torch.float32).view(-1, 3),
                    indices=take(m.indices, ni, "<i4", torch.int32),
                    dmc_quads=take(m.dmc_quads, nq * 4, "<i4", torch.int32).view(-1, 4),
                    near_zero_samples=int(m.near_zero_samples),
                )
            torch.cuda.current_stream().synchronize()  # the copies above are done before the library frees its arrays
            for k in range(3):
                self._lib.upg_device_mesh_free(C.byref(ms[k]))
        return out

    def field_mesh_slab(self, field, slab, slab_count):
        """slab `slab` of `slab_count` of field_mesh() (upg_field_mesh_slab): indices / dmc_quads are relative to the slab's
        first vertex (uint32, two's complement); see stitch_slabs()."""
        m = capi.Mesh()
        capi.check(self._lib.upg_field_mesh_slab(self._h, capi.FIELDS[field], slab, slab_count, C.byref(m)))
        return _mesh_to_numpy(self._lib, m)

    def field_points(self, field, positions):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.float32)
        capi.check(self._lib.upg_field_points(self._h, capi.FIELDS[field], _ptr(pos), len(pos), _ptr(out)))
        return out

    # K5..K7 --------------------------------------------------------------------------------------
    def set_cellular_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float32).reshape(-1)
        assert t.size == 1024
        capi.check(self._lib.upg_planet_set_cellular_table(self._h, _ptr(t)))

    def tiles(self, purpose, positions, normals):
        """terrainTile on explicit (position, unit normal) pairs -> (n, 13) float32, columns capi.TILE_COLUMNS."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(pos), 13), dtype=np.float32)
        capi.check(self._lib.upg_tiles(self._h, capi.PURPOSES[purpose], _ptr(pos), _ptr(nrm), len(pos), _ptr(out)))
        return out

    def bake(self, water, positions, normals, uvs, indices, w, h):
        """generateTexturesLand / Water (upg_bake) -> (albedo u8 (h,w,3|4), special u8 (h,w,2), height u8 (h,w,1))."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        albedo = np.zeros((h, w, 4 if water else 3), dtype=np.uint8)
        special = np.zeros((h, w, 2), dtype=np.uint8)
        height = np.zeros((h, w, 1), dtype=np.uint8)
        capi.check(self._lib.upg_bake(self._h, int(bool(water)), _ptr(pos), _ptr(nrm), _ptr(uv), len(pos), _ptr(idx), len(idx), w, h, _ptr(albedo), _ptr(special), _ptr(height)))
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())
        return albedo, special, height

    def bake_rows(self, water, positions, normals, uvs, indices, w, h, row0, row1):
        """rows [row0, row1) of the textures bake() returns (upg_bake_rows): the unit of texel-row sharding across GPUs."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        rows = row1 - row0
        albedo = np.zeros((rows, w, 4 if water else 3), dtype=np.uint8)
        special = np.zeros((rows, w, 2), dtype=np.uint8)
        height = np.zeros((rows, w, 1), dtype=np.uint8)
        capi.check(self._lib.upg_bake_rows(self._h, int(bool(water)), _ptr(pos), _ptr(nrm), _ptr(uv), len(pos), _ptr(idx), len(idx), w, h, row0, row1, _ptr(albedo), _ptr(special), _ptr(height)))
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())
        return albedo, special, height

    def bake_export(self, water, positions, normals, uvs, indices, w, h, row0=0, row1=None):
        """bake + texture tail (upg_bake_export_rows): the three PNG payloads of a chunk -> (albedo, pbr (rows,w,3), normal map (rows,w,3))."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        row1 = h if row1 is None else row1
        rows = row1 - row0
        albedo = np.zeros((rows, w, 4 if water else 3), dtype=np.uint8)
        pbr = np.zeros((rows, w, 3), dtype=np.uint8)
        nmap = np.zeros((rows, w, 3), dtype=np.uint8)
        capi.check(self._lib.upg_bake_export_rows(self._h, int(bool(water)), _ptr(pos), _ptr(nrm), _ptr(uv), len(pos), _ptr(idx), len(idx), w, h, row0, row1, _ptr(albedo), _ptr(pbr), _ptr(nmap)))
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())
        return albedo, pbr, nmap

    def bake_export_png(self, water, positions, normals, uvs, indices, w, h, out=None):
        """upg_bake_export_png: the three PNG files of a chunk (albedo, pbr, normal map) as bytes.  `out`: three reusable uint8 host
        arrays of at least upg_png_bound bytes (e.g. pinned torch tensors' numpy views); allocated when absent."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        chans = (4 if water else 3, 3, 3)
        caps = [int(self._lib.upg_png_bound(w, h, c)) for c in chans]
        if out is None:
            out = [np.empty(c, dtype=np.uint8) for c in caps]
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in out])
        cap = (C.c_uint64 * 3)(*[min(c, o.size) for c, o in zip(caps, out)])
        size = (C.c_uint64 * 3)()
        t = time.perf_counter()
        capi.check(self._lib.upg_bake_export_png(self._h, int(bool(water)), _ptr(pos), _ptr(nrm), _ptr(uv), len(pos), _ptr(idx), len(idx), w, h, ptrs, cap, size))
        self.last_call_s = time.perf_counter() - t
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())
        return [o[: int(n)] for o, n in zip(out, size)]

    def bake_export_device(self, water, d_pos, d_nrm, d_uv, nverts, d_idx, nidx, w, h, row0, row1, d_albedo, d_pbr, d_normal, stream=0):
        """upg_bake_export_device: every pointer is an integer DEVICE address; synchronises `stream` before returning."""
        vp = C.c_void_p
        capi.check(self._lib.upg_bake_export_device(self._h, int(bool(water)), vp(d_pos), vp(d_nrm), vp(d_uv), nverts, vp(d_idx), nidx, w, h, row0, row1, vp(d_albedo), vp(d_pbr), vp(d_normal), vp(stream)))
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())

    def bake_device(self, water, d_pos, d_nrm, d_uv, nverts, d_idx, nidx, w, h, row0, row1, d_albedo, d_special, d_height, stream=0):
        """upg_bake_device: every pointer is an integer DEVICE address; synchronises `stream` before returning."""
        vp = C.c_void_p
        capi.check(self._lib.upg_bake_device(self._h, int(bool(water)), vp(d_pos), vp(d_nrm), vp(d_uv), nverts, vp(d_idx), nidx, w, h, row0, row1, vp(d_albedo), vp(d_special), vp(d_height), vp(stream)))
        self.last_covered_texels = int(self._lib.upg_bake_last_covered_texels())
