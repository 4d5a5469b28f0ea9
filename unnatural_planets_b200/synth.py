"""Synthetic inputs for bench.py and the large-size tests.  NOT part of the hot path: the reference obtains texture
coordinates from xatlas (externals/cage/sources/libcore/mesh/xatlas.cpp:54-141 via meshUnwrap, sources/meshGeneration.cpp:
184-197), a serial third-party unwrapper that is out of scope.  To bake a whole surface at a stated
resolution without it, vertices get an octahedral parametrisation of their direction from the origin: one chart, area
distortion <= 2x, every triangle that would straddle a fold of the octahedron is dropped (a few 0.1 % of the surface)."""
import numpy as np


def octahedral_uv(positions, margin=0.004):
    d = np.asarray(positions, dtype=np.float64)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    p = d / np.abs(d).sum(axis=1, keepdims=True)
    x, y = p[:, 0].copy(), p[:, 1].copy()
    lower = p[:, 2] < 0
    fx = (1 - np.abs(y)) * np.where(x >= 0, 1.0, -1.0)
    fy = (1 - np.abs(x)) * np.where(y >= 0, 1.0, -1.0)
    x[lower], y[lower] = fx[lower], fy[lower]
    uv = np.stack([x, y], axis=1) * (0.5 - margin) + 0.5
    return uv.astype(np.float32)


def drop_folded_triangles(uv, indices, limit=0.03):
    tri = np.asarray(indices, dtype=np.uint32).reshape(-1, 3)
    a, b, c = uv[tri[:, 0]], uv[tri[:, 1]], uv[tri[:, 2]]
    longest = np.maximum(np.maximum(np.abs(a - b).max(axis=1), np.abs(b - c).max(axis=1)), np.abs(c - a).max(axis=1))
    return np.ascontiguousarray(tri[longest < limit].reshape(-1))


def uv_window_submesh(positions, normals, uv, indices, u0, v0, size):
    """triangles whose three corners lie in the UV window [u0, u0+size) x [v0, v0+size), UVs rescaled to the unit square:
    baking it at (size * W)^2 evaluates exactly the texels of that window of the full W^2 bake (a bounded CPU sample)"""
    tri = np.asarray(indices, dtype=np.uint32).reshape(-1, 3)
    inside = np.ones(len(tri), dtype=bool)
    for k in range(3):
        q = uv[tri[:, k]]
        inside &= (q[:, 0] >= u0) & (q[:, 0] < u0 + size) & (q[:, 1] >= v0) & (q[:, 1] < v0 + size)
    tri = tri[inside]
    used, inv = np.unique(tri.reshape(-1), return_inverse=True)
    suv = ((uv[used] - np.array([u0, v0], dtype=np.float32)) / np.float32(size)).astype(np.float32)
    return positions[used], normals[used], np.clip(suv, 0, 0.99999), inv.astype(np.uint32)
