"""Deterministic synthetic inputs for the texel-baking tests (shared by the tests and tests/golden/make_golden_bake.py).

Land tiles take their elevation from the base shape alone (terrainSdfElevation = shape * 10, sources/coloringModes.cpp:51-57),
so a patch of a radius-1100 sphere displaced radially by d world units sits at tile elevation 10 * d: the patches below are
placed under water, on the beach, in the temperate band and above the snow line to reach every colouring layer."""
import numpy as np

# configurations of the colouring fixtures: (shape, elevation, colouring, poles, flowers)
TILE_CONFIGS = (
    ("sphere", "lakes", "default", False, False),
    ("sphere", "lakes", "default", True, True),
    ("sphere", "islands", "barren", False, True),
    ("sphere", "simple", "debug", True, False),
    ("torus", "craters", "default", True, True),
)
PURPOSES = ("land", "water", "navigation")
BAKE_CONFIGS = (
    ("sphere", "lakes", "default", False, False),
    ("sphere", "lakes", "default", True, True),
    ("sphere", "legacy", "barren", False, True),
    ("sphere", "none", "debug", True, False),
)


def config_name(cfg):
    s, e, c, p, f = cfg
    return "%s_%s_%s_p%d_f%d" % (s, e, c, int(p), int(f))


def tile_probes(shape, n=1500, seed=11):
    """(positions, unit normals): points around the surface of `shape` (sphere r=1100 / torus R=950 r=530)"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    off = rng.uniform(-12, 35, size=(n, 1))
    off[: n // 4] = rng.uniform(-1.5, 3.0, size=(n // 4, 1))  # shoreline / beach band
    if shape == "torus":
        # torus around the z axis in the xy plane (sdfTorus: major 950, minor 530)
        ang = rng.uniform(0, 2 * np.pi, size=n)
        radial = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], axis=1)
        phi = rng.uniform(0, 2 * np.pi, size=(n, 1))
        d = radial * np.cos(phi) + np.array([[0.0, 0.0, 1.0]]) * np.sin(phi)
        pos = radial * 950 + d * (530 + off)
        nrm = d
    else:
        pos = d * (1100 + off)
        nrm = d
    tilt = rng.normal(size=(n, 3)) * rng.choice([0.02, 0.15, 0.4], size=(n, 1))
    nrm = nrm + tilt
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return pos.astype(np.float32), nrm.astype(np.float32)


def _patch(center_dir, up, radial_offset, size, n, uv0, uv1, bump):
    """n x n vertex grid on a sphere patch of `size` world units around center_dir"""
    c = np.asarray(center_dir, dtype=np.float64)
    c /= np.linalg.norm(c)
    a = np.cross(c, up)
    a /= np.linalg.norm(a)
    b = np.cross(c, a)
    s, t = np.meshgrid(np.linspace(-0.5, 0.5, n), np.linspace(-0.5, 0.5, n), indexing="xy")
    d = c[None, None, :] * 1100 + a[None, None, :] * (s[..., None] * size) + b[None, None, :] * (t[..., None] * size)
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    r = 1100 + radial_offset + bump * (np.sin(s * 19.0) * np.cos(t * 13.0) + 0.5 * np.sin((s + t) * 41.0))
    pos = d * r[..., None]
    uv = np.stack([uv0[0] + (s + 0.5) * (uv1[0] - uv0[0]), uv0[1] + (t + 0.5) * (uv1[1] - uv0[1])], axis=2)
    idx = []
    for j in range(n - 1):
        for i in range(n - 1):
            v = j * n + i
            idx += [v, v + 1, v + n, v + 1, v + n + 1, v + n]
    return pos.reshape(-1, 3), uv.reshape(-1, 2), np.asarray(idx, dtype=np.uint32)


def vertex_normals(pos, idx, outward=True):
    tri = idx.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    nrm = np.zeros_like(pos)
    for k in range(3):
        np.add.at(nrm, tri[:, k], fn)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    if outward:
        flip = np.sum(nrm * pos, axis=1) < 0
        nrm[flip] *= -1
    return nrm


def atlas_mesh(texels=192, n=20):
    """four sphere patches in a 2x2 atlas (+ one sliver chart that overlaps the first: 'the last triangle wins').
    Patch size is chosen for ~1.35 texels per world unit (sources/meshGeneration.cpp:193) at `texels`^2."""
    size = (texels * 0.44) / 1.35
    charts = [
        ((1.0, 0.2, 0.1), -6.0, 3.0),  # under water .. shoreline
        ((-0.3, 0.9, 0.3), 1.0, 1.2),  # beach band, near a pole
        ((0.2, -0.1, 1.0), 9.0, 4.0),  # temperate land with cliffs
        ((-0.7, -0.6, -0.4), 25.0, 2.5),  # above the snow line
    ]
    P, U, I = [], [], []
    base = 0
    for k, (c, off, bump) in enumerate(charts):
        cx, cy = k % 2, k // 2
        uv0 = (0.03 + 0.5 * cx, 0.03 + 0.5 * cy)
        uv1 = (0.47 + 0.5 * cx, 0.47 + 0.5 * cy)
        p, u, i = _patch(c, (0.0, 0.0, 1.0) if abs(c[2]) < 0.9 else (0.0, 1.0, 0.0), off, size, n, uv0, uv1, bump)
        P.append(p)
        U.append(u)
        I.append(i + base)
        base += len(p)
    # an extra sliver chart drawn LAST over the corner of chart 0 (different surface position: visible if the order is wrong)
    p, u, i = _patch((0.1, 0.3, -1.0), (0.0, 1.0, 0.0), 12.0, size * 0.2, 5, (0.02, 0.02), (0.13, 0.11), 1.0)
    P.append(p)
    U.append(u)
    I.append(i + base)
    pos = np.concatenate(P)
    uv = np.concatenate(U)
    idx = np.concatenate(I)
    nrm = vertex_normals(pos, idx)
    return pos.astype(np.float32), nrm.astype(np.float32), uv.astype(np.float32), idx.astype(np.uint32)


def random_triangles(count=400, seed=3):
    """triangle soup in UV space for the rasteriser: slivers, flat tops / bottoms, sub-texel, overlapping, edge-touching"""
    rng = np.random.default_rng(seed)
    uv = rng.random((count * 3, 2))
    scale = rng.choice([0.01, 0.05, 0.2, 1.0], size=(count, 1, 1))
    c = uv.reshape(count, 3, 2).mean(axis=1, keepdims=True)
    uv = (c + (uv.reshape(count, 3, 2) - c) * scale).reshape(-1, 2)
    uv[0:3] = [(0.1, 0.5), (0.3, 0.5), (0.2, 0.7)]  # flat bottom
    uv[3:6] = [(0.5, 0.2), (0.6, 0.4), (0.7, 0.4)]  # flat top
    uv[6:9] = [(0.4, 0.4), (0.4, 0.4), (0.4, 0.4)]  # a point
    uv[9:12] = [(0.1, 0.1), (0.9, 0.1001), (0.5, 0.1002)]  # horizontal sliver
    uv[12:15] = [(0.0, 0.0), (0.999, 0.0), (0.0, 0.999)]  # touches the borders
    uv = np.clip(uv, 0.0, 0.9995).astype(np.float32)
    idx = np.arange(count * 3, dtype=np.uint32)
    return uv, idx


def dilation_image(h=70, w=90, ch=3, seed=5, hole=0.6):
    rng = np.random.default_rng(seed)
    img = rng.random((h, w, ch)).astype(np.float32)
    mask = rng.random((h, w)) < hole
    yy, xx = np.mgrid[0:h, 0:w]
    mask |= ((yy - h / 2) ** 2 + (xx - w / 2) ** 2) < (min(h, w) / 4) ** 2  # a hole wider than 7 rounds can fill
    img[mask] = np.nan
    return img


def tail_images():
    """[(name, special u8 (h,w,2), height u8 (h,w,1))]: inputs of the texture-tail fixture (texture_tail.npz): degenerate and ragged
    sizes (every border clamp of the Sobel), random bytes (every quantisation bucket) and a smooth relief with flat areas"""
    rng = np.random.default_rng(77)
    out = []
    for h, w in ((1, 1), (1, 9), (9, 1), (2, 2), (3, 3), (37, 23), (64, 130)):
        out.append(("rand_%dx%d" % (h, w), rng.integers(0, 256, (h, w, 2), dtype=np.uint8), rng.integers(0, 256, (h, w, 1), dtype=np.uint8)))
    yy, xx = np.mgrid[0:96, 0:80]
    relief = 127 + 90 * np.sin(xx / 9.0) * np.cos(yy / 13.0)
    relief[20:40, 10:50] = 200  # a plateau: zero gradient -> (127, 127, 255)
    relief[60:, :30] = 0
    out.append(("relief_96x80", rng.integers(0, 256, (96, 80, 2), dtype=np.uint8), relief.astype(np.uint8)[..., None]))
    return out
