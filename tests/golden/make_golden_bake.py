#!/usr/bin/env python3
"""Regenerates the colouring / baking fixtures in tests/golden/ from the REFERENCE ITSELF (oracle/_ref/libup_ref.so).

    python tests/golden/make_golden_bake.py

  tiles_<config>.npz     terrainTile outputs (13 columns, oracle/ref.py TILE_COLUMNS) for land / water / navigation purposes
                         at tests/bake_common.tile_probes, thread RNG pinned per tile position (upref_tiles_seeded)
  bake_<config>.npz      generateTexturesLand + generateTexturesWater outputs (u8) for bake_common.atlas_mesh at 192^2
  raster.npz             meshGenerateTexture coverage (owner, weights, visit count) for bake_common.random_triangles
  dilation.npz           imageDilation / imageConvert results incl. cage's own 11x11 test case (test-core/image.cpp:567-583)
  texture_tail.npz       imageConvertSpecialToGltfPbr + imageConvertHeigthToNormal(1) on bake_common.tail_images (python ... tail = only this)
  colors.npz             colorRgbToHsv / HsvToRgb / RgbToHsluv / HsluvToRgb / interpolateColor samples
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bake_common as bc  # noqa: E402
from oracle.ref import RefPlanet  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def texture_tail_fixture():
    ref = RefPlanet(init=False)
    d = {}
    for name, sp, hm in bc.tail_images():
        d[name + "_pbr"], d[name + "_normal"] = ref.texture_tail(sp, hm, 1.0)
    name, sp, hm = bc.tail_images()[-1]
    d["strength4_normal"] = ref.texture_tail(sp, hm, 4.0)[1]
    flat = d["relief_96x80_normal"][25:35, 15:45]
    assert (flat == (127, 127, 255)).all(), "a plateau maps to the +z normal"
    np.savez_compressed(os.path.join(OUT, "texture_tail.npz"), **d)
    print("texture tail ok")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "tail":
        return texture_tail_fixture()
    texture_tail_fixture()
    for cfg in bc.TILE_CONFIGS:
        s, e, c, p, f = cfg
        ref = RefPlanet(s, e, c, p, f)
        pos, nrm = bc.tile_probes(s)
        d = {"positions": pos, "normals": nrm}
        for purpose in bc.PURPOSES:
            d[purpose] = ref.tiles_seeded(purpose, pos, nrm)
        np.savez_compressed(os.path.join(OUT, "tiles_%s.npz" % bc.config_name(cfg)), **d)
        print("tiles", cfg, "biomes", np.bincount(d["land"][:, 11].astype(int), minlength=12))

    pos, nrm, uv, idx = bc.atlas_mesh()
    for cfg in bc.BAKE_CONFIGS:
        s, e, c, p, f = cfg
        ref = RefPlanet(s, e, c, p, f)
        d = {}
        for water in (0, 1):
            a, sp, hm = ref.bake_seeded(water, pos, nrm, uv, idx, 192, 192)
            tag = "water" if water else "land"
            d[tag + "_albedo"], d[tag + "_special"], d[tag + "_height"] = a, sp, hm
            d[tag + "_covered"] = np.array([ref.last_covered_texels], dtype=np.uint64)
            if c != "default":
                # without generateSand the thread RNG is never read: the re-assembled pipeline must equal the reference's own
                a2, sp2, hm2 = ref.bake(water, pos, nrm, uv, idx, 192, 192)
                assert (a2 == a).all() and (sp2 == sp).all() and (hm2 == hm).all(), "bake_seeded deviates from generateTextures*"
        np.savez_compressed(os.path.join(OUT, "bake_%s.npz" % bc.config_name(cfg)), **d)
        print("bake", cfg, "covered", int(d["land_covered"][0]))

    ref = RefPlanet(init=False)
    ruv, ridx = bc.random_triangles()
    rpos = np.zeros((len(ruv), 3), dtype=np.float32)
    d = {}
    for w, h in ((64, 64), (97, 41)):
        owner, weights, calls = ref.raster(rpos, ruv, ridx, w, h)
        d["owner_%dx%d" % (w, h)] = owner
        d["weights_%dx%d" % (w, h)] = weights
        d["calls_%dx%d" % (w, h)] = np.array([calls], dtype=np.uint64)
    auv = bc.atlas_mesh()[2]
    aidx = bc.atlas_mesh()[3]
    owner, weights, calls = ref.raster(np.zeros((len(auv), 3), np.float32), auv, aidx, 192, 192)
    d["atlas_owner"], d["atlas_weights"], d["atlas_calls"] = owner, weights, np.array([calls], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "raster.npz"), **d)
    print("raster", calls)

    d = {}
    for ch in (1, 2, 3, 4):
        img = bc.dilation_image(ch=ch, seed=5 + ch)
        for rounds in (1, 2, 7, 8):
            f, q = ref.dilate(img, rounds, use_nan=True, to_u8=True)
            d["nan_c%d_r%d_f" % (ch, rounds)], d["nan_c%d_r%d_u8" % (ch, rounds)] = f, q
    img = np.zeros((11, 11, 3), dtype=np.float32)  # cage's own test image (test-core/image.cpp:567-583), useNan = false
    img[0, 0] = (1, 0, 0)
    img[5, 5] = (0, 1, 0)
    img[0, 10] = (0, 0, 1)
    f, q = ref.dilate(img, 8, use_nan=False, to_u8=True)
    d["kat_input"], d["kat_f"], d["kat_u8"] = img, f, q
    assert (q[0, 0] == (255, 0, 0)).all() and (q[5, 5] == (0, 255, 0)).all() and (q[0, 10] == (0, 0, 255)).all()
    assert (q[10, 0] == (0, 255, 0)).all() and (q[10, 10] == (0, 255, 0)).all()  # get3(0,10), get3(10,10) are green
    small = bc.dilation_image(h=2, w=9, ch=3, seed=1, hole=0.5)  # the w < 3 || h < 3 path (imageAlgorithms.cpp:436-441)
    d["small_input"] = small
    d["small_f"], d["small_u8"] = ref.dilate(small, 7, use_nan=True, to_u8=True)
    np.savez_compressed(os.path.join(OUT, "dilation.npz"), **d)
    print("dilation ok")

    rng = np.random.default_rng(9)
    cols = rng.random((200, 3)).astype(np.float32)
    cols[:8] = [(0, 0, 0), (1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0.5, 0.5, 0.5), (1, 0.5, 0.5), (0.2, 0.2, 0.2000001)]
    d = {"input": cols}
    for what in ("rgb2hsv", "hsv2rgb", "rgb2hsluv", "hsluv2rgb"):
        d[what] = np.stack([ref.color(what, c) for c in cols])
    fs = rng.random(199).astype(np.float32)
    d["interp_f"] = fs
    d["interp"] = np.stack([ref.color_interpolate(cols[i], cols[i + 1], float(fs[i])) for i in range(199)])
    np.savez_compressed(os.path.join(OUT, "colors.npz"), **d)
    print("done")


if __name__ == "__main__":
    main()
