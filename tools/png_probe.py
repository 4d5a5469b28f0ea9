#!/usr/bin/env python3
"""PNG encoding of one chunk's three textures: upg_png_encode (device) next to zlib on the host (cage's exportFile = libpng + zlib at its
default level, single-threaded per image; libpng is an un-vendored submodule, so PIL's libpng stands in).   python tools/png_probe.py [--texture 4096]"""
import argparse
import io
import json
import os
import statistics
import sys
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet, png_encode, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--texture", type=int, default=4096)
ap.add_argument("--grid", type=int, default=500)
a = ap.parse_args()
T = a.texture
p = Planet(shape="sphere", elevation="lakes", grid=a.grid)
m = p.field_meshes_clean(("land",))["land"]
uv = synth.octahedral_uv(m["positions"])
idx = synth.drop_folded_triangles(uv, m["indices"])
images = p.bake_export(0, m["positions"], m["normals"], uv, idx, T, T)
out = {"texture": T}
from PIL import Image

for name, img in zip(("albedo", "pbr", "normal"), images):
    dev = torch.from_numpy(img).cuda()
    png_encode(dev)
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        data = png_encode(dev)
        ts.append((time.perf_counter() - t) * 1e3)
    hts = []
    for _ in range(3):
        t = time.perf_counter()
        data_h = png_encode(img)
        hts.append((time.perf_counter() - t) * 1e3)
    t = time.perf_counter()
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="PNG")  # libpng, zlib level 6, adaptive filtering: what cage's exportFile does
    pil_ms = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    z1 = zlib.compress(img.tobytes(), 1)
    z1_ms = (time.perf_counter() - t) * 1e3
    ok = bool(np.array_equal(np.asarray(Image.open(io.BytesIO(data))), img))
    out[name] = {"raw_mb": round(img.nbytes / 1e6, 1), "gpu_png_mb": round(len(data) / 1e6, 2), "gpu_ms_device_input": round(statistics.median(ts), 2), "gpu_ms_host_input": round(statistics.median(hts), 2),
                 "libpng_mb": round(buf.tell() / 1e6, 2), "libpng_ms": round(pil_ms, 1), "zlib1_mb": round(len(z1) / 1e6, 2), "zlib1_ms": round(z1_ms, 1), "decodes_equal": ok}
print(json.dumps(out))
