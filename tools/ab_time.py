#!/usr/bin/env python3
"""A/B timing aid: K1 land sampling and the device-resident land / water bake for whichever library UPG_LIBRARY selects.
    UPG_LIBRARY=.../libupg_variant.so python tools/ab_time.py [--grid 500] [--texture 4096]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=500)
ap.add_argument("--texture", type=int, default=4096)
ap.add_argument("--flowers", type=int, default=0)
ap.add_argument("--poles", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
p = Planet(shape="sphere", elevation="lakes", coloring="default", grid=a.grid, poles=bool(a.poles), flowers=bool(a.flowers))
s = torch.cuda.current_stream()
buf = torch.empty(a.grid**3, dtype=torch.float32, device=dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(n):
        fn()
    e1.record(s)
    e1.synchronize()
    return e0.elapsed_time(e1) / n


k1 = timed(lambda: p.field_samples_device("land", 0, a.grid, buf.data_ptr(), s.cuda_stream))
out = {"lib": os.path.basename(os.environ.get("UPG_LIBRARY", "libupg.so")), "k1_land_ms": round(k1, 3)}
T = a.texture
for name, water in (("land", 0), ("water", 1)):
    m = p.field_mesh(name)
    uv = synth.octahedral_uv(m["positions"])
    idx = synth.drop_folded_triangles(uv, m["indices"])
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (m["positions"], -m["normals"], uv, idx.view(np.int32))]
    o = [torch.empty(T * T * c, dtype=torch.uint8, device=dev) for c in ((4 if water else 3), 2, 1)]
    ms = timed(lambda: p.bake_device(water, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[0].shape[0], d[3].data_ptr(), d[3].shape[0], T, T, 0, T, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), s.cuda_stream), 3)
    out["bake_%s_ms" % name] = round(ms, 3)
    out["mtexels_per_s_%s" % name] = round(p.last_covered_texels / ms / 1e3, 1)
    out["checksum_%s" % name] = int(o[0].to(torch.int64).sum().item())
    # position-weighted sums of the three images: equal outputs <=> equal numbers (for A/B builds that must stay bit-identical)
    out["digest_%s" % name] = [int((x.to(torch.int64) * (torch.arange(x.numel(), device=dev) % 65521 + 1)).sum().item()) for x in o]
print(out)
