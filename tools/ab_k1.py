#!/usr/bin/env python3
"""Dense K1 (fused land + water + navigation launch) in both arithmetic builds: time per launch and how far the contracted build
(UPG_ARITH_CONTRACT) is from the exact one.   python tools/ab_k1.py [--grid 1000] [--shape sphere] [--elevation lakes]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1000)
ap.add_argument("--shape", default="sphere")
ap.add_argument("--elevation", default="lakes")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
res = a.grid
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
out = {"grid": res, "shape": a.shape, "elevation": a.elevation, "lib": os.path.basename(os.environ.get("UPG_LIBRARY", "libupg.so"))}
fields = {}
for mode in ("exact", "contract"):
    p = Planet(shape=a.shape, elevation=a.elevation, grid=res, arithmetic=mode)
    bufs = [torch.empty(res**3, dtype=torch.float32, device=dev) for _ in range(3)]

    def step():
        p.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.cuda_stream)

    step()
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.steps):
        step()
    e1.record(s)
    e1.synchronize()
    out[mode + "_ms"] = round(e0.elapsed_time(e1) / a.steps, 3)
    fields[mode] = bufs
    p.async_status(s.cuda_stream)
    p.close()
for i, name in enumerate(("land", "water", "navigation")):
    x, c = fields["exact"][i], fields["contract"][i]
    err = (x.double() - c.double()).abs()
    tol = 1e-4 * torch.maximum(x.abs(), c.abs()).double()
    bad = err > tol
    near = (x.abs() < 1e-3) | (c.abs() < 1e-3)
    flips = ((x >= 0) != (c >= 0))
    worst = torch.topk(err.flatten(), 5)
    out[name] = {
        "bit_identical_fraction": float((x.view(torch.int32) == c.view(torch.int32)).double().mean()),
        "max_abs_err": float(err.max()), "max_rel_err_over_1e-4": float((err / tol.clamp_min(1e-30)).max()),
        "outside_1e-4_rel": int(bad.sum()), "sign_flips": int(flips.sum()), "sign_flips_outside_near_zero": int((flips & ~near).sum()),
        "worst_abs": [round(float(v), 6) for v in worst.values], "worst_exact_values": [round(float(x.flatten()[j]), 4) for j in worst.indices],
    }
print(json.dumps(out))
