#!/usr/bin/env python3
"""The fused dense K1 launch of bench.py's timed region (three fields of one grid), for ncu: python tools/profile_dense.py [--grid 1000]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unnatural_planets_b200 import Planet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1000)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--arithmetic", default="exact")
a = ap.parse_args()
p = Planet(shape="sphere", elevation="lakes", grid=a.grid, arithmetic=a.arithmetic)
bufs = [torch.empty(a.grid**3, dtype=torch.float32, device="cuda") for _ in range(3)]
s = torch.cuda.current_stream()
for _ in range(a.repeat):
    p.fields_samples_device(0, a.grid, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.cuda_stream)
s.synchronize()
print("ok")
