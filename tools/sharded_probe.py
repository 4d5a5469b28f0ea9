#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/sharded_probe.py: phases of the sharded base meshes (UPG_TRACE) at 1024^3"""
import os
import sys
import time

os.environ.setdefault("UPG_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from unnatural_planets_b200 import sharded as sh

rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", 1)) > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sp = sh.ShardedPlanet(device=local, shape="sphere", elevation="lakes", grid=1024)
for rep in range(4):
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    m = sp.field_meshes(("land", "water", "navigation"), dst=0)
    if rank == 0:
        print("rep %d: %.2f ms %s" % (rep, (time.perf_counter() - t) * 1e3, sp.timing), file=sys.stderr, flush=True)
if dist.is_initialized():
    dist.destroy_process_group()
