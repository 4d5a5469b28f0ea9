#!/usr/bin/env python3
"""bench.py -- throughput of the generation hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid R] [--texture T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2], the largest single-GPU configuration: `--shape sphere` (elevation lakes, colouring
default, default seed) at 2x the meshing grid resolution (1000^3 samples per field, 3 fields: land, water, navigation;
sources/generator.cpp:432-439) with a 4096 x 4096 albedo / special / height bake of the land and the water surface.

The metric is two-part (BASELINE.json: "SDF samples/s and baked texels/s ... + planet gen time vs host CPU"):
  * `value` (metric sdf_samples_per_s): one step = the dense SDF sampling of one planet's three fields, samples resident in
    HBM, timed with CUDA events over exactly K steps after W warm-up steps, max over ranks;  `e2e` = the same samples
    through the host-buffer C-ABI call upg_field_samples into pinned host memory (device->host copies inside the timing).
  * `texel_bake` (baked_texels_per_s): one step = upg_bake_device of the land + water textures (K5 raster, K6 colouring,
    K7 dilation + quantisation) with mesh and outputs resident in HBM;  its `e2e` = upg_bake with host mesh and host images.
  * `planet`: wall-clock of the whole hot path of one planet through the C ABI with host buffers (3 x upg_field_mesh =
    sampling + isosurface + mesh download, 2 x upg_bake), next to the CPU reference's extrapolated time for the same work.
Weak scaling: every rank generates its own planet (planet i uses the default seed stream advanced by i draws); there is no
data-path collective; whole-job value = units of all ranks / max-over-ranks device time.
`roofline`: the dominant kernel of each leg against the FP32 FMA ceiling measured in this run (both legs are FP32 / INT32
issue bound, nothing on this path is a contraction; DESIGN.md "Rooflines").  `cpu_baseline`: the reference's own loops
(oracle/_ref = the reference sources compiled verbatim) on a bounded sample, all host threads, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_STDOUT = sys.stdout
METRIC = "sdf_samples_per_s"
UNIT = "samples/s"
FIELDS = ("land", "water", "navigation")
SHAPE, ELEVATION, COLORING = "sphere", "lakes", "default"
# Algorithmic FP32 work per unit (FMA = 2): EXACT COUNTS from the instrumented build of the CPU restatement (oracle/up_count.cpp,
# tests/hostlib/upg_host_count.cpp; tools/count_ops.py wrote profiles/r02_work_counts_config3.json, tests/test_oracle_counts.py pins
# it).  Per SDF sample as the reference evaluates it (each field on its own: sources/generator.cpp:432-439), per baked land texel.
def _work_counts():
    c = json.load(open(os.path.join(ROOT, "profiles", "r02_work_counts_config3.json")))
    f = c["fields"]
    return {
        "per_grid_node": c["flops_per_grid_node"],  # land + water + navigation, three evaluations like the reference's three tasks
        "per_sample": c["flops_per_grid_node"] / 3.0,
        "one_evaluation": f["land"]["flops"] + 3.0,  # ONE shape + ONE elevation evaluation feeding the three combinations (what the fused launch needs)
        "special_per_grid_node": c["special_per_grid_node"],
        "land_texel": c["texels"]["land"]["flops"], "land_texel_fp64": c["texels"]["land"]["flops_fp64"],
        "water_texel": c["texels"]["water"]["flops"],
    }


WORK = _work_counts()


def _ncu(name):
    """per-unit figures of a committed ncu capture (profiles/<name>.json, tools/ncu_ops_json.py); None when absent"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


K1_NCU = _ncu("r02_k1_ops_exact.json")
K6_NCU = _ncu("r02_k6_ops.json")


def workload(args):
    return "BASELINE configs[2]: --shape sphere (elevation lakes, colouring default, default seed), 3 fields (land, water, navigation) x %d^3 grid, %d x %d land + water bake" % (args.grid, args.texture, args.texture)


def bench_config(args):
    """the SAME object in both arms (ours / --impl reference): what is generated and what one unit of the metric is"""
    n = args.grid**3
    return {
        "workload": workload(args), "grid": args.grid, "fields": list(FIELDS), "texture": args.texture, "planets_per_step": args.gpus,
        "sample": "one terrainSdf* value; every grid node yields three (land, water, navigation) and BOTH arms evaluate all three fields of the nodes they time",
        "gpu_arithmetic": "%s (UPG_ARITH_*: the GPU arm's dense sampling build; the CPU arm is the reference as compiled)" % args.arithmetic,
        "l2": "a sampling step writes %.1f GB per GPU and a bake step touches > 1 GB of float images (>> 126 MB L2); inputs are a ~2 KB parameter block and a mesh" % (12e-9 * n),
    }


def seed_for(index):
    """seed stream of planet `index`: the default xoroshiro128** state advanced by `index` draws"""
    s = [1346519564496, 42245614964156]
    mask = (1 << 64) - 1

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & mask

    for _ in range(index):
        s0, s1 = s
        s1 ^= s0
        s = [rotl(s0, 24) ^ s1 ^ ((s1 << 16) & mask), rotl(s1, 37)]
    return tuple(s)


class ClockSampler:
    """polls SM clock / throttle reasons during the timed regions (B200_PROFILING.md 'clocks line')"""

    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.max_mhz = None
        self.reasons = set()
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _poll(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            self.samples.append(float(f[0]))
            self.max_mhz = float(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._poll()
            self._stop.wait(0.05)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU reference legs (oracle/_ref; oracle/ is only ever the checker or the timed baseline, never the product) ----
def cpu_sampling_rate(res, slices, threads):
    """the reference's sampling loops (MarchingCubes::updateByPosition x terrainSdfLand / Water / Navigation: the three tasks of
    sources/generator.cpp:432-439) on `slices` z-slices around the equator, every loop split over `threads` threads;
    returns (samples/s counting all three fields, seconds, kind)"""
    from oracle import ref as oref

    z0 = res // 2 - slices // 2
    if not oref.available():
        from oracle import port as oport  # the standalone restatement

        p = oport.OraclePlanet(SHAPE, ELEVATION)
        t = time.perf_counter()
        for f in FIELDS:
            p.grid(f, res, z0=z0, z1=z0 + slices, threads=threads)
        dt = time.perf_counter() - t
        return len(FIELDS) * slices * res * res / dt, dt, "port"
    planet = oref.RefPlanet(shape=SHAPE, elevation=ELEVATION)
    t = time.perf_counter()
    for f in FIELDS:
        planet.grid(f, res, z0=z0, z1=z0 + slices, threads=threads)
    dt = time.perf_counter() - t
    return len(FIELDS) * slices * res * res / dt, dt, "reference"


def cpu_bake_rate(mesh, texture, window, threads):
    """generateTexturesLand (rasterise + terrainTile per visit + 7 dilations + u8) of `threads` UV windows of the full bake,
    one window per host thread -- the reference bakes one chunk per pool thread (sources/generator.cpp:159)"""
    from oracle import ref as oref
    from unnatural_planets_b200 import synth

    if not oref.available():
        return None
    planet = oref.RefPlanet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING)
    pos, nrm, uv, idx = mesh
    size = window / texture
    jobs = []
    for k in range(threads):
        # windows spread over the upper-hemisphere part of the octahedral map (no folds there)
        u0 = 0.27 + (k % 4) * size * 1.01
        v0 = 0.27 + (k // 4) * size * 1.01
        jobs.append(synth.uv_window_submesh(pos, nrm, uv, idx, u0, v0, size))
    covered = [0] * threads

    def work(k):
        p, n, u, i = jobs[k]
        a, s, h = planet.bake(0, p, n, u, i, window, window)
        covered[k] = int((a != 0).any(axis=2).sum())

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t = time.perf_counter()
    [x.start() for x in ts]
    [x.join() for x in ts]
    dt = time.perf_counter() - t
    return sum(covered) / dt, dt, sum(covered)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    res = args.grid
    slices = max(1, (6 * threads * 250000) // (res * res))  # ~1.5e6 nodes x 3 fields per thread and step: seconds of CPU work
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, kind = cpu_sampling_rate(res, slices, threads)
        if i >= args.warmup:
            rates.append((r, dt))
    value = statistics.mean(r for r, _ in rates)
    ms = statistics.mean(dt for _, dt in rates) * 1e3
    sample = "land + water + navigation of %d z-slices x %d^2 per step (%.3g of one planet's %.3g samples), %d threads" % (slices, res, slices / float(res), 3.0 * res**3, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_STDOUT, flush=True)


def bench_mesh(raw):
    """the land and water surfaces of the planet with synthetic texture coordinates (see unnatural_planets_b200/synth.py)"""
    import numpy as np

    from unnatural_planets_b200 import synth

    out = {}
    for name in ("land", "water"):
        m = raw[name]
        uv = synth.octahedral_uv(m["positions"])
        idx = synth.drop_folded_triangles(uv, m["indices"])
        # meshFlipNormals (sources/meshGeneration.cpp:42): the generator bakes the flipped mesh
        out[name] = (m["positions"], np.ascontiguousarray(-m["normals"]), uv, idx)
    return out


# ---- the one-planet-over-N-GPUs legs (BASELINE configs[4]) run ISOLATED: each in its own process (group) started by rank 0 after the
# headline legs are done and the other ranks have left, under a timeout.  A failure or a hang there costs its own scalars, never the line.
def leg_sharded(args, rank, world, local_rank):
    """ONE planet sharded over all ranks, one process per GPU over NCCL (unnatural_planets_b200/sharded.py): z-slab sampling + extraction,
    row-tile bake.  Returns the leg's record on rank 0, None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from unnatural_planets_b200 import sharded as sh

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sp = sh.ShardedPlanet(device=local_rank, shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=args.sharded_grid)
    stex = args.sharded_texture
    sh.gather_varlen(np.zeros(4, dtype=np.float32), dst=0)  # opens NCCL's point-to-point channels outside the timing
    barrier()
    phases = {}
    for _ in range(2):  # first calls: NCCL channels of this message size, memory pools, pinned staging
        sp.field_meshes(FIELDS, dst=0)
    mesh_times = []
    for _ in range(3):  # median of three calls, each between barriers (max over ranks below)
        barrier()
        t = time.perf_counter()
        smeshes = sp.field_meshes(FIELDS, dst=0)
        mesh_times.append((time.perf_counter() - t, sp.timing))
        barrier()
    mesh_times.sort(key=lambda x: x[0])
    t_mesh, phases["meshes"] = mesh_times[1]
    # the bake input lives on rank 0 after the gather: broadcast it (a few tens of MB) like the generator would hand every
    # device the chunk mesh; the synthetic UVs are host work outside the timing
    src = None
    if rank == 0:
        from unnatural_planets_b200 import synth

        src = {}
        for name in ("land", "water"):
            m = smeshes[name]
            uv = synth.octahedral_uv(m["positions"])
            src[name] = [m["positions"], np.ascontiguousarray(-m["normals"]), uv, synth.drop_folded_triangles(uv, m["indices"])]
    # only rank 0 holds the mesh (like the generator's host process would): it uploads once and the mesh reaches the other GPUs as
    # one NCCL broadcast inside the timed bake
    none4 = [None, None, None, None]
    for name, water in (("land", 0), ("water", 1)):  # first call: NCCL channels of this message size, memory pools
        sp.bake(water, *(src[name] if rank == 0 else none4), stex, stex, dst=0, src=0)
    bake_times = []
    for _ in range(3):  # median of three, like the meshes
        barrier()
        t = time.perf_counter()
        stexels = 0
        ph = {}
        for name, water in (("land", 0), ("water", 1)):
            sp.bake(water, *(src[name] if rank == 0 else none4), stex, stex, dst=0, src=0)
            ph["bake_" + name] = sp.timing
            stexels += sp.last_covered_texels
        barrier()
        bake_times.append((time.perf_counter() - t, ph))
    bake_times.sort(key=lambda x: x[0])
    t_bake = bake_times[1][0]
    phases.update(bake_times[1][1])
    st = torch.tensor([t_mesh, t_bake], dtype=torch.float64, device=dev)
    sc = torch.tensor([float(stexels)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        dist.all_reduce(sc, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    return {
        "what": "ONE planet over %d GPU(s), strong scaling: upg_field_meshes_slab_device (own z-slab + 3-slice recomputed halo sampled once for the 3 fields, 3 extractions, chunks stay in HBM) + count all_gather + NCCL gather of the mesh chunks to rank 0 + one download there; then, per surface, the mesh uploaded by rank 0 and broadcast over NCCL, upg_bake_device on own texel rows (+ 7-row apron) + NCCL gather of the row tiles; wall clock, max over ranks, median of three calls after two warm-up calls" % world,
        "gpus": world, "grid": args.sharded_grid, "texture": stex, "isosurface_ms": float(st[0]) * 1e3, "bake_ms": float(st[1]) * 1e3,
        "samples_per_s": 3.0 * args.sharded_grid**3 / float(st[0]), "texels_per_s": float(sc[0]) / float(st[1]),
        "land_triangles": int(len(smeshes["land"]["indices"]) // 3), "covered_texels": int(sc[0]), "rank0_phases_ms": phases,
    }


def leg_multi_device(args, devices):
    """the same planet over `devices` GPUs from ONE process through the C ABI (upg_planet_desc::device_count, csrc/upg_multi.cu)"""
    import ctypes as C

    import torch

    from unnatural_planets_b200 import Planet, capi
    from unnatural_planets_b200 import synth as _synth

    L = capi.load_library()
    n = len(devices)
    mres, mtex = args.sharded_grid, args.sharded_texture
    mp = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=mres, devices=list(devices))
    for _ in range(2):
        mm = mp.field_meshes_clean(FIELDS)
    mt = []
    for _ in range(3):
        mm = mp.field_meshes_clean(FIELDS)
        mt.append(mp.last_call_s)
    msrc = {}
    for name in ("land", "water"):
        uv = _synth.octahedral_uv(mm[name]["positions"])
        msrc[name] = (mm[name]["positions"], mm[name]["normals"], uv, _synth.drop_folded_triangles(uv, mm[name]["indices"]))
    mimg = {"land": [torch.empty(mtex * mtex * c, dtype=torch.uint8, pin_memory=True) for c in (3, 3, 3)], "water": [torch.empty(mtex * mtex * c, dtype=torch.uint8, pin_memory=True) for c in (4, 3, 3)]}
    mtexels = [0]
    failures = []

    def mbake(name, water):
        try:
            pos, nrm, uv, idx = msrc[name]
            o = mimg[name]
            capi.check(L.upg_bake_export(mp._h, water, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), mtex, mtex, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))
            mtexels[0] += int(L.upg_bake_last_covered_texels())
        except Exception as ex:
            failures.append(ex)

    def mbake_both():
        mtexels[0] = 0
        ts = [threading.Thread(target=mbake, args=x) for x in (("land", 0), ("water", 1))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if failures:
            raise failures[0]

    mbake_both()
    bt = []
    for _ in range(3):
        t = time.perf_counter()
        mbake_both()
        bt.append(time.perf_counter() - t)
    m_ms, b_ms = statistics.median(mt) * 1e3, statistics.median(bt) * 1e3
    line = {
        "what": "ONE planet over %d GPU(s) from ONE process through the C ABI (upg_planet_desc::device_count): upg_field_meshes_clean (z-slab per device, chunks gathered over NVLink per field, K9 + download on three devices) + 2 x upg_bake_export from two host threads (mesh parts uploaded per device and exchanged over NVLink, texel rows per device written straight into the pinned host images); host buffers, wall clock, median of three calls" % n,
        "multi_device_planet_devices": n, "multi_device_planet_grid": mres, "multi_device_planet_texture": mtex,
        "multi_device_planet_base_meshes_ms": m_ms, "multi_device_planet_bake_ms": b_ms, "multi_device_planet_ms": m_ms + b_ms,
        "multi_device_planet_texels_per_s": mtexels[0] / (b_ms * 1e-3), "multi_device_planet_samples_per_s": 3.0 * mres**3 / (m_ms * 1e-3),
        "multi_device_planet_land_triangles": int(len(mm["land"]["indices"]) // 3),
    }
    mp.close()
    return line


def leg_export(args):
    """one chunk's export: the land bake of a 500^3 planet at the bench's texture size as three PNG FILES (albedo, pbr, normal map) --
    what chunkEntry does with three Image::exportFile calls (sources/generator.cpp:135-139; libpng + zlib on one host thread per image).
    Every part reports on its own: a failure of one costs its scalars only."""
    import io

    import numpy as np
    import torch

    from unnatural_planets_b200 import Planet, png_encode, synth

    T = args.texture
    p = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=500)
    m = p.field_meshes_clean(("land",))["land"]
    uv = synth.octahedral_uv(m["positions"])
    idx = synth.drop_folded_triangles(uv, m["indices"])
    images = p.bake_export(0, m["positions"], m["normals"], uv, idx, T, T)
    t = time.perf_counter()
    images = p.bake_export(0, m["positions"], m["normals"], uv, idx, T, T)
    line = {"chunk_export_texture": T, "chunk_export_bake_ms": (time.perf_counter() - t) * 1e3, "chunk_export_raw_mb": sum(i.nbytes for i in images) / 1e6}
    try:  # K10 on images that are already on the device (tests/test_gpu_png.py checks the files byte for byte through two decoders)
        dev = [torch.from_numpy(i).cuda() for i in images]
        files = [png_encode(d) for d in dev]
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            files = [png_encode(d) for d in dev]
            ts.append((time.perf_counter() - t) * 1e3)
        line.update({"chunk_export_png_encode_ms": statistics.median(ts), "chunk_export_png_mb": sum(len(f) for f in files) / 1e6})
        from PIL import Image

        line["chunk_export_png_decodes_equal"] = bool(all(np.array_equal(np.asarray(Image.open(io.BytesIO(f))), i) for f, i in zip(files, images)))
        del dev
    except Exception as ex:
        line["chunk_export_png_error"] = repr(ex)[:200]
    try:  # bake + tail + PNG in one call, host mesh in, three files out (first run on hardware is the driver's: DESIGN.md section 9)
        out = [torch.empty(int(capi_bound(T, T, c)), dtype=torch.uint8, pin_memory=True).numpy() for c in (3, 3, 3)]
        p.bake_export_png(0, m["positions"], m["normals"], uv, idx, T, T, out=out)
        ts = []
        for _ in range(3):
            files1 = p.bake_export_png(0, m["positions"], m["normals"], uv, idx, T, T, out=out)
            ts.append(p.last_call_s * 1e3)
        from PIL import Image

        line.update({"chunk_export_one_call_ms": statistics.median(ts), "chunk_export_one_call_decodes_equal": bool(all(np.array_equal(np.asarray(Image.open(io.BytesIO(f.tobytes()))), i) for f, i in zip(files1, images)))})
    except Exception as ex:
        line["chunk_export_one_call_error"] = repr(ex)[:200]
    try:  # the host side of the comparison: libpng (PIL: zlib level 6, adaptive filters -- cage's exportFile settings) on ONE image, one thread
        from PIL import Image

        t = time.perf_counter()
        buf = io.BytesIO()
        Image.fromarray(images[0]).save(buf, format="PNG")
        line.update({"chunk_export_libpng_albedo_ms": (time.perf_counter() - t) * 1e3, "chunk_export_libpng_albedo_mb": buf.tell() / 1e6})
    except Exception as ex:
        line["chunk_export_libpng_error"] = repr(ex)[:200]
    p.close()
    return line


def tensor_checksums(bufs, step=1 << 24):
    """two 64-bit checksums (plain and position-weighted sum of the bit patterns, modulo 2^64) of float32 tensors, computed where they live"""
    import torch

    a = b = 0
    mask = (1 << 64) - 1
    w = None
    for buf in bufs:
        v = buf.view(torch.int32)
        if w is None:
            w = torch.arange(1, step + 1, dtype=torch.int64, device=v.device)
        for at in range(0, v.numel(), step):
            c = v[at:at + step].to(torch.int64)
            a = (a + int(c.sum())) & mask
            b = (b + int((c * w[: c.numel()]).sum())) & mask
    return "%016x%016x" % (a, b)


def leg_experiments(args):
    """dense K1 launches of whichever build of the library UPG_LIBRARY names (the parent runs this leg once per build and compares):
    config 3 (sphere + lakes, column kernel) and config 2 (torus + craters at 500^3, sample kernel), exact and tolerance arithmetic --
    time per fused launch (CUDA events) and two 64-bit checksums of the three sample fields"""
    import torch

    from unnatural_planets_b200 import Planet, library_path

    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    line = {"library": os.path.basename(library_path())}

    for tag, shape, elevation, res in (("lakes", "sphere", "lakes", args.grid), ("craters", "torus", "craters", 500)):
        bufs = [torch.empty(res**3, dtype=torch.float32, device=dev) for _ in FIELDS]
        for mode in ("exact", "contract"):
            p = Planet(shape=shape, elevation=elevation, grid=res, arithmetic=mode)

            def step():
                p.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream)

            step()
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                step()
            e1.record(stream)
            e1.synchronize()
            p.async_status(stream.cuda_stream)
            line["%s_%s_ms" % (tag, mode)] = e0.elapsed_time(e1) / 5
            line["%s_%s_checksum" % (tag, mode)] = tensor_checksums(bufs)
            p.close()
        del bufs
        torch.cuda.empty_cache()
    # the land bake of the bench planet (K5 raster + K6 colouring + K7 dilation / quantisation, everything resident in HBM): time per call and the
    # albedo written to UPG_EXPERIMENT_DUMP for the parent to compare between the builds (the experiments build evaluates HSLuv in single precision)
    import numpy as np

    from unnatural_planets_b200 import synth

    tex = args.texture
    p = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=500)
    m = p.field_meshes_clean(("land",))["land"]
    uv = synth.octahedral_uv(m["positions"])
    idx = synth.drop_folded_triangles(uv, m["indices"])
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (m["positions"], -m["normals"], uv, idx.view(np.int32))]
    o = [torch.empty(tex * tex * c, dtype=torch.uint8, device=dev) for c in (3, 2, 1)]

    def bake():
        p.bake_device(0, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[0].shape[0], d[3].data_ptr(), d[3].shape[0], tex, tex, 0, tex, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream.cuda_stream)

    bake()
    bake()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        bake()
    e1.record(stream)
    e1.synchronize()
    line["land_bake_ms"] = e0.elapsed_time(e1) / 5
    line["land_bake_covered_texels"] = int(p.last_covered_texels)
    dump = os.environ.get("UPG_EXPERIMENT_DUMP")
    if dump:
        np.save(dump, o[0].cpu().numpy())
    p.close()
    return line


def compare_albedo_dumps(a_path, b_path):
    """share of texels whose albedo differs between two dumps of leg_experiments, and by how much"""
    import numpy as np

    a, b = np.load(a_path).reshape(-1, 3).astype(np.int16), np.load(b_path).reshape(-1, 3).astype(np.int16)
    step = np.abs(a - b).max(axis=1)
    covered = (a != 0).any(axis=1) | (b != 0).any(axis=1)
    n = max(1, int(covered.sum()))
    return {"texels": int(covered.sum()), "differing_share": float((step > 0).sum() / n), "more_than_one_step_share": float((step > 1).sum() / n), "max_step": int(step.max())}


def leg_chunking(args):
    """upg_mesh_chunking (cage's meshChunking as meshSplit calls it, maxSurfaceArea 800000) on the cleaned land mesh of a 500^3 planet, next to
    the reference's own meshChunking compiled from the reference tree (oracle/_ref: the timed CPU baseline and the checker of this leg)"""
    import numpy as np

    from unnatural_planets_b200 import Planet, mesh_chunking

    m = Planet(shape=SHAPE, elevation=ELEVATION, grid=500).field_meshes_clean(("land",))["land"]
    got = mesh_chunking(m["positions"], m["normals"], m["indices"], max_surface_area=800000.0)
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        got = mesh_chunking(m["positions"], m["normals"], m["indices"], max_surface_area=800000.0)
        ts.append((time.perf_counter() - t) * 1e3)
    line = {"chunking_triangles": int(len(m["indices"]) // 3), "chunking_chunks": len(got), "chunking_ms": statistics.median(ts)}
    try:
        from oracle import ref as oref

        if oref.available():
            r = oref.RefPlanet(init=False)
            t = time.perf_counter()
            want = r.chunking(m["positions"], m["normals"], m["indices"], max_surface_area=800000.0)
            line["chunking_reference_cpu_ms"] = (time.perf_counter() - t) * 1e3
            line["chunking_equals_reference"] = bool(len(got) == len(want) and all(a[k].shape == b[k].shape and np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for a, b in zip(got, want) for k in ("positions", "normals", "indices")))
    except Exception as ex:
        line["chunking_reference_error"] = repr(ex)[:200]
    return line


def capi_bound(w, h, ch):
    from unnatural_planets_b200 import capi

    return capi.load_library().upg_png_bound(w, h, ch)


def run_leg(args, rank, world, local_rank):
    """`--leg sharded|multi`: one isolated leg, one JSON line (rank 0)"""
    if args.leg == "multi":
        line = leg_multi_device(args, list(range(args.gpus)))
    elif args.leg == "export":
        line = leg_export(args)
    elif args.leg == "experiments":
        line = leg_experiments(args)
    elif args.leg == "chunking":
        line = leg_chunking(args)
    else:
        line = leg_sharded(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(line), file=_STDOUT, flush=True)


_TORCHRUN_ENV = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK", "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS", "NCCL_ASYNC_ERROR_HANDLING", "TORCH_NCCL_ASYNC_ERROR_HANDLING")


def run_selftest_leg(args, rank, world):
    """`--leg selftest` / `--leg hang`: the orchestration of the isolated legs checked on CPU (gloo), tests/test_bench_orchestration.py"""
    if args.leg == "hang":
        time.sleep(3600)
    total = float(rank + 1)
    if world > 1:
        import torch
        import torch.distributed as dist

        dist.init_process_group("gloo")
        t = torch.tensor([total], dtype=torch.float64)
        dist.all_reduce(t)
        total = float(t[0])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"selftest_sum": total, "world": world, "inherited_rank_env": [k for k in _TORCHRUN_ENV if k in os.environ and world == 1]}), file=_STDOUT, flush=True)


def run_orchestration_selftest(args, rank, world):
    """the control flow of run_ours without a GPU: collective, group torn down on every rank, the other ranks leave, rank 0 runs the
    isolated legs (one nested launch over `world` processes, one that hangs and is killed) and prints the line"""
    if world > 1:
        import torch
        import torch.distributed as dist

        dist.init_process_group("gloo")
        t = torch.tensor([1.0], dtype=torch.float64)
        dist.all_reduce(t)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ok = run_isolated_leg("selftest", args, world, 120)
    single = run_isolated_leg("selftest", args, 1, 120)
    hung = run_isolated_leg("hang", args, world, 5)
    print(json.dumps({"nested": ok, "single": single, "hung": hung}), file=_STDOUT, flush=True)


def isolated_leg_command(leg, args, world, port):
    """argv of the process (group) that runs one isolated leg over `world` GPUs"""
    tail = [os.path.join(ROOT, "bench.py"), "--leg", leg, "--gpus", str(world), "--sharded-grid", str(args.sharded_grid), "--sharded-texture", str(args.sharded_texture)]
    if leg in ("export", "experiments", "chunking"):  # one process on one GPU whatever the size of the job
        return [sys.executable, os.path.join(ROOT, "bench.py"), "--leg", leg, "--gpus", "1", "--grid", str(args.grid), "--texture", str(args.texture)]
    if leg != "multi" and world > 1:
        return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", str(port)] + tail
    return [sys.executable] + tail


def free_port():
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def kill_process_tree(proc):
    """SIGKILL for the process this module started and every descendant of it (torchrun puts each worker into a session of its own,
    so the process group of the launcher is not enough); by pid, nothing is matched by name"""
    import signal

    victims = []
    try:
        import psutil

        victims = [c.pid for c in psutil.Process(proc.pid).children(recursive=True)]
    except Exception:
        pass
    for pid in [proc.pid] + victims:
        try:
            os.kill(pid, signal.SIGKILL)
        except Exception:
            pass
    try:
        os.killpg(proc.pid, signal.SIGKILL)  # the session this call started
    except Exception:
        pass


_LEGS_DEADLINE = [None]  # wall-clock instant after which no isolated leg is started or kept running (args.legs_budget after the first one)


def run_isolated_leg(leg, args, world, timeout_s, extra_env=None):
    """runs `bench.py --leg <leg>` in its own session under a timeout; returns its record, or {"error": "..."}"""
    if _LEGS_DEADLINE[0] is None:
        _LEGS_DEADLINE[0] = time.perf_counter() + getattr(args, "legs_budget", 420)
    left = _LEGS_DEADLINE[0] - time.perf_counter()
    if left < 15:
        return {"error": "skipped: the time budget of the isolated legs is spent"}
    timeout_s = min(timeout_s, left)
    env = {k: v for k, v in os.environ.items() if k not in _TORCHRUN_ENV and not k.startswith("TORCHELASTIC")}
    env.update(extra_env or {})
    cmd = isolated_leg_command(leg, args, world, free_port())
    t = time.perf_counter()
    try:
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=None, env=env, cwd=ROOT, start_new_session=True, text=True)
    except Exception as ex:
        return {"error": "could not start: %r" % (ex,)}
    try:
        out, _ = proc.communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        kill_process_tree(proc)
        try:
            proc.communicate(timeout=10)
        except Exception:
            pass
        return {"error": "timeout after %d s" % int(timeout_s)}
    record = None
    for text in reversed((out or "").strip().splitlines()):
        try:
            record = json.loads(text)
            break
        except Exception:
            continue
    if proc.returncode != 0 or not isinstance(record, dict):
        return {"error": "exit code %s, no record" % proc.returncode}
    record["leg_wall_s"] = time.perf_counter() - t
    return record


def wait_for_other_processes(timeout_s=20.0):
    """after the process group is gone the other ranks exit: wait until no other compute process is left on the node's GPUs (bounded)"""
    t = time.perf_counter()
    while time.perf_counter() - t < timeout_s:
        try:
            out = subprocess.run(["nvidia-smi", "--query-compute-apps=pid", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout
            # (a container sees host pids here, so the count is what can be compared: at most this process is left)
            if len([x for x in out.split() if x.strip().isdigit()]) <= 1:
                return True
        except Exception:
            time.sleep(3.0)
            return False
        time.sleep(0.5)
    return False


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from unnatural_planets_b200 import Planet, capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    res, tex = args.grid, args.texture
    # the dense sampling of the headline runs in the TOLERANCE build by default (UPG_ARITH_CONTRACT: FMA contraction + SFU pow / reciprocal
    # inside the elevation curves; within the 1e-4 relative tolerance north_star states, integer lattice bit-exact; tests/test_gpu_fullsize.py
    # checks it against the reference at the BASELINE sizes); --arithmetic exact times the bit-reproducible build instead.  The mesh
    # entry points and the bake always run exact.
    ARITH, OTHER = args.arithmetic, ("exact" if args.arithmetic == "contract" else "contract")
    planet = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=res, device=local_rank, seed=seed_for(rank), arithmetic=ARITH)
    n = res**3
    bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in FIELDS]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: SDF sampling (the contract's timed region) ---------------------------------------------------------
    def step():
        # the three fields of one planet from ONE fused K1 launch: shape + elevation evaluated once per grid node, three stores
        planet.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)

    # the same kernel restricted to one field each (what three separate upg_field_samples_device calls cost), CUDA events on
    # the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    single_ms = {}
    for f, b in zip(FIELDS, bufs):
        kernel_ms = []
        for _ in range(max(3, min(args.steps, 5))):
            k0.record(stream)
            planet.field_samples_device(f, 0, res, b.data_ptr(), stream.cuda_stream)
            k1.record(stream)
            k1.synchronize()
            kernel_ms.append(k0.elapsed_time(k1))
        single_ms[f] = statistics.mean(kernel_ms)

    # the same fused launch in the OTHER arithmetic build
    cplanet = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=res, device=local_rank, seed=seed_for(rank), arithmetic=OTHER)
    cplanet.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream)
    k0.record(stream)
    csteps = max(3, min(args.steps, 10))
    for _ in range(csteps):
        cplanet.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream)
    k1.record(stream)
    k1.synchronize()
    contract_ms = k0.elapsed_time(k1) / csteps
    cplanet.async_status(stream.cuda_stream)
    cplanet.close()
    planet.async_status(stream.cuda_stream)

    # end to end through the host-buffer C-ABI call: device->host copies inside the timed region
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host_np = host.numpy().reshape(res, res, res)
    e2e_steps = max(1, min(args.steps, 2))

    def e2e_step():
        for f in FIELDS:
            planet.field_samples(f, 0, res, host_np)

    e2e_step()
    barrier()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t) / e2e_steps
    del host, host_np
    del bufs
    torch.cuda.empty_cache()

    # ---- leg 2: isosurface (sampling + K2..K4 + download, through the C ABI) -------------------------------------------
    for f, m in planet.field_meshes_clean(FIELDS).items():  # warm the memory pools like a second planet of a batch would find them
        del m
    barrier()
    planet.field_meshes(FIELDS)  # warms the (larger) pinned blocks of the raw meshes
    raw = planet.field_meshes(FIELDS)
    mesh_raw_s = planet.last_call_s
    # the complete meshGenerateBaseLand / Water / Navigation bodies: + K9 clean-up (non-manifold filter, box clip, largest component,
    # flip, water culling) before the meshes leave HBM; the C call alone is timed (host meshes in pinned memory when it returns)
    final = planet.field_meshes_clean(FIELDS)
    mesh_s = planet.last_call_s
    final_stats = {f: {"vertices": int(len(m["positions"])), "triangles": int(len(m["indices"]) // 3)} for f, m in final.items()}
    del final
    t = time.perf_counter()
    m2 = planet.field_mesh("land")
    field_mesh_land_s = time.perf_counter() - t
    mesh_stats = {"land_vertices": int(len(m2["positions"])), "land_triangles": int(len(m2["indices"]) // 3), "navigation_triangles": int(len(raw["navigation"]["indices"]) // 3), "near_zero_samples": m2["near_zero_samples"]}
    meshes = bench_mesh(raw)  # synthetic UVs (host numpy, outside every timing); the input of the bake
    del raw, m2
    # how the grid got filled for those meshes: band sampling (K1b: water field densely, elevation only where the isosurface can see it)
    # against the dense fused launch of leg 1 -- device timed, plus the same mesh call on a planet forced to dense sampling
    bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in FIELDS]
    evaluated = planet.fields_samples_band_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream, count=True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    band_steps = max(3, min(args.steps, 10))
    s0.record(stream)
    for _ in range(band_steps):
        planet.fields_samples_band_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), stream.cuda_stream)
    s1.record(stream)
    s1.synchronize()
    band_ms = s0.elapsed_time(s1) / band_steps
    del bufs
    torch.cuda.empty_cache()
    dense_planet = Planet(shape=SHAPE, elevation=ELEVATION, coloring=COLORING, grid=res, device=local_rank, seed=seed_for(rank), mesh_sampling="dense")
    dense_planet.field_meshes_clean(FIELDS)
    dense_planet.field_meshes_clean(FIELDS)
    mesh_dense_s = dense_planet.last_call_s
    dense_planet.close()
    sampling_stats = {"mode": "band", "band_sampling_ms": band_ms, "evaluated_nodes": int(evaluated), "evaluated_fraction": evaluated / n, "base_meshes_ms_with_dense_sampling": mesh_dense_s * 1e3,
                      "what": "upg_fields_samples_band_device over the whole grid (3 fields): dense water launch + band_sample_kernel, CUDA events; the extracted meshes are bit-identical to dense sampling (tests/test_gpu_mesh.py)"}

    # ---- leg 3: texel baking -------------------------------------------------------------------------------------------
    dmesh = {}
    for name, (pos, nrm, uv, idx) in meshes.items():
        dmesh[name] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pos, nrm, uv, idx.view(np.int32)))
    outs = {"land": [torch.empty(tex * tex * c, dtype=torch.uint8, device=dev) for c in (3, 2, 1)], "water": [torch.empty(tex * tex * c, dtype=torch.uint8, device=dev) for c in (4, 2, 1)]}
    covered = {}

    def bake_step():
        for name, water in (("land", 0), ("water", 1)):
            p, nm, u, i = dmesh[name]
            o = outs[name]
            planet.bake_device(water, p.data_ptr(), nm.data_ptr(), u.data_ptr(), p.shape[0], i.data_ptr(), i.shape[0], tex, tex, 0, tex, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream.cuda_stream)
            covered[name] = planet.last_covered_texels

    bake_steps = max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 3)):
        bake_step()
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for _ in range(bake_steps):
        bake_step()
    b1.record(stream)
    barrier()
    bake_ms = b0.elapsed_time(b1) / bake_steps
    # land alone, for the roofline of the colouring kernel
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    p, nm, u, i = dmesh["land"]
    o = outs["land"]
    planet.bake_device(0, p.data_ptr(), nm.data_ptr(), u.data_ptr(), p.shape[0], i.data_ptr(), i.shape[0], tex, tex, 0, tex, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream.cuda_stream)
    l1.record(stream)
    l1.synchronize()
    bake_land_ms = l0.elapsed_time(l1)
    clocks.__exit__()

    # end to end: host mesh in, host images out (pinned)
    himg = {"land": [torch.empty(tex * tex * c, dtype=torch.uint8, pin_memory=True) for c in (3, 3, 3)], "water": [torch.empty(tex * tex * c, dtype=torch.uint8, pin_memory=True) for c in (4, 3, 3)]}
    L = capi.load_library()
    import ctypes as C

    def bake_e2e():
        for name, water in (("land", 0), ("water", 1)):
            pos, nrm, uv, idx = meshes[name]
            o = himg[name]
            capi.check(L.upg_bake_export(planet._h, water, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), tex, tex, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))

    def bake_e2e_concurrent():
        # the generator bakes its chunks from pool threads (sources/generator.cpp:159); the entry points are re-entrant, so the upload
        # of one surface overlaps the colouring of the other
        import threading

        def one(name, water):
            pos, nrm, uv, idx = meshes[name]
            o = himg[name]
            capi.check(L.upg_bake_export(planet._h, water, pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(pos), idx.ctypes.data_as(C.c_void_p), len(idx), tex, tex, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))

        ts = [threading.Thread(target=one, args=a) for a in (("land", 0), ("water", 1))]
        for th in ts:
            th.start()
        for th in ts:
            th.join()

    bake_e2e()
    barrier()
    t = time.perf_counter()
    bake_e2e()
    bake_seq_s = time.perf_counter() - t
    bake_e2e_concurrent()
    barrier()
    t = time.perf_counter()
    bake_e2e_concurrent()
    bake_e2e_s = time.perf_counter() - t
    h2d_bake = sum(a.nbytes for m in meshes.values() for a in m)
    d2h_bake = tex * tex * (3 + 3 + 3 + 4 + 3 + 3)

    # ---- leg 5: a batch of planets back to back (BASELINE configs[3]: random base shape per seed, every elevation mode) ----
    batch_line = None
    if args.batch > 0:
        from unnatural_planets_b200 import ELEVATIONS, SHAPES, UpgError, synth

        try:
            del dmesh, outs
        except NameError:
            pass
        torch.cuda.empty_cache()
        bres, btex = args.batch_grid, args.batch_texture
        mine = list(range(rank, args.batch, world))  # replicas: planet i on rank i mod N, no communication
        bhost = {"land": [torch.empty(btex * btex * 3, dtype=torch.uint8, pin_memory=True) for _ in range(3)], "water": [torch.empty(btex * btex * c, dtype=torch.uint8, pin_memory=True) for c in (4, 3, 3)]}

        def batch_planet(i, timed):
            seed = seed_for(1000 + i)
            shape = SHAPES[seed[0] % len(SHAPES)]  # `--shape random` draws randomRange(0u, 34) (sources/modesConfiguration.cpp:121)
            elev = ELEVATIONS[i % len(ELEVATIONS)]
            t0 = time.perf_counter()
            bp = Planet(shape=shape, elevation=elev, coloring=COLORING, grid=bres, device=local_rank, seed=seed)
            try:
                final = bp.field_meshes_clean(FIELDS)
            except UpgError as ex:  # the reference throws for the same planets ("generated empty base land mesh", NaN sdf)
                bp.close()
                return {"i": i, "shape": shape, "elevation": elev, "error": ex.code}, time.perf_counter() - t0, 0.0, 0
            t_mesh = time.perf_counter() - t0
            # the generator would simplify, chunk and unwrap here (host geometry, out of scope): synthetic chart, outside the timing
            t_bake, texels_b = 0.0, 0
            for name, water in (("land", 0), ("water", 1)):
                m = final[name]
                if len(m["indices"]) == 0:
                    continue
                uv = synth.octahedral_uv(m["positions"])
                idx = synth.drop_folded_triangles(uv, m["indices"])
                o = bhost[name]
                t1 = time.perf_counter()
                capi.check(L.upg_bake_export(bp._h, water, m["positions"].ctypes.data_as(C.c_void_p), m["normals"].ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), len(uv), idx.ctypes.data_as(C.c_void_p), len(idx), btex, btex, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr())))
                t_bake += time.perf_counter() - t1
                texels_b += int(L.upg_bake_last_covered_texels())
            info = {"i": i, "shape": shape, "elevation": elev, "land_triangles": int(len(final["land"]["indices"]) // 3), "mesh_ms": round(t_mesh * 1e3, 2), "bake_ms": round(t_bake * 1e3, 2)}
            bp.close()
            return info, t_mesh, t_bake, texels_b

        if mine:
            batch_planet(mine[0], False)  # warms the pools for this grid / texture size
        barrier()
        infos, sum_mesh, sum_bake, sum_texels = [], 0.0, 0.0, 0
        for i in mine:
            info, tm, tb, tx = batch_planet(i, True)
            infos.append(info)
            sum_mesh += tm
            sum_bake += tb
            sum_texels += tx
        bt = torch.tensor([sum_mesh, sum_bake], dtype=torch.float64, device=dev)
        bc = torch.tensor([float(sum_texels), float(sum(1 for x in infos if "error" in x))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bc, op=dist.ReduceOp.SUM)
        if rank == 0:
            shapes_seen = sorted({x["shape"] for x in infos})
            batch_line = {
                "what": "BASELINE configs[3]: %d planets back to back, seed i -> shape = seed %% 34 (the reference's `random`), elevation = mode i %% 6, %d^3 grid, %d^2 land + water export bake; per planet: upg_planet_create + upg_field_meshes_clean + 2 x upg_bake_export, host buffers, wall clock of the C calls (synthetic UV chart on the host outside the timing); planets are replicas across ranks" % (args.batch, bres, btex),
                "planets": args.batch, "grid": bres, "texture": btex, "failed": int(bc[1]),
                "generation_ms_per_planet": (float(bt[0]) + float(bt[1])) * 1e3 / max(1, len(mine)), "planets_per_s": args.batch / max(1e-9, float(bt[0]) + float(bt[1])),
                "meshes_ms_total": float(bt[0]) * 1e3, "bake_ms_total": float(bt[1]) * 1e3, "baked_texels": int(bc[0]),
                "rank0_shapes": shapes_seen, "rank0_planets": infos,
            }

    texels = covered["land"] + covered["water"]
    times = torch.tensor([ms_total, e2e_s * 1e3, bake_ms, bake_e2e_s * 1e3, mesh_s * 1e3, mesh_raw_s * 1e3, contract_ms], dtype=torch.float64, device=dev)
    counts = torch.tensor([float(texels)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, bake_ms_max, bake_e2e_ms, mesh_ms, mesh_raw_ms, contract_ms_max = (float(x) for x in times)
    texels_all = float(counts[0])
    # the last collective of the job is behind us: the process group goes away on every rank together, the other ranks leave, and rank 0
    # finishes alone (CPU baseline, the isolated one-planet-over-N-GPUs legs, the JSON line)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    samples_per_step = len(FIELDS) * n * world
    value = samples_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = samples_per_step / (e2e_ms * 1e-3)

    peak = fp32_peak(capi, local_rank)
    fused_ms = ms_total / args.steps  # one launch per step
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    peak_source = "FP32 FMA ceiling measured in this run by upg_measure_fp32_peak (MEASURED_PEAKS.json has no FP32 SIMT figure)"

    def tflops(flops_per_node, ms):
        return flops_per_node * n / (ms * 1e-3) / 1e12

    # `achieved`: ALGORITHMIC flops of the launch over its time.  The numerator credits ONE evaluation of the reference's expression per
    # grid node -- one shape + one elevation function + the three combinations, instrumented count (804.5 flop at 1000^3) -- because that
    # is the work a launch that produces the three fields of a node has to do; it does NOT credit the kernel's own exact eliminations
    # (column caches, zero-mask skipping), which make the EXECUTED flops smaller still.  Both other views ride along as scalars: the
    # reference's arithmetic for the same 3e9 samples evaluated field by field like its three tasks do (can exceed the peak: two of
    # three evaluations are not needed), and the executed flops from the committed ncu instruction counts.
    achieved = tflops(WORK["one_evaluation"], fused_ms)
    three_eval = tflops(WORK["per_grid_node"], fused_ms)
    executed = tflops(K1_NCU["executed_flops_per_unit"], fused_ms) if K1_NCU else None
    issue_peak = None
    if K1_NCU and clocks.summary()["sm_mhz"]:
        issue_peak = 148 * 4 * 32 * clocks.summary()["sm_mhz"] * 1e6  # thread-instruction issue slots per second
    k1_ncu_this = _ncu("r02_k1_ops_%s.json" % ARITH) or K1_NCU
    roofline = {
        "bound": "fp32", "kernel": "sdf_column_kernel<sphere,lakes> (fused land + water + navigation launch, z-column walk), %s arithmetic build" % ARITH, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None, "traffic": (k1_ncu_this["dram_bytes_read"] + k1_ncu_this["dram_bytes_write"]) if k1_ncu_this else None, "peak_source": peak_source,
        "kernel_ms": fused_ms, "arithmetic": ARITH, "flops_per_grid_node": WORK["one_evaluation"], "samples_per_grid_node": len(FIELDS),
        "flops_source": "instrumented build of the CPU restatement, every 2nd node per axis of the 1000^3 grid (profiles/r02_work_counts_config3.json): one land evaluation + the water / navigation combinations",
        "frac_reference_three_evaluations": three_eval / peak if peak else None, "flops_reference_three_evaluations_per_node": WORK["per_grid_node"],
        "frac_executed": executed / peak if (peak and executed) else None, "executed_flops_per_grid_node": K1_NCU["executed_flops_per_unit"] if K1_NCU else None,
        "issue_slot_frac": (k1_ncu_this["issue_slots_x32_per_unit"] * n / (fused_ms * 1e-3) / issue_peak) if (issue_peak and k1_ncu_this) else None,
        "thread_instructions_per_grid_node": k1_ncu_this["thread_instructions_per_unit"] if k1_ncu_this else None,
        "fp32_instruction_share": (k1_ncu_this["fp32_instructions_per_unit"] / k1_ncu_this["thread_instructions_per_unit"]) if k1_ncu_this else None,
        "hbm_achieved_gbs": 4.0 * len(FIELDS) * n / (fused_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
        "other_arithmetic": OTHER, "other_kernel_ms": contract_ms_max, "other_samples_per_s": samples_per_step / (contract_ms_max * 1e-3), "other_frac": tflops(WORK["one_evaluation"], contract_ms_max) / peak if peak else None,
        "single_field_kernel_ms_land": single_ms["land"], "single_field_kernel_ms_water": single_ms["water"], "single_field_kernel_ms_navigation": single_ms["navigation"],
    }
    bake_achieved = WORK["land_texel"] * covered["land"] / (bake_land_ms * 1e-3) / 1e12
    texel_bake = {
        "metric": "baked_texels_per_s", "value": texels_all / (bake_ms_max * 1e-3), "unit": "texels/s", "steps": bake_steps, "ms_per_step": bake_ms_max,
        "texels_per_step": int(texels_all), "covered": {k: int(v) for k, v in covered.items()}, "texture": [tex, tex],
        "note": "texels = covered texels (one terrainTile each; the reference evaluates every (triangle, texel) visit) of the land (3+2+1 ch) and water (4+2+1 ch) surfaces; UVs: synthetic octahedral chart",
        "e2e": {"value": texels_all / (bake_e2e_ms * 1e-3), "unit": "texels/s", "h2d_bytes_per_step": int(h2d_bake), "d2h_bytes_per_step": int(d2h_bake), "path": "upg_bake_export: host mesh -> device, pinned host albedo / pbr / normal-map images <- device, wall clock"},
        "roofline": {"bound": "fp32", "kernel": "tile_color_kernel<default> + its K1 probes (land bake, whole upg_bake_device call)", "achieved": bake_achieved, "peak": peak, "unit": "TFLOP/s", "frac": bake_achieved / peak if peak else None,
                     "traffic": (K6_NCU["dram_bytes_per_unit"] * covered["land"]) if K6_NCU else None, "peak_source": peak_source, "call_ms": bake_land_ms, "flops_per_texel": WORK["land_texel"], "fp64_flops_per_texel": WORK["land_texel_fp64"]},
    }
    # the second half of the metric and the planet times as SCALARS inside `roofline` (the driver's record keeps scalars there)
    roofline.update({
        "texel_bake_texels_per_s": texel_bake["value"], "texel_bake_e2e_texels_per_s": texel_bake["e2e"]["value"], "texel_bake_ms_per_step": bake_ms_max, "texel_bake_land_call_ms": bake_land_ms,
        "texel_bake_frac": texel_bake["roofline"]["frac"], "texel_bake_flops_per_texel": WORK["land_texel"], "texel_bake_traffic": texel_bake["roofline"]["traffic"], "texel_bake_covered_texels": int(texels_all),
    })
    planet_ms = mesh_ms + bake_e2e_ms
    planet_line = {
        "ms": planet_ms, "base_meshes_ms": mesh_ms, "base_meshes_without_cleanup_ms": mesh_raw_ms, "bake_ms": bake_e2e_ms, "bake_one_after_the_other_ms": bake_seq_s * 1e3, "field_mesh_land_ms": field_mesh_land_s * 1e3,
        "what": "upg_field_meshes_clean (band sampling of the 3 fields + 3 concurrent Dual Marching Cubes extractions + K9 clean-up = the bodies of meshGenerateBaseLand/Water/Navigation + mesh downloads) + 2 x upg_bake_export (bake + pbr / normal-map tail; land and water from two host threads like the generator's chunk tasks), host buffers, wall clock",
        "mesh": mesh_stats, "base_meshes": final_stats, "mesh_sampling": sampling_stats,
    }
    roofline.update({"planet_ms": planet_ms, "planet_base_meshes_ms": mesh_ms, "planet_bake_ms": bake_e2e_ms, "planet_band_sampling_ms": sampling_stats["band_sampling_ms"], "planet_land_triangles": final_stats["land"]["triangles"]})

    threads = os.cpu_count() or 1
    cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "unavailable", "sample": ""}
    try:
        slices = max(1, min(res, (10 * threads * 250000) // (res * res)))  # ~10 s of CPU work in total, spread over every host thread
        cpu_rate, cpu_dt, kind = cpu_sampling_rate(res, slices, threads)
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": "land + water + navigation of %d z-slices x %d^2 (%.1f s on %d threads)" % (slices, res, cpu_dt, threads)}
        planet_line["cpu_sampling_ms_extrapolated"] = 3.0 * n / cpu_rate * 1e3
        cpu["planet_sampling_ms_extrapolated"] = planet_line["cpu_sampling_ms_extrapolated"]
        window = 256
        r = cpu_bake_rate(meshes["land"], tex, window, threads)
        if r:
            cpu["texel_bake_value"] = r[0]
            cpu["texel_bake_sample"] = "%d windows of %d^2 texels of the %d^2 land bake, one per thread (%.1f s, %d covered texels)" % (threads, window, tex, r[1], r[2])
            planet_line["cpu_bake_ms_extrapolated"] = texels_all / world / r[0] * 1e3
            cpu["planet_bake_ms_extrapolated"] = planet_line["cpu_bake_ms_extrapolated"]
    except Exception as ex:  # the baseline is reported, never required for the GPU number
        cpu["sample"] = "failed: %r" % (ex,)

    # ---- legs 4 / 4b: ONE planet over all GPUs of the job (BASELINE configs[4]), each leg in its own process (group) under a timeout ----
    sharded_line = multi_line = None
    if args.sharded_grid > 0:
        try:  # device tensors of the earlier legs (the batch leg may have dropped them already)
            del dmesh, outs
        except NameError:
            pass
        planet.close()
        torch.cuda.empty_cache()
        try:
            from unnatural_planets_b200 import planet as _pl

            _pl.trim(local_rank, 0)
        except Exception:
            pass
        if world > 1:
            wait_for_other_processes()
        sharded_line = run_isolated_leg("sharded", args, world, args.leg_timeout)
        multi_line = run_isolated_leg("multi", args, world, args.leg_timeout)
        export_line = run_isolated_leg("export", args, 1, args.leg_timeout)
        for name, rec in (("sharded_planet", sharded_line), ("multi_device_planet", multi_line), ("chunk_export", export_line)):
            if "error" in rec:
                print("[bench] leg %s failed: %s" % (name, rec["error"]), file=sys.stderr, flush=True)
                roofline[name + "_error"] = rec["error"][:200]
        if "error" in sharded_line:
            sharded_line = None
        if "error" in multi_line:
            multi_line = None
        if "error" not in export_line:
            roofline.update({k: v for k, v in export_line.items() if isinstance(v, (int, float, bool, str)) and k != "leg_wall_s"})
        if world == 1:
            # measurements the round could not take itself (its GPU budget ended with the 8-GPU run): the device meshChunking next to the
            # reference's, and the two unmeasured K1 experiments (streaming stores, block-shared Voronoi table) as an alternative build of
            # the library next to the shipped one -- each in its own process, none of them on the path of any number above
            rec = run_isolated_leg("chunking", args, 1, args.leg_timeout)
            if "error" in rec:
                roofline["chunking_error"] = rec["error"][:200]
            else:
                roofline.update({k: v for k, v in rec.items() if isinstance(v, (int, float, bool, str)) and k != "leg_wall_s"})
            variant = os.path.join(ROOT, "unnatural_planets_b200", "libupg_experiments.so")
            if os.path.exists(variant):
                import tempfile

                dumps = tempfile.mkdtemp(prefix="upg_bench_")
                da, db = os.path.join(dumps, "shipped.npy"), os.path.join(dumps, "experiments.npy")
                shipped = run_isolated_leg("experiments", args, 1, args.leg_timeout, extra_env={"UPG_EXPERIMENT_DUMP": da})
                other = run_isolated_leg("experiments", args, 1, args.leg_timeout, extra_env={"UPG_LIBRARY": variant, "UPG_EXPERIMENT_DUMP": db})
                for label, rec in (("shipped", shipped), ("experiments_build", other)):
                    if "error" in rec:
                        roofline["ab_%s_error" % label] = rec["error"][:200]
                    else:
                        roofline.update({"ab_%s_%s" % (label, k): v for k, v in rec.items() if k.endswith("_ms")})
                if "error" not in shipped and "error" not in other:
                    for k in shipped:
                        if k.endswith("_checksum"):
                            roofline["ab_experiments_build_%s_samples_identical" % k[: -len("_checksum")]] = bool(shipped[k] == other.get(k))
                    roofline["ab_experiments_build_what"] = "libupg_experiments.so = -DUPG_K1_STORE=1 (st.global.cs in the column kernel: lakes) -DUPG_CRATERS_TABLE=1 (block-shared Voronoi points in the sample kernel: craters) -DUPG_HSLUV_REAL=float (HSLuv of the colouring kernel in single precision: land_bake)"
                    try:
                        roofline.update({"ab_experiments_build_albedo_" + k: v for k, v in compare_albedo_dumps(da, db).items()})
                    except Exception as ex:
                        roofline["ab_experiments_build_albedo_error"] = repr(ex)[:200]
                for f in (da, db):
                    try:
                        os.remove(f)
                    except OSError:
                        pass
    if sharded_line:
        roofline.update({"sharded_planet_grid": sharded_line["grid"], "sharded_planet_texture": sharded_line["texture"], "sharded_planet_isosurface_ms": sharded_line["isosurface_ms"], "sharded_planet_bake_ms": sharded_line["bake_ms"]})
    if multi_line:
        roofline.update({k: v for k, v in multi_line.items() if isinstance(v, (int, float))})
    if batch_line:
        roofline.update({"batch_planets": batch_line["planets"], "batch_planets_per_s": batch_line["planets_per_s"], "batch_ms_per_planet": batch_line["generation_ms_per_planet"], "batch_failed": batch_line["failed"]})
        # the per-planet list is long: stderr, not the JSON line
        print("[bench] batch planets of rank 0: " + json.dumps(batch_line.pop("rank0_planets")), file=sys.stderr, flush=True)

    launches_sampling = args.steps  # one fused sdf_column_kernel launch per step
    param_bytes = int(capi.load_library().upg_kernel_param_bytes())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * param_bytes * max(1, (n * 4 + (64 << 20) - 1) // (64 << 20)), "d2h_bytes_per_step": 4 * n * len(FIELDS),
                "path": "3 x upg_field_samples -> pinned host buffer, wall clock incl. D2H; the only host -> device input of a generator step is the planet's parameter block (%d B), sent with every slab launch as kernel arguments" % param_bytes},
        "gpu_launches": launches_sampling,
        "texel_bake": texel_bake,
        "planet": planet_line,
        "sharded_planet": sharded_line,
        "multi_device_planet": multi_line,
        "batch": batch_line,
        "cpu_baseline": cpu,
        "roofline": roofline,
    }
    print(json.dumps(line), file=_STDOUT, flush=True)


def fp32_peak(capi, device):
    import ctypes as C

    v = C.c_float(0)
    capi.check(capi.load_library().upg_measure_fp32_peak(device, C.byref(v)))
    return float(v.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arithmetic", default="contract", choices=["exact", "contract"], help="arithmetic build of the dense sampling kernel the headline times (include/upg.h UPG_ARITH_*)")
    ap.add_argument("--grid", type=int, default=1000, help="samples per axis of the meshing grid (BASELINE configs[2]: 1000)")
    ap.add_argument("--texture", type=int, default=4096, help="bake resolution (BASELINE configs[2]: 4096)")
    ap.add_argument("--sharded-grid", type=int, default=1024, help="grid of the one-planet-over-all-ranks leg (BASELINE configs[4]: 1024; 0 = skip)")
    ap.add_argument("--batch", type=int, default=64, help="planets of the back-to-back leg (BASELINE configs[3]: 64; 0 = skip)")
    ap.add_argument("--batch-grid", type=int, default=500, help="grid of the batch planets (the reference's constant)")
    ap.add_argument("--batch-texture", type=int, default=2048, help="bake resolution of the batch planets")
    ap.add_argument("--sharded-texture", type=int, default=8192, help="bake resolution of that leg (BASELINE configs[4]: 8192)")
    ap.add_argument("--leg", default="", choices=["", "sharded", "multi", "export", "experiments", "chunking", "selftest", "hang"], help="internal: run ONE of the isolated one-planet-over-N-GPUs legs and print its record (selftest / hang: CPU checks of the orchestration)")
    ap.add_argument("--legs-budget", type=int, default=420, help="seconds all isolated legs together may take; legs that do not fit are skipped and reported as such")
    ap.add_argument("--orchestration-selftest", action="store_true", help="internal: the multi-process control flow of the bench on CPU (gloo), no GPU work")
    ap.add_argument("--leg-timeout", type=int, default=180, help="seconds each isolated leg may take before it is killed and reported as failed")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner does) are sent to stderr
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.leg in ("selftest", "hang"):
        run_selftest_leg(args, rank, world)
        return
    if args.orchestration_selftest:
        run_orchestration_selftest(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.leg:
            run_leg(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
