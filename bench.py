#!/usr/bin/env python3
"""bench.py -- throughput of the generation hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = the dense SDF sampling of one planet: the three terrain fields the generator meshes (land, water,
navigation; sources/generator.cpp:432-439) over the 500^3 meshing grid (sources/meshGeneration.cpp:19-28) of the
reference's default configuration (`--shape sphere`, elevation lakes; BASELINE.json configs[0]) = 3.75e8 samples.
Weak scaling: every rank samples one such planet per step (planet i of the batch gets the seed stream advanced by
i draws), there is no data-path collective; `value` = samples of all ranks / max-over-ranks device time.

Prints ONE JSON line (rank 0).  `value`: inputs (a ~1 KB parameter block) and outputs resident in HBM.
`e2e`: the same samples through the host-buffer C-ABI call upg_field_samples (device->host copies inside the timed
region).  `roofline`: K1 against the FP32 FMA ceiling measured in this run.  `cpu_baseline`: the reference's own
sampling loop (oracle/_ref, the reference sources compiled verbatim) on a bounded z-slab sample, all host threads.
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

METRIC = "sdf_samples_per_s"
UNIT = "samples/s"
RES = 500
FIELDS = ("land", "water", "navigation")
SHAPE, ELEVATION = "sphere", "lakes"
WORKLOAD = "BASELINE configs[0]: --shape sphere (elevation lakes, default seed), 3 fields (land, water, navigation) x 500^3 grid per planet"
# algorithmic FP32 work of one sample (FMA = 2), per field, for sphere+lakes: see DESIGN.md "K1 work per sample"
# (instruction counts of the un-contracted arithmetic, cross-checked with ncu smsp__sass_thread_inst_executed_op_f*)
FLOPS_PER_SAMPLE = {"land": 640.0, "navigation": 640.0, "water": 9.0}


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
    """polls SM clock / throttle reasons during the timed region (B200_PROFILING.md 'clocks line')"""

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
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_rate(slices, threads):
    """the reference's sampling loop (oracle/_ref) on `slices` z-slices around the equator of the land field"""
    from oracle import ref as oref

    if not oref.available():
        from oracle import port as oport  # the standalone restatement

        return oport.sample_rate(SHAPE, ELEVATION, RES, slices, threads) + ("port",)
    planet = oref.RefPlanet(shape=SHAPE, elevation=ELEVATION)
    z0 = RES // 2 - slices // 2
    t = time.perf_counter()
    planet.grid("land", RES, z0=z0, z1=z0 + slices, threads=threads)
    dt = time.perf_counter() - t
    return slices * RES * RES / dt, dt, "reference"


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    slices = 12 * threads  # ~12 x 250k samples per thread: a few seconds of CPU work per step
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, kind = cpu_reference_rate(slices, threads)
        if i >= args.warmup:
            rates.append((r, dt))
    value = statistics.mean(r for r, _ in rates)
    ms = statistics.mean(dt for _, dt in rates) * 1e3
    sample = "%d z-slices x %d^2 of the land field per step (%.3g of one planet's 3.75e8 samples), %d threads" % (slices, RES, slices * RES * RES / (3.0 * RES**3), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "host CPU, no GPU involved"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from unnatural_planets_b200 import Planet, capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    planet = Planet(shape=SHAPE, elevation=ELEVATION, grid=RES, device=local_rank, seed=seed_for(rank))
    n = RES**3
    bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in FIELDS]
    stream = torch.cuda.current_stream()

    def step():
        for f, b in zip(FIELDS, bufs):
            planet.field_samples_device(f, 0, RES, b.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)

    # per-kernel timing of the dominant kernel (land field) with CUDA events on the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    for _ in range(max(3, args.steps)):
        k0.record(stream)
        planet.field_samples_device("land", 0, RES, bufs[0].data_ptr(), stream.cuda_stream)
        k1.record(stream)
        k1.synchronize()
        kernel_ms.append(k0.elapsed_time(k1))
    land_ms = statistics.mean(kernel_ms)

    # end to end through the host-buffer C-ABI call: device->host copies inside the timed region
    host = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in FIELDS]
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        for f, h in zip(FIELDS, host):
            planet.field_samples(f, 0, RES, h.numpy().reshape(RES, RES, RES))

    e2e_step()
    barrier()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t) / e2e_steps

    times = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(times[0]), float(times[1])
    if rank != 0:
        return

    samples_per_step = len(FIELDS) * n * world
    value = samples_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = samples_per_step / (e2e_ms * 1e-3)

    peak = C_float_peak(capi, local_rank)
    flops_land = FLOPS_PER_SAMPLE["land"] * n
    achieved = flops_land / (land_ms * 1e-3) / 1e12
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    roofline = {
        "bound": "fp32", "kernel": "sdf_sample_kernel<sphere,lakes> (land field)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None, "traffic": None,
        "peak_source": "FP32 FMA ceiling measured in this run by upg_measure_fp32_peak (MEASURED_PEAKS.json has no FP32 SIMT figure)",
        "kernel_ms": land_ms, "flops_per_sample": FLOPS_PER_SAMPLE["land"],
        "hbm": {"achieved_gbs": 4.0 * n / (land_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "note": "4 B/sample written; the kernel is compute bound"},
    }

    threads = os.cpu_count() or 1
    slices = min(RES, 30 * threads)  # 10-30 s of CPU work in total, spread over every host thread
    try:
        cpu_rate, cpu_dt, kind = cpu_reference_rate(slices, threads)
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": "%d z-slices x %d^2 of the land field (%.1f s of CPU work on %d threads)" % (slices, RES, cpu_dt, threads)}
    except Exception as ex:  # the baseline is reported, never required for the GPU number
        cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "unavailable", "sample": str(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grid": RES, "fields": list(FIELDS), "planets_per_step": world, "l2": "each step writes 1.5 GB of samples per GPU (>> 126 MB L2); inputs are a ~1 KB parameter block"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n * len(FIELDS), "path": "upg_field_samples -> pinned host buffers, wall clock incl. D2H"},
        "gpu_launches": len(FIELDS) * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def C_float_peak(capi, device):
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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
