"""bench.py's multi-process control flow, on CPU (gloo): the last collective, the process group torn down on every rank, the other ranks
leaving, rank 0 running the isolated one-planet-over-N-GPUs legs -- one nested torchrun launch, one plain process, one that hangs and
is killed at its timeout -- and exactly one JSON line on stdout.  (The 8-GPU run of round 2 lost its bench line to a leg that hung.)"""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_isolated_legs_under_torchrun():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "bench.py"), "--orchestration-selftest"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, lines  # stdout carries exactly one JSON line
    rec = json.loads(lines[0])
    assert rec["nested"]["selftest_sum"] == 3.0 and rec["nested"]["world"] == 2  # the nested launch formed its own two-rank group
    assert rec["single"]["selftest_sum"] == 1.0 and rec["single"]["inherited_rank_env"] == []  # no RANK / WORLD_SIZE leaked into the plain process
    assert rec["hung"] == {"error": "timeout after 5 s"}


def test_isolated_leg_commands():
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    a = argparse.Namespace(sharded_grid=1024, sharded_texture=8192, texture=4096, grid=1000)
    multi = bench.isolated_leg_command("multi", a, 8, 1234)
    assert multi[0] == sys.executable and "torch.distributed.run" not in multi and multi[-6:] == ["--gpus", "8", "--sharded-grid", "1024", "--sharded-texture", "8192"]
    sharded = bench.isolated_leg_command("sharded", a, 8, 1234)
    assert sharded[1:3] == ["-m", "torch.distributed.run"] and "--nproc-per-node" in sharded and sharded[sharded.index("--nproc-per-node") + 1] == "8" and "127.0.0.1" in sharded
    assert "torch.distributed.run" not in bench.isolated_leg_command("sharded", a, 1, 1234)
    export = bench.isolated_leg_command("export", a, 8, 1234)  # one process on one GPU whatever the size of the job
    assert "torch.distributed.run" not in export and export[2:6] == ["--leg", "export", "--gpus", "1"] and export[-2:] == ["--texture", "4096"]


def test_bench_helpers_of_the_ab_leg():
    """the pure helpers of bench.py's A/B leg (the leg itself needs a GPU): checksums of sample fields, comparison of two albedo dumps"""
    sys.path.insert(0, ROOT)
    import tempfile

    import numpy as np
    import torch

    import bench

    rng = np.random.default_rng(3)
    x = rng.normal(size=5000).astype(np.float32)
    y = x.copy()
    y[1234] = np.nextafter(y[1234], np.float32(10))
    cx = bench.tensor_checksums([torch.from_numpy(x), torch.from_numpy(x[:100].copy())], step=1024)
    assert cx == bench.tensor_checksums([torch.from_numpy(x.copy()), torch.from_numpy(x[:100].copy())], step=1024) and len(cx) == 32
    assert cx != bench.tensor_checksums([torch.from_numpy(y), torch.from_numpy(x[:100].copy())], step=1024)  # one ulp in one sample shows
    bits = x.view(np.int32).astype(np.int64)
    plain = (int(bits.sum()) + int(bits[:100].sum())) & ((1 << 64) - 1)
    assert cx[:16] == "%016x" % plain
    d = tempfile.mkdtemp()
    a = rng.integers(0, 255, size=(64, 64, 3)).astype(np.uint8)
    b = a.copy()
    b[3, 4, 1] += 1
    b[7, 7, 2] = (int(b[7, 7, 2]) + 9) % 256 if b[7, 7, 2] < 200 else b[7, 7, 2] - 9
    np.save(os.path.join(d, "a.npy"), a.reshape(-1))
    np.save(os.path.join(d, "b.npy"), b.reshape(-1))
    r = bench.compare_albedo_dumps(os.path.join(d, "a.npy"), os.path.join(d, "b.npy"))
    assert r["max_step"] == 9 and abs(r["differing_share"] * r["texels"] - 2) < 1e-6 and abs(r["more_than_one_step_share"] * r["texels"] - 1) < 1e-6

