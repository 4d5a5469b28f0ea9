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
