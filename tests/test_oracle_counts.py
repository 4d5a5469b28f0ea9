"""The counted builds behind the work-per-unit figures (oracle/up_count.cpp, tests/hostlib/upg_host_count.cpp): the same
sources with `float` / `double` replaced by an instrumented type must return the SAME BITS as the plain builds (tools/count_ops.py
asserts that on the grids it counts), and the counts themselves have closed forms where the code is short enough to count by hand."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import count_ops  # noqa: E402


def test_counted_sampling_matches_plain_oracle_and_hand_counts():
    L = count_ops.counted_lib()
    water = count_ops.count_field(L, "sphere", "none", "water", 24, 1, 2)  # asserts bit-identity on two slices internally
    # position(): per axis 4 fadd + 1 fmul + 1 fdiv (marchingCubes.cpp:247-257); sdfSphere: length = 3 fmul + 3 fadd (dot starts at 0) + sqrt, - radius
    assert water["fmul"] == 6 and water["fdiv"] == 3 and water["fsqrt"] == 1 and water["fadd"] == 16
    assert water["special"] == 0 and water["nodes"] == 24**3
    land = count_ops.count_field(L, "sphere", "none", "land", 24, 1, 2)
    # elevation "none" is the constant 100: land = shape - 100 / 10
    assert land["flops"] == water["flops"] + 2
    lakes = count_ops.count_field(L, "sphere", "lakes", "land", 40, 1, 2)
    assert 500 < lakes["flops"] < 1100 and 1 < lakes["special"] < 4  # four Value octaves + pow always; mountains where cover > 0


def test_counted_colouring_matches_plain_host_build():
    out = count_ops.count_texels("sphere", "lakes", "default", 48)  # asserts bit-identity of all 13 tile outputs internally
    assert 4000 < out["land"]["flops"] < 9000 and out["land"]["flops_fp64"] > 100
    assert out["water"]["flops"] < out["land"]["flops"]


def test_committed_work_counts_are_what_bench_reads():
    c = json.load(open(os.path.join(ROOT, "profiles", "r02_work_counts_config3.json")))
    assert c["grid"] == 1000 and c["shape"] == "sphere" and c["elevation"] == "lakes"
    f = c["fields"]
    assert abs(c["flops_per_grid_node"] - (f["land"]["flops"] + f["water"]["flops"] + f["navigation"]["flops"])) < 1e-6
    assert c["texels"]["land"]["points"] > 50000
