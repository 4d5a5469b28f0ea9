#!/usr/bin/env python3
"""profiles/k1_column_ncu.json (what bench.py reads roofline.traffic and the issue-slot view from) out of an `ncu --set full` capture of the
fused dense K1 launch:  python tools/k1_capture_json.py gpurun_out/<report>.ncu-rep "<how it was captured>" [grid] > profiles/k1_column_ncu.json"""
import csv
import io
import json
import subprocess
import sys

rep, how = sys.argv[1], sys.argv[2]
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]


def val(key, unit_scale=None):
    v = float(r[hdr.index(key)].replace(",", ""))
    u = units[hdr.index(key)]
    if unit_scale and u in unit_scale:
        v *= unit_scale[u]
    return v


BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
nodes = grid**3
rd, wr = val("dram__bytes_read.sum", BYTES), val("dram__bytes_write.sum", BYTES)
inst = val("smsp__inst_executed.sum")
print(json.dumps({
    "source": "%s: %s" % (rep, how),
    "kernel": r[hdr.index("Kernel Name")],
    "grid_nodes": nodes,
    "dram__bytes_read.sum": rd,
    "dram__bytes_write.sum": wr,
    "dram_bytes_per_grid_node": (rd + wr) / nodes,
    "algorithmic_bytes_per_grid_node": 12,
    "gpu__time_duration_ms": val("gpu__time_duration.sum", TIME),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_alu.pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_fma.pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_xu.pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_lsu.pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "smsp__inst_executed.sum": inst,
    "registers_per_thread": val("launch__registers_per_thread"),
    "sm__warps_active.pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "smsp__thread_inst_executed_per_inst_executed": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "issue_slots_x32_per_grid_node": inst * 32.0 / nodes,
}, indent=1))
