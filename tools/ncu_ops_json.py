#!/usr/bin/env python3
"""ncu --csv metric dump of ONE kernel launch -> small JSON of per-unit figures (committed under profiles/, read by bench.py):
    python tools/ncu_ops_json.py gpurun_out/r02a/k1_ops_exact.csv --units 1e9 --unit-name grid_node --out profiles/r02_k1_ops_exact.json"""
import argparse
import csv
import json

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--units", type=float, required=True, help="units (grid nodes, texels) the launch processed")
ap.add_argument("--unit-name", default="unit")
ap.add_argument("--out", required=True)
ap.add_argument("--note", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
m = {}
kernel = None
for r in rows[hi + 1:]:
    d = dict(zip(hdr, r))
    kernel = d["Kernel Name"]
    m[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
g = lambda k: m.get(k, 0.0)
fadd, fmul, ffma = g("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"), g("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"), g("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")
out = {
    "kernel": kernel, "source": a.csv, "note": a.note, "units": a.units, "unit": a.unit_name,
    "gpu_time_ms": g("gpu__time_duration.sum") * 1e-6,
    "executed_flops_per_unit": (fadd + fmul + 2 * ffma) / a.units,
    "fadd_per_unit": fadd / a.units, "fmul_per_unit": fmul / a.units, "ffma_per_unit": ffma / a.units,
    "fp32_instructions_per_unit": g("smsp__sass_thread_inst_executed_op_fp32_pred_on.sum") / a.units,
    "integer_instructions_per_unit": g("smsp__sass_thread_inst_executed_op_integer_pred_on.sum") / a.units,
    "conversion_instructions_per_unit": g("smsp__sass_thread_inst_executed_op_conversion_pred_on.sum") / a.units,
    "control_instructions_per_unit": g("smsp__sass_thread_inst_executed_op_control_pred_on.sum") / a.units,
    "memory_instructions_per_unit": g("smsp__sass_thread_inst_executed_op_memory_pred_on.sum") / a.units,
    "thread_instructions_per_unit": g("smsp__thread_inst_executed.sum") / a.units,
    "issue_slots_x32_per_unit": g("smsp__inst_executed.sum") * 32.0 / a.units,
    "issue_active_frac": g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "pipe_fma_warp_instructions": g("sm__inst_executed_pipe_fma.sum"), "pipe_alu_warp_instructions": g("sm__inst_executed_pipe_alu.sum"),
    "pipe_xu_warp_instructions": g("sm__inst_executed_pipe_xu.sum"), "pipe_lsu_warp_instructions": g("sm__inst_executed_pipe_lsu.sum"),
    "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
    "dram_bytes_per_unit": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / a.units,
    "registers_per_thread": g("launch__registers_per_thread"), "warps_active_per_scheduler": g("smsp__warps_active.avg.per_cycle_active"),
}
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out, indent=1))
