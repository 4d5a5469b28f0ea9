#!/usr/bin/env python3
"""Markdown summary of `ncu --set full` captures (.ncu-rep read HERE with `ncu -i`): speed of light, scheduler, occupancy, pipe
utilisation, warp stall reasons per issued instruction.  tools/ncu_details_md.py out.md label=path.ncu-rep [label=path ...]"""
import csv
import subprocess
import sys


def page(path, name):
    out = subprocess.run(["ncu", "-i", path, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def details(path):
    rows = page(path, "details")
    h = rows[0]
    d = {}
    kernel = None
    for r in rows[1:]:
        x = dict(zip(h, r))
        kernel = kernel or x.get("Kernel Name")
        if x.get("Metric Name"):
            d[(x["Section Name"], x["Metric Name"])] = (x["Metric Value"], x["Metric Unit"])
    return kernel, d


def raw(path):
    rows = page(path, "raw")
    return dict(zip(rows[0], rows[2]))


PICK = [
    ("GPU Speed Of Light Throughput", "Duration"), ("GPU Speed Of Light Throughput", "SM Frequency"), ("GPU Speed Of Light Throughput", "Compute (SM) Throughput"),
    ("GPU Speed Of Light Throughput", "Memory Throughput"), ("GPU Speed Of Light Throughput", "DRAM Throughput"), ("GPU Speed Of Light Throughput", "L1/TEX Cache Throughput"),
    ("GPU Speed Of Light Throughput", "L2 Cache Throughput"), ("Compute Workload Analysis", "Executed Ipc Active"), ("Compute Workload Analysis", "Issue Slots Busy"),
    ("Memory Workload Analysis", "Memory Throughput"), ("Memory Workload Analysis", "L1/TEX Hit Rate"), ("Memory Workload Analysis", "L2 Hit Rate"),
    ("Scheduler Statistics", "Active Warps Per Scheduler"), ("Scheduler Statistics", "Eligible Warps Per Scheduler"), ("Scheduler Statistics", "No Eligible"),
    ("Warp State Statistics", "Warp Cycles Per Issued Instruction"), ("Warp State Statistics", "Avg. Active Threads Per Warp"), ("Warp State Statistics", "Avg. Not Predicated Off Threads Per Warp"),
    ("Launch Statistics", "Grid Size"), ("Launch Statistics", "Block Size"), ("Launch Statistics", "Registers Per Thread"), ("Launch Statistics", "Static Shared Memory Per Block"),
    ("Launch Statistics", "Dynamic Shared Memory Per Block"), ("Launch Statistics", "Waves Per SM"), ("Occupancy", "Block Limit Registers"), ("Occupancy", "Block Limit Shared Mem"),
    ("Occupancy", "Theoretical Occupancy"), ("Occupancy", "Achieved Occupancy"), ("Source Counters", "Branch Efficiency"),
]


def main():
    out, items = sys.argv[1], [a.split("=", 1) for a in sys.argv[2:]]
    caps = []
    for label, path in items:
        kernel, d = details(path)
        caps.append((label, path, kernel, d, raw(path)))
    lines = ["# ncu `--set full` captures: what bounds each kernel", "", "Read with `ncu -i <file> --page details|raw --csv` (tools/ncu_details_md.py).  Times under ncu are single cold launches with replay;", "the shares and ratios are what carries over to the bench.", ""]
    lines += ["| | " + " | ".join(l for l, *_ in caps) + " |", "|---|" + "---|" * len(caps)]
    lines.append("| kernel | " + " | ".join("`%s`" % (k or "")[:90] for _, _, k, _, _ in caps) + " |")
    for key in PICK:
        vals = []
        for _, _, _, d, _ in caps:
            v = d.get(key)
            vals.append("%s %s" % v if v else "")
        lines.append("| %s: %s | " % key + " | ".join(vals) + " |")
    lines += ["", "## Warp stall reasons (warps stalled per issued instruction; `selected` = 1 is the issuing warp itself)", "", "| reason | " + " | ".join(l for l, *_ in caps) + " |", "|---|" + "---|" * len(caps)]
    reasons = sorted({k for *_, r in caps for k in r if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")})
    table = []
    for k in reasons:
        name = k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        vals = [float(r.get(k, "0").replace(",", "") or 0) for *_, r in caps]
        table.append((max(vals), name, vals))
    for _, name, vals in sorted(table, reverse=True):
        lines.append("| %s | " % name + " | ".join("%.2f" % v for v in vals) + " |")
    lines += ["", "## Pipe utilisation (% of peak, active cycles)", "", "| pipe | " + " | ".join(l for l, *_ in caps) + " |", "|---|" + "---|" * len(caps)]
    pipes = sorted({k for *_, r in caps for k in r if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active")})
    rowsp = []
    for k in pipes:
        vals = [float(r.get(k, "0").replace(",", "") or 0) for *_, r in caps]
        if max(vals) >= 1.0:
            rowsp.append((max(vals), k[len("sm__inst_executed_pipe_"):-len(".avg.pct_of_peak_sustained_active")], vals))
    for _, name, vals in sorted(rowsp, reverse=True):
        lines.append("| %s | " % name + " | ".join("%.1f" % v for v in vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
