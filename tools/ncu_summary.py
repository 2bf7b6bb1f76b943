#!/usr/bin/env python
"""Summarise an Nsight Compute report for profiles/: one JSON object per profiled launch with the
metrics DESIGN.md argues from (duration, DRAM bytes, pipe/throughput percentages, hit rates, occupancy,
top warp-stall reasons).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "what was run" > profiles/ncu_rNN_summary.json
    python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/launches_rNN_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    # on-chip traffic and pipes (round 2: what actually binds the bit-sliced kernels)
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    head, units = rd[0], rd[1]
    return head, units, rd[2:]


def summarise(rep, note):
    head, units, rows = rows_of(rep)
    col = {h: i for i, h in enumerate(head)}
    kernels = []
    for r in rows:
        k = {"kernel": r[col["Kernel Name"]], "id": r[col["ID"]]}
        for m in KEEP:
            if m in col:
                k[m] = ("%s %s" % (r[col[m]], units[col[m]])).strip()
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX) and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        k["top_stalls_warps_per_issue"] = {n: round(v, 3) for v, n in stalls[:5]}
        kernels.append(k)
    return {"source": note, "report": rep, "kernels": kernels}


def launches(path):
    """Launch list of `ncu --metrics gpu__time_duration.sum --csv --log-file`: per kernel count / total / share."""
    text = open(path).read()
    start = text.index('"ID"')
    rd = list(csv.DictReader(io.StringIO(text[start:])))
    agg = {}
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(r["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    return {"source": path, "total_us": round(total, 2),
            "kernels": {k: {"launches": a[0], "us_total": round(a[1], 2), "us_avg": round(a[1] / a[0], 2),
                            "share": round(a[1] / total, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}}


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        print(json.dumps(launches(sys.argv[2]), indent=1))
    else:
        print(json.dumps(summarise(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""), indent=1))
