#!/usr/bin/env python
"""Summarise `ncu --set full` reports into profiles/<round>/ncu_full_summary.json: python profiles/ncu_summarise.py out.json name=report.ncu-rep ...
(reads each report with `ncu -i ... --page raw --csv`; keeps the metrics DESIGN.md and bench.py's roofline.traffic quote)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def summarise(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for n, u, v in zip(names, units, vals):
        if n == "Kernel Name" or n in KEEP:
            out[n] = {"value": v, "unit": u}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in out:
            tot += float(out[k]["value"].replace(",", "")) * SCALE.get(out[k]["unit"], 1.0)
    out["dram_bytes_total"] = tot
    return out


if __name__ == "__main__":
    dst = sys.argv[1]
    try:
        cur = json.load(open(dst))
    except Exception:
        cur = {}
    for a in sys.argv[2:]:
        name, rep = a.split("=", 1)
        cur[name] = summarise(rep)
    json.dump(cur, open(dst, "w"), indent=1)
    print(json.dumps({k: {"ms": v.get("gpu__time_duration.sum"), "dram": v["dram_bytes_total"]} for k, v in cur.items()}, indent=1))
