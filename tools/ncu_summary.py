#!/usr/bin/env python
"""Summarise Nsight Compute captures for profiles/: per kernel, the counters the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...] --out profiles/r2_ncu_full.txt \
        [--traffic profiles/ncu_traffic.json] [--note "what was captured"]

Reads each report with `ncu -i <rep> --page raw --csv` (works without a GPU).  --traffic updates the JSON table that
bench.py reads `roofline.traffic` from: {kernel name: {"dram_bytes_per_launch": read + write, "source": report}}.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.min.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(header)}
    recs = []
    for r in rows[2:]:
        if len(r) != len(header):
            continue
        rec = {"name": r[col["Kernel Name"]], "id": r[col["ID"]]}
        for m in KEEP:
            if m in col:
                rec[m] = (r[col[m]], units[col[m]])
        recs.append(rec)
    return recs


def to_bytes(val, unit):
    return float(val.replace(",", "")) * UNIT_SCALE.get(unit, 1.0)


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    lines = []
    if a.note:
        lines.append("# " + a.note)
    table = {}
    if a.traffic and os.path.exists(a.traffic):
        table = json.load(open(a.traffic))
    for rep in a.reports:
        lines.append(f"# report: {os.path.basename(rep)}")
        for rec in read_report(rep):
            lines.append(f"\n## [{rec['id']}] {rec['name'][:150]}")
            for m in KEEP:
                if m in rec:
                    v, u = rec[m]
                    lines.append(f"{m:<100} {v:>16} {u}")
            if "dram__bytes_read.sum" in rec and "dram__bytes_write.sum" in rec:
                total = to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])
                lines.append(f"{'dram bytes read + write per launch':<100} {total / 1e6:>16.1f} MB")
                key = short(rec["name"])
                # keep the LAST launch of a kernel in a report (earlier ones are warm-ups)
                table[key] = {"dram_bytes_per_launch": total, "source": os.path.basename(rep),
                              "duration_ms_under_ncu": float(rec["gpu__time_duration.sum"][0].replace(",", ""))
                              * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(rec["gpu__time_duration.sum"][1], 1.0)
                              if "gpu__time_duration.sum" in rec else None, "kernel": rec["name"][:120]}
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    if a.traffic:
        with open(a.traffic, "w") as f:
            json.dump(table, f, indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
