#!/usr/bin/env python3
"""Summaries of ncu captures for profiles/ (the raw .ncu-rep stays in gpurun_out/, which is scratch).

  python scripts/ncu_summary.py launches gpurun_out/x_launches.csv            -> per-kernel launch count / total ms / share
  python scripts/ncu_summary.py full gpurun_out/x.ncu-rep [kernel regex ...]  -> selected metrics per captured kernel
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex lsu wavefronts %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("kb2::", "")
    return name.strip()


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns / 1e6
    tot = sum(v[1] for v in agg.values())
    print("kernel | launches | total ms | share")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}%")
    print(f"TOTAL | {sum(v[0] for v in agg.values())} | {tot:.3f} |")


def full(path, pats):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    seen = set()
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kn = short(r[col["Kernel Name"]])
        if pats and not any(re.search(p, kn) for p in pats):
            continue
        if kn in seen:
            continue
        seen.add(kn)
        print(f"\n## `{kn}`\n\n| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in col:
                print(f"| {label} (`{m}`) | {r[col[m]]} {units[col[m]]} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3:])
