#!/usr/bin/env python3
"""Summarise an ncu report (read here, without a GPU) into the small text files kept under profiles/.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_v2_prof      -> *_raw.csv (selected metrics), *_details.txt
  python tools/ncu_summary.py --launches gpurun_out/launches.csv                -> per-kernel totals of a launch list"""
import csv
import subprocess
import sys
from collections import defaultdict

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "sm__inst_executed_pipe",
        "smsp__average_warp", "smsp__warp_issue_stalled", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_local", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "sm__pipe_fma", "sm__pipe_alu", "smsp__inst_executed_pipe", "gpc__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hd = rows[h]
    ki, vi, ui = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[h + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0]
        tot[n] += v
        cnt[n] += 1
    unit = rows[h + 1][ui]
    all_ = sum(tot.values())
    print("kernel,launches,total_%s,avg_%s,share" % (unit, unit))
    for n in sorted(tot, key=lambda n: -tot[n]):
        print("%s,%d,%.0f,%.0f,%.4f" % (n, cnt[n], tot[n], tot[n] / cnt[n], tot[n] / all_))


def report(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hd, units = rows[0], rows[1]
    with open(prefix + "_raw.csv", "w") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            w.writerow(["kernel", r[hd.index("Kernel Name")]])
            for h, u, v in zip(hd, units, r):
                if any(h.startswith(k) for k in KEEP):
                    w.writerow([h, u, v])
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(prefix + "_details.txt", "w").write(det)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1], sys.argv[2])
