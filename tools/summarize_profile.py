#!/usr/bin/env python
"""Summarise an ncu capture for profiles/: per-kernel launch table from a `--metrics gpu__time_duration.sum` CSV log and
the headline counters of a `--set full` .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed).

    python tools/summarize_profile.py gpurun_out/r1c profiles/r01
"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("t2d::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.2f | %.1f%% |\n" % (k, n, t / 1e3, t / 1e3 / n, 100 * t / tot))


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; one column per captured launch\n")
        names = [r[ik].split("(")[0].replace("void ", "").replace("t2d::", "") for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches(os.path.join(src, "launches.csv"), dst + "_launches.md")
    if os.path.exists(os.path.join(src, "prof.ncu-rep")):
        full(os.path.join(src, "prof.ncu-rep"), dst + "_ncu_full.md")
    for f in ("bench.json", "bench_c2.json", "bench_c3.json", "bench_ref.json", "gpu.txt"):
        p = os.path.join(src, f)
        if os.path.exists(p) and os.path.getsize(p):
            open(dst + "_" + f, "w").write(open(p).read())
