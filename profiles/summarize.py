#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/*.ncu-rep, launch-list CSVs) into the small tracked summaries in profiles/.
usage: python profiles/summarize.py rep <file.ncu-rep> <out.md> | launches <launches.csv> <out.md>"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % path)
        for d in data:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, d[i], units[i]))
            f.write("\n")


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of %s\n\n" % path)
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.2f | %.3f |\n" % (k, n, t / 1e3, t / n / 1e3, t / tot))


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
