#!/usr/bin/env python3
"""Turn the two ncu outputs of a bench run into the markdown summary kept under profiles/.

  python profiles/ncu_summary.py LAUNCHES.csv FULL_RAW.csv > profiles/rNN_..._ncu.md

LAUNCHES.csv : `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`
FULL_RAW.csv : `ncu -i capture.ncu-rep --page raw --csv` of a `--set full` capture of the dominant kernel
"""
import csv
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        name = r[k].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.2f%% |" % (name, c, ms, 100 * ms / tot))


def full(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("| metric | value | unit |\n|---|---:|---|")
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            print("| `%s` | %s | %s |" % (m, vals[i], units[i]))
    stall = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stall[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(vals[i].replace(",", ""))
            except ValueError:
                pass
    tot = sum(stall.values())
    if tot:
        print("\nTop warp stall reasons (pc sampling):\n\n| reason | share |\n|---|---:|")
        for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]:
            print("| %s | %.1f%% |" % (k, 100 * v / tot))


if __name__ == "__main__":
    print("## Launch list of the timed steps\n")
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        print("\n## `ncu --set full` of the dominant kernel (one launch)\n")
        full(sys.argv[2])
