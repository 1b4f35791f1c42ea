#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of one kernel launch in an .ncu-rep (needs -lineinfo and
--import-source on).  Usage: ncu_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, l = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    for h, u, v in zip(hdr, units, l):
        if h == w:
            print(f"{h:84s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
cur = None
ix = None
agg = []
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ix = {k: i for i, k in enumerate(r)}
        continue
    if r[0] in ("Function Name",) or ix is None or r[0] == "":
        continue
    try:
        samp = int(r[ix["# Samples"]])
        ex = int(r[ix["Instructions Executed"]])
        lsb = int(r[ix["stall_long_sb"]])
        wait = int(r[ix["stall_wait"]])
        loc = int(r[ix["L2 Theoretical Sectors Local"]])
        glb = int(r[ix["L2 Theoretical Sectors Global"]])
    except (ValueError, KeyError):
        continue
    agg.append((samp, ex, lsb, wait, loc, glb, cur, r[0], r[1].strip()[:84]))
tot = sum(a[0] for a in agg) or 1
print("samples", tot, "warp instrs", sum(a[1] for a in agg), "L2 local sectors", sum(a[4] for a in agg), "global", sum(a[5] for a in agg))
for a in sorted(agg, reverse=True)[:top]:
    print(f"{a[0]:6d} {100 * a[0] / tot:5.1f}% ex={a[1]:9d} lsb={a[2]:6d} wait={a[3]:5d} loc={a[4]:8d} glb={a[5]:9d} {a[6]}:{a[7]} {a[8]}")
