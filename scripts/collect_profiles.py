#!/usr/bin/env python
"""Copy one GPU visit (gpurun_out/<tag>/, written by scripts/gpu_round.sh) into profiles/<tag>/ as
small text summaries: bench lines, pytest/smoke logs, the ncu launch list, selected raw metrics of the
`--set full` capture of k_iterate and the opcode/stall summary of its source page.
Usage: python scripts/collect_profiles.py <tag>"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles", tag)
os.makedirs(dst, exist_ok=True)

for name in os.listdir(src):
    if name.endswith((".json", ".log", ".txt")) and os.path.getsize(os.path.join(src, name)) > 0:
        if name in ("ncu_full.log", "ncu_launches.log"):
            continue
        shutil.copy(os.path.join(src, name), os.path.join(dst, name))
if os.path.exists(os.path.join(src, "launches.csv")):
    shutil.copy(os.path.join(src, "launches.csv"), os.path.join(dst, "ncu_launches.csv"))

WANT = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sass__inst_executed_register_spilling", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum",
]

for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    base = rep[:-len(".ncu-rep")]
    path = os.path.join(src, rep)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no raw page in", rep)
        continue
    hdr, units, launches = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    with open(os.path.join(dst, f"ncu_{base}_full_metrics.csv"), "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        f.write("Kernel Name,," + ",".join('"%s"' % l[ix["Kernel Name"]] for l in launches) + "\n")
        for m in WANT:
            if m in ix:
                f.write(f"{m},{units[ix[m]]}," + ",".join(l[ix[m]].replace(",", "") for l in launches) + "\n")
    # DRAM traffic of the first launch -> profiles/iterate_traffic.json (bench.py's roofline.traffic)
    keymap = {"prof_iterate": "lattice-1000x1000", "prof_iterate_rings": "rings-100000",
              "prof_axis": "lattice-1000x1000", "prof_axis_rings": "rings-100000"}
    if base in keymap:
        def to_bytes(col):
            v = float(launches[0][ix[col]].replace(",", ""))
            u = units[ix[col]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        tj = os.path.join(ROOT, "profiles", "iterate_traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
        d[keymap[base]] = {
            "dram_bytes_per_launch": rd + wr,
            "source": f"profiles/{tag}/ncu_{base}_full_metrics.csv (dram__bytes_read.sum {rd / 1e6:.2f} MB + "
                      f"dram__bytes_write.sum {wr / 1e6:.2f} MB, {launches[0][ix['Kernel Name']]}, after commit {commit})",
        }
        json.dump(d, open(tj, "w"), indent=2)
    srcp = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join(src, base + "_source.csv")
    open(tmp, "w").write(srcp)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_summary.py"), tmp, "30"],
                         capture_output=True, text=True)
    open(os.path.join(dst, f"ncu_{base}_source_summary.txt"), "w").write(out.stdout + out.stderr[-2000:])
    # stall samples by CUDA source line (needs -lineinfo and --import-source on)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), path, "60"],
                         capture_output=True, text=True)
    open(os.path.join(dst, f"ncu_{base}_source_lines.txt"), "w").write(out.stdout + out.stderr[-2000:])
print(sorted(os.listdir(dst)))
