#!/bin/bash
# Dense workload: ncu launch list + full capture of k_edge_messages and k_iterate.  Usage: gpu_dense3.sh tag
set -u
TAG=${1:-dense3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file "$OUT/launches.csv" \
  python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_iterate$|^k_edge_messages$' -s 60 -c 2 -o "$OUT/prof_general_dense" -f \
  python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/ncu_full_dense.log" 2>&1; echo "ncu full dense rc=$?"
tail -3 "$OUT/ncu_full_dense.log" | cut -c1-200
ls -la "$OUT"
