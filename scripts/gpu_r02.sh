#!/bin/bash
# Round-2 GPU visit: parity tests, bench lines, ncu launch list, ncu full capture of k_iterate_axis<1,1>.
# Usage (under gpurun, from the repo root): bash scripts/gpu_r02.sh <tag> [skip-tests]
set -u
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvidia_smi.csv" 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -2 "$OUT/smoke.log"
if [ "${2:-}" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
  tail -8 "$OUT/pytest_gpu.log"
fi
for wl in lattice rings; do
  timeout 600 python bench.py --workload $wl --steps 10 --no-cpu-baseline > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$wl.json"))
    print("$wl", "value %.1f M/s"%(d["value"]/1e6), "e2e %.1f M/s"%(d["e2e"]["value"]/1e6), "ms/step %.2f"%d["ms_per_step"], {k:(v["count"],round(v["ms"],1)) for k,v in d["profile_ms"].items()})
except Exception as e: print("$wl bench failed", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iterate_axis -s 40 -c 2 -o "$OUT/prof_axis" -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iterate_axis -s 40 -c 1 -o "$OUT/prof_axis_rings" -f \
  python bench.py --workload rings --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_rings.log" 2>&1; echo "ncu full rings rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench reference rc=$?"
cut -c1-400 "$OUT/bench_reference.json"
ls -la "$OUT"
