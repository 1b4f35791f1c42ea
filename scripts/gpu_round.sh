#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the iterate kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvidia_smi.csv" 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
cat "$OUT/bench.json"
timeout 600 python bench.py --workload rings --no-cpu-baseline > "$OUT/bench_rings100k.json" 2> "$OUT/bench_rings100k.err"; echo "bench rings rc=$?"
cat "$OUT/bench_rings100k.json"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench ref rc=$?"
cat "$OUT/bench_reference.json"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 40 -c 2 -o "$OUT/prof_iterate" -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 40 -c 1 -o "$OUT/prof_iterate_rings" -f \
  python bench.py --workload rings --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_rings.log" 2>&1; echo "ncu full rings rc=$?"
ls -la "$OUT"
