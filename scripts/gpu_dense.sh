#!/bin/bash
# One-GPU visit: full -m gpu suite, then an ncu --set full capture of the general kernel k_iterate<1,1> on the dense
# workload (active InterRobot factors).  Usage: gpu_dense.sh tag
set -u
TAG=${1:-dense}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -8 "$OUT/pytest_gpu.log"
timeout 600 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/bench_dense.json" 2> "$OUT/bench_dense.err"; echo "bench dense rc=$?"
cut -c1-1500 "$OUT/bench_dense.json"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_iterate<" -s 30 -c 1 -o "$OUT/prof_general_dense" -f \
  python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/ncu_full_dense.log" 2>&1; echo "ncu full dense rc=$?"
tail -3 "$OUT/ncu_full_dense.log"
ls -la "$OUT"
