#!/bin/bash
# Quick one-GPU visit: all -m gpu tests, then the two bench workloads (no CPU baseline).
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -15 "$OUT/pytest_gpu.log"
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
