#!/bin/bash
# GPU visit for the two-kernel iterate path: the new A/B tests first, then the whole suite, then bench lines.
set -u
TAG=${1:-axis}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -3 "$OUT/smoke.log"
timeout 900 python -m pytest tests/test_gpu_axis.py -x -q > "$OUT/pytest_axis.log" 2>&1; echo "pytest axis rc=$?"
tail -25 "$OUT/pytest_axis.log"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_axis.py > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
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
