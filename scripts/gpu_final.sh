#!/bin/bash
# Final check of a round: the whole -m gpu suite, then bench.py exactly as the driver calls it (default flags) and the
# reference arm.  Usage: gpu_final.sh tag
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -1 "$OUT/smoke.log"
if [ "${2:-}" != "skip-tests" ]; then timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; fi
tail -6 "$OUT/pytest_gpu.log"
SECONDS=0; timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench default rc=$?"
echo "wall ${SECONDS}s"
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
print("value %.1f M/s e2e %.1f M/s ms/step %.2f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), d["clocks"], "frac %.3f traffic_frac %.3f"%(d["roofline"]["frac"], d["roofline"]["traffic_frac"]), d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["gpu_launches"])
PY
SECONDS=0; timeout 900 python bench.py --impl reference > "$OUT/bench_reference_default.json" 2> "$OUT/bench_reference_default.err"; echo "bench reference rc=$?"
echo "wall ${SECONDS}s"; cut -c1-300 "$OUT/bench_reference_default.json"
