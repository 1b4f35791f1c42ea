#!/bin/bash
# One GPU visit: full parity suite on the in-tree build, then tick timings of every variant library
# under gpurun_variants/ (scripts/build_variants.sh) with scripts/variant_bench.py.
# Usage: bash scripts/gpu_variants_round.sh <tag> [variant ...]
set -u
TAG=${1:-variants}; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -2 "$OUT/smoke.log"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest_gpu.log"
for wl in lattice rings; do
  timeout 300 python scripts/variant_bench.py --workload $wl 2>>"$OUT/err.log" | tee -a "$OUT/variants.jsonl"
done
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 300 python scripts/variant_bench.py --workload lattice --check 2>>"$OUT/err.log" | tee -a "$OUT/variants.jsonl"
  timeout 300 python scripts/variant_bench.py --workload rings 2>>"$OUT/err.log" | tee -a "$OUT/variants.jsonl"
done
tail -5 "$OUT/err.log" 2>/dev/null
