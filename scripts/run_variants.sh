#!/bin/bash
# On the GPU box: parity subset + bench (rings-100k and lattice-1M) for each variant library.
# Usage: run_variants.sh tag name...   (libraries built by scripts/build_variants.sh)
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 600 python -m pytest tests -m gpu -x -q -k "rings_2000 or circle_30_ticks or junction_twoway_all_factor_kinds or idle_robots or lattice_slabs" > $OUT/pytest_$name.log 2>&1
  echo "$name pytest rc=$? $(tail -1 $OUT/pytest_$name.log)"
  for wl in rings lattice; do
    timeout 300 python bench.py --workload $wl --steps 10 --no-cpu-baseline > $OUT/bench_${name}_$wl.json 2>$OUT/bench_${name}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${name}_$wl.json"))
    print("$name $wl", "value %.1f M/s"%(d["value"]/1e6), "iterate_ext_int avg %.1f us"%(d["roofline"]["avg_launch_ms"]*1e3), "e2e %.1f M/s"%(d["e2e"]["value"]/1e6))
except Exception as e: print("$name $wl bench failed", e)
PY
  done
done
