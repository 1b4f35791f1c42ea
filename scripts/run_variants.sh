#!/bin/bash
# On the GPU box: parity subset + bench for each variant library.  Usage: run_variants.sh tag name...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 600 python -m pytest tests -m gpu -x -q -k "rings_2000 or circle_30 or junction or idle_robots" > $OUT/pytest_$name.log 2>&1
  echo "$name pytest rc=$? $(tail -1 $OUT/pytest_$name.log)"
  timeout 300 python bench.py --steps 10 --no-cpu-baseline > $OUT/bench_$name.json 2>$OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name", "value %.1f M/s"%(d["value"]/1e6), "iterate_ext_int avg %.1f us"%(d["roofline"]["avg_launch_ms"]*1e3), "e2e %.1f M/s"%(d["e2e"]["value"]/1e6))
except Exception as e: print("$name bench failed", e)
PY
done
