#!/bin/bash
# Full -m gpu suite on the in-tree build, then interleaved same-box A/B of engine builds on lattice and rings.
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -3 "$OUT/pytest_gpu.log"
for rep in 1 2; do
  for name in "$@"; do
    for wl in lattice rings; do
      GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload $wl --steps 6 > $OUT/vb_${name}_${wl}_$rep.json 2> $OUT/vb_${name}_${wl}_$rep.err
      echo "$name $wl $rep $(cut -c1-300 $OUT/vb_${name}_${wl}_$rep.json)"
    done
  done
done
