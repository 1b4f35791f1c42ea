#!/bin/bash
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for name in "$@"; do
  GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload dense --steps 4 --check > $OUT/vb_${name}_dense.json 2> $OUT/vb_${name}_dense.err
  echo "$name dense $(cut -c1-330 $OUT/vb_${name}_dense.json)"
done
