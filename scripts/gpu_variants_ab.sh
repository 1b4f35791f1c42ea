#!/bin/bash
# Same-box A/B of engine builds: lattice and rings, each variant twice (interleaved).  Usage: gpu_variants_ab.sh tag name...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_axis.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -3 "$OUT/pytest_gpu.log"
for rep in 1 2; do
  for name in "$@"; do
    for wl in lattice rings; do
      GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload $wl --steps 6 --check > $OUT/vb_${name}_${wl}_$rep.json 2> $OUT/vb_${name}_${wl}_$rep.err
      echo "$name $wl $rep $(cut -c1-330 $OUT/vb_${name}_${wl}_$rep.json)"
    done
  done
done
