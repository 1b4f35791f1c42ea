#!/bin/bash
# Default build: axis + parity tests; then ticks of lattice-1M per variant library (scripts/build_variants.sh).
# Usage: gpu_variants2.sh tag name...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_axis.py tests/test_gpu_parity.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  for wl in lattice rings; do
    timeout 300 python scripts/variant_bench.py --workload $wl --steps 5 --check > $OUT/vb_${name}_$wl.json 2> $OUT/vb_${name}_$wl.err
    echo "$name $wl $(cat $OUT/vb_${name}_$wl.json | cut -c1-400)"
  done
done
