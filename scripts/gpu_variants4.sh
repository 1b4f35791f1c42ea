#!/bin/bash
# ticks of lattice-1M per variant library + one smaller lattice on the default build.  Usage: gpu_variants4.sh tag name...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 300 python scripts/variant_bench.py --workload lattice --steps 5 --check > $OUT/vb_${name}_lattice.json 2> $OUT/vb_${name}_lattice.err
  echo "$name lattice $(cat $OUT/vb_${name}_lattice.json | cut -c1-420)"
done
unset GBP_B200_LIB
for n in 500000 125000; do
  timeout 300 python scripts/variant_bench.py --workload lattice --robots $n --steps 10 > $OUT/vb_default_$n.json 2> $OUT/vb_default_$n.err
  echo "default $n $(cat $OUT/vb_default_$n.json | cut -c1-420)"
done
