#!/bin/bash
# Selected tests on the default build, then ticks of lattice-1M per variant library.  Usage: gpu_variants3.sh tag "pytest args" name...
TAG=$1; PYT=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest $PYT -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 300 python scripts/variant_bench.py --workload lattice --steps 5 --check > $OUT/vb_${name}_lattice.json 2> $OUT/vb_${name}_lattice.err
  echo "$name lattice $(cat $OUT/vb_${name}_lattice.json | cut -c1-420)"
done
