#!/bin/bash
# A/B of run-time toggles (GBP_PDL) x variant libraries, 1 M and 125 k lattice.  Usage: gpu_variants5.sh tag name...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_axis.py tests/test_gpu_parity.py tests/test_gpu_shards.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
for name in "$@"; do
  for pdl in 0 1; do
    for n in 1000000 125000; do
      GBP_PDL=$pdl GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload lattice --robots $n --steps 8 > $OUT/vb_${name}_pdl${pdl}_$n.json 2> $OUT/vb_${name}_pdl${pdl}_$n.err
      echo "$name pdl=$pdl n=$n $(cat $OUT/vb_${name}_pdl${pdl}_$n.json | cut -c1-420)"
    done
  done
done
