#!/bin/bash
# ncu --set full capture of k_iterate<1,1> for variant libraries: ncu_variant.sh tag workload name...
TAG=$1; WL=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for name in "$@"; do
  export GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 40 -c 1 -o "$OUT/prof_$name" -f \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_$name.log" 2>&1; echo "ncu $name rc=$?"
done
ls -la $OUT
