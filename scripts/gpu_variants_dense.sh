#!/bin/bash
# Same-box A/B of engine builds on the dense workload (general kernel k_iterate) and, for safety, the lattice.
# Usage: gpu_variants_dense.sh tag name...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for name in "$@"; do
  for wl in dense lattice; do
    chk=""; [ "$wl" = "dense" ] && chk="--check"
    GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload $wl --steps 4 $chk > $OUT/vb_${name}_$wl.json 2> $OUT/vb_${name}_$wl.err
    echo "$name $wl $(cut -c1-600 $OUT/vb_${name}_$wl.json)"
  done
done
GBP_ITER_GRID_PER_SM=4 GBP_B200_LIB=$PWD/gpurun_variants/libgbp_base.so timeout 300 python scripts/variant_bench.py --workload dense --steps 4 > $OUT/vb_base_grid4_dense.json 2> $OUT/vb_base_grid4_dense.err
echo "base grid4 dense $(cut -c1-600 $OUT/vb_base_grid4_dense.json)"
