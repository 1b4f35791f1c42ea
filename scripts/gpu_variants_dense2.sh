#!/bin/bash
# Full -m gpu suite on the in-tree build, then same-box A/B of engine builds on the dense workload and the lattice.
# Usage: gpu_variants_dense2.sh tag name...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -12 "$OUT/pytest_gpu.log"
for name in "$@"; do
  for wl in dense lattice; do
    chk=""; [ "$wl" = "dense" ] && chk="--check"
    GBP_B200_LIB=$PWD/gpurun_variants/libgbp_$name.so timeout 300 python scripts/variant_bench.py --workload $wl --steps 4 $chk > $OUT/vb_${name}_$wl.json 2> $OUT/vb_${name}_$wl.err
    echo "$name $wl $(cut -c1-560 $OUT/vb_${name}_$wl.json)"
  done
done
