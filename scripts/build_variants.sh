#!/bin/bash
# Build tuning variants of the engine into gpurun_variants/ (git-ignored, travels to the GPU box).
# Usage: scripts/build_variants.sh name1:"-DFLAG ..." name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  echo "building $name with [$flags]"
  (cd magics_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
     -Xcompiler -fPIC -shared $flags -o ../../gpurun_variants/libgbp_$name.so gbp_engine.cu gbp_collide_host.cpp) &
done
wait
ls -la gpurun_variants
