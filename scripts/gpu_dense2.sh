#!/bin/bash
# ncu --set full of the general kernel k_iterate on the dense workload + the evaluation tests.  Usage: gpu_dense2.sh tag
set -u
TAG=${1:-dense}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_evaluation.py tests/test_gpu_axis.py -m gpu -q > "$OUT/pytest_eval.log" 2>&1; echo "pytest rc=$?"
tail -4 "$OUT/pytest_eval.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_iterate$' -s 30 -c 1 -o "$OUT/prof_general_dense" -f \
  python bench.py --workload dense --steps 2 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/ncu_full_dense.log" 2>&1; echo "ncu full dense rc=$?"
tail -3 "$OUT/ncu_full_dense.log" | cut -c1-300
ls -la "$OUT"
