#!/bin/bash
# One-GPU visit for the sharding code path: local-shard parity tests (+ memcheck of one), then the rest.
set -u
TAG=${1:-shards}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_shards.py -x -q > "$OUT/pytest_shards.log" 2>&1; echo "pytest shards rc=$?"
tail -30 "$OUT/pytest_shards.log"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_shards.py -x -q \
  -k "lattice_slabs and 3" > "$OUT/memcheck.log" 2>&1; echo "memcheck rc=$?"
tail -15 "$OUT/memcheck.log"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shards.py > "$OUT/pytest_gpu_rest.log" 2>&1; echo "pytest rest rc=$?"
tail -5 "$OUT/pytest_gpu_rest.log"
timeout 600 python bench.py --steps 5 --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
