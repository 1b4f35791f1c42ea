#!/bin/bash
# Two-GPU visit: the whole -m gpu suite (NCCL tests included), bench at N = 1 and N = 2.  Usage: gpu_pair.sh tag
set -u
TAG=${1:-pair}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"
tail -8 "$OUT/pytest_gpu.log"
summ() {
python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$1", "value %.1f M/s"%(d["value"]/1e6), "e2e %.1f"%(d["e2e"]["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["state_hash"]["means"], {k:(v["count"],round(v["ms"]/max(1,v["count"]),4)) for k,v in d["profile_ms"].items() if v["count"]})
except Exception as e: print("failed", e)
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench n=1 rc=$?"; summ "$OUT/bench_n1.json"
timeout 600 python bench.py --workload rings --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_rings.json" 2> "$OUT/bench_rings.err"; echo "bench rings rc=$?"; summ "$OUT/bench_rings.json"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"; echo "bench n=2 rc=$?"; summ "$OUT/bench_n2.json"

