#!/bin/bash
# N-GPU bench with the comm stream at the highest / default priority (same box).  Usage: gpu_scale_prio.sh tag N
set -u
TAG=${1:-prio}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
for pr in 1 0; do
  GBP_COMM_PRIORITY=$pr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n${N}_prio$pr.json" 2> "$OUT/bench_n${N}_prio$pr.err"
  echo "bench n=$N prio=$pr rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n${N}_prio$pr.json"))
    print("N=$N prio=$pr value %.1f M/s"%(d["value"]/1e6), "e2e %.1f"%(d["e2e"]["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["state_hash"]["means"], {k:(v["count"],round(v["ms"]/max(1,v["count"]),4)) for k,v in d["profile_ms"].items() if v["count"]})
except Exception as e: print("failed", e)
PY
done
