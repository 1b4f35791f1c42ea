#!/bin/bash
# N-GPU bench with the halo overlap on and off (same box).  Usage: gpu_scale_ab.sh tag N
set -u
TAG=${1:-scaleab}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
for ov in 1 0; do
  GBP_HALO_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n${N}_ov$ov.json" 2> "$OUT/bench_n${N}_ov$ov.err"
  echo "bench n=$N overlap=$ov rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n${N}_ov$ov.json"))
    print("N=$N ov=$ov value %.1f M/s"%(d["value"]/1e6), "e2e %.1f"%(d["e2e"]["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["state_hash"]["means"], {k:(v["count"],round(v["ms"]/max(1,v["count"]),4)) for k,v in d["profile_ms"].items() if v["count"]})
except Exception as e: print("failed", e)
PY
done
