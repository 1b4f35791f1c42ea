#!/bin/bash
# Scaling visit (gpurun --gpus N): bench.py at the listed rank counts, back to back.  Usage: gpu_scale.sh tag "1 2 4 8"
set -u
TAG=${1:-scale}; NS=${2:-"8"}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv > "$OUT/nvidia_smi.csv" 2>&1
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
  fi
  echo "bench n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$n.json"))
    print("N=$n value %.1f M/s"%(d["value"]/1e6), "e2e %.1f"%(d["e2e"]["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["state_hash"]["means"], d["state_hash"]["connectivity"], {k:(v["count"],round(v["ms"]/max(1,v["count"]),4)) for k,v in d["profile_ms"].items() if v["count"]})
except Exception as e: print("N=$n failed", e)
PY
  tail -2 "$OUT/bench_n$n.err"
done
