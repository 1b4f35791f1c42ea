#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL transport parity, then the bench at N GPUs (and 1 for reference).
set -u
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv > "$OUT/nvidia_smi.csv" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1; echo "build rc=$?"
if [ "${4:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_nccl.py -x -q > "$OUT/pytest_nccl.log" 2>&1; echo "pytest nccl rc=$?"
  tail -30 "$OUT/pytest_nccl.log"
fi
for n in $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"; echo "bench n=$n rc=$?"
  cat "$OUT/bench_n$n.json"; tail -5 "$OUT/bench_n$n.err"
done
if [ "${3:-1}" = "1" ]; then timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench n=1 rc=$?"; cat "$OUT/bench_n1.json"; fi
