#!/bin/bash
# N-GPU weak-scaling check, launched the way the driver launches bench.py.  Usage: bash tools/gpu_scaleN.sh N
N=${1:-8}
mkdir -p gpurun_out
for wl in v2 v4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --workload $wl > gpurun_out/bench_${wl}_n$N.log 2>&1; echo "$wl rc=$?"
  tail -n 1 gpurun_out/bench_${wl}_n$N.log | cut -c1-330
done
