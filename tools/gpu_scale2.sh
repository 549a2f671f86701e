#!/bin/bash
# 2-GPU weak-scaling check of both workloads, launched the way the driver launches bench.py
mkdir -p gpurun_out
for wl in v2 v4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --workload $wl > gpurun_out/bench_${wl}_n2.log 2>&1; echo "$wl rc=$?"
  tail -n 1 gpurun_out/bench_${wl}_n2.log | cut -c1-330
done
