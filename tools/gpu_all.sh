#!/bin/bash
# Whole GPU suite as the driver runs it, smoke, and the three bench workloads.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/t_all.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/smoke.log
echo "== bench v2"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v2.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/bench_v2.log | cut -c1-700
echo "== bench v4"
timeout 900 python bench.py --workload v4 --steps 2 --warmup 3 > gpurun_out/bench_v4.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/bench_v4.log | cut -c1-1500
