#!/bin/bash
# One gpurun call: the whole GPU suite as the driver runs it, smoke, then the v3 workload (bench line + launch list).
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/t_all.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/smoke.log
echo "== bench v3"
timeout 900 python bench.py --workload v3 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_v3.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/bench_v3.log | cut -c1-900
echo "== v3 launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv python tools/profile_step.py --v3 --batch 128 --steps 2 --no-decode > gpurun_out/ncu.log 2>&1; echo "rc=$?"
