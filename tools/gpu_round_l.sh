#!/bin/bash
# Round evidence r01_l (after the last kernel changes): bench lines of the three workloads + launch lists of v3 / v4 steps.
mkdir -p gpurun_out
for wl in v2 v3 v4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r01_l_bench_$wl.log 2>&1; echo "bench $wl rc=$?"
  tail -n 1 gpurun_out/r01_l_bench_$wl.log | cut -c1-160
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_l_launches_v4_step.csv python tools/pix_profile.py --batch 64 --steps 2 --reps 1 --no-graph > gpurun_out/ncu2.log 2>&1; echo "v4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_l_launches_v3_step.csv python tools/profile_step.py --v3 --batch 128 --steps 2 --no-decode > gpurun_out/ncu3.log 2>&1; echo "v3 rc=$?"
timeout 300 python tools/pix_profile.py --batch 64 --steps 50 2>&1 | tail -1
timeout 300 python tools/pix_profile.py --batch 256 --steps 20 2>&1 | tail -1
