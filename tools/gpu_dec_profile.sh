#!/bin/bash
# decoder: launch list (gpu__time_duration) of one B=256 decode in bf16
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --steps 1 > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dec.csv python tools/profile_step.py --steps 1 > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu.log
