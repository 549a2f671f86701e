#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --steps 3 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log
timeout 600 python -m pytest tests -m gpu -q -k "bf16" --durations=8 --timeout=600 -p no:cacheprovider 2>&1 | tail -15
