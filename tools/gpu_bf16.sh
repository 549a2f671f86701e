#!/bin/bash
mkdir -p gpurun_out
echo "== pytest bf16"
timeout 900 python -m pytest tests -m gpu -q -k "bf16 or full_size" --timeout=600 -p no:cacheprovider > gpurun_out/t_bf16.log 2>&1; echo "rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|assert .* <|^E  " gpurun_out/t_bf16.log | head -60
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/smoke.log
echo "== bench bf16"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_bf16.log
