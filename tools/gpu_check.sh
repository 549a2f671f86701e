#!/bin/bash
# One gpurun call: parity tests per precision (separate processes, so a fault in one cannot poison the other),
# smoke, and short benches.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== build" ; python -c "import __graft_entry__ as g; print(g.build())" 2>&1 | tail -2
echo "== pytest fp32"
timeout 900 python -m pytest tests -m gpu -q -k "not bf16 and not full_size" --timeout=600 -p no:cacheprovider > gpurun_out/t_fp32.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/t_fp32.log
echo "== pytest bf16"
timeout 900 python -m pytest tests -m gpu -q -k "bf16 or full_size" --timeout=600 -p no:cacheprovider > gpurun_out/t_bf16.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/t_bf16.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/smoke.log
echo "== bench fp32"
timeout 900 python bench.py --precision fp32 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_fp32.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_fp32.log
echo "== bench bf16"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_bf16.log
