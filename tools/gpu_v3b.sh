#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v3.py tests/test_gpu_parity.py tests/test_ublock.py -q -m gpu --timeout=600 -p no:cacheprovider --tb=short > gpurun_out/t_v3b.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_v3b.log
for k in 0 1; do
  echo "== LDM_GEMM_SPLITK=$k"
  LDM_GEMM_SPLITK=$k timeout 600 python bench.py --workload v3 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v3', d['value'], d['ms_per_step'], d['gpu_launches'])"
  LDM_GEMM_SPLITK=$k LDM_CHAIN=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2 per-layer', d['value'], d['ms_per_step'])"
done
