#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v3.py -q -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert|rror" | head -12
for a in 0 1; do
  echo "== LDM_ATTN_TC=$a"
  LDM_ATTN_TC=$a timeout 600 python bench.py --workload v3 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v3', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv python tools/profile_step.py --v3 --batch 128 --steps 2 --no-decode > gpurun_out/ncu.log 2>&1; echo "rc=$?"
grep -E "attn" gpurun_out/launches_v3.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | tail -10
