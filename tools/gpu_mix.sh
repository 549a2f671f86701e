#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v3.py tests/test_gpu_parity.py tests/test_pix.py -q -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert|rror" | head -12
timeout 600 python bench.py --workload v3 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v3', d['value'], d['ms_per_step'], d['gpu_launches'])"
for t in 0 1; do
  LDM_PIX_IN_TC=$t timeout 300 python tools/pix_profile.py --batch 64 --steps 50 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pix.csv python tools/pix_profile.py --batch 64 --steps 2 --reps 1 --no-graph > gpurun_out/ncu_pix.log 2>&1; echo "ncu rc=$?"
grep -E "conv_in" gpurun_out/launches_pix.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | tail -2
