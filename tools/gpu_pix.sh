#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pix.py -q -m gpu --timeout=300 -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert [0-9]|rror" | head -12
for h in 1 3; do
  echo "== LDM_PIX_HALO=$h"
  LDM_PIX_HALO=$h timeout 300 python tools/pix_profile.py --batch 64 --steps 50 2>&1 | tail -1
  LDM_PIX_HALO=$h timeout 300 python tools/pix_profile.py --batch 256 --steps 20 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pix.csv python tools/pix_profile.py --batch 64 --steps 2 --reps 1 --no-graph > gpurun_out/ncu_pix.log 2>&1; echo "ncu rc=$?"
