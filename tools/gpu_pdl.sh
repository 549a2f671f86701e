#!/bin/bash
mkdir -p gpurun_out
echo "== PDL on: parity"
LDM_PDL=1 timeout 900 python -m pytest tests/test_pix.py tests/test_gpu_parity.py -q -m gpu --timeout=600 -p no:cacheprovider 2>&1 | tail -3
for p in 0 1; do
  echo "== LDM_PDL=$p"
  LDM_PDL=$p timeout 300 python tools/pix_profile.py --batch 64 --steps 50 2>&1 | tail -1
  LDM_PDL=$p timeout 300 python tools/pix_profile.py --batch 256 --steps 20 2>&1 | tail -1
  LDM_PDL=$p timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2', d['value'], d['decode'])"
done
