#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v3.py tests/test_gpu_parity.py -q -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert|rror" | head -12
timeout 600 python bench.py --workload v3 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v3', d['value'], d['ms_per_step'], d['gpu_launches'])"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2', d['value'], d['ms_per_step'], d['gpu_launches'], d['decode'])"
