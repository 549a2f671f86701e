#!/bin/bash
# Run the GPU suite several times (different orders) and keep every log: hunts order- or timing-dependent failures.
mkdir -p gpurun_out
for i in 1 2 3 4; do
  if [ $((i % 2)) = 0 ]; then ORDER="tests/test_v3.py tests/test_ublock.py tests/test_gpu_parity.py tests/test_pix.py"; else ORDER="tests"; fi
  timeout 900 python -m pytest $ORDER -q -m gpu --timeout=600 -p no:cacheprovider --tb=short > gpurun_out/flake_$i.log 2>&1
  echo "run $i rc=$? $(tail -1 gpurun_out/flake_$i.log)"
done
