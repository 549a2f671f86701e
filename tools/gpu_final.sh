#!/bin/bash
# What the driver runs at round end, in its order: GPU suite, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/final_bench_ref.log | cut -c1-160
timeout 900 python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/final_bench.log
