#!/bin/bash
# one ncu --set full capture of the persistent chain kernel (50 reverse steps, B = 256), source-level
mkdir -p gpurun_out
export LDM_CHAIN=1
timeout 300 python tools/profile_step.py --steps 50 --no-decode > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/chain_full -f python tools/profile_step.py --steps 50 --no-decode > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu.log; ls -la gpurun_out/chain_full.ncu-rep
