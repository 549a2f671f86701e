#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|pix_conv_in" -c 4 -f -o gpurun_out/pix_full python tools/pix_profile.py --batch 64 --steps 1 --reps 1 --no-graph > gpurun_out/ncu_pix_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/pix_full.ncu-rep
