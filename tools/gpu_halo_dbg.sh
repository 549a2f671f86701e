#!/bin/bash
mkdir -p gpurun_out
for d in 0 32 8 16 24 56; do
  LDM_HALO_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/halo_dbg_$d.csv python tools/pix_profile.py --batch 64 --steps 1 --reps 1 --no-graph > gpurun_out/ncu_dbg.log 2>&1
  echo "debug=$d: $(grep conv_halo gpurun_out/halo_dbg_$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
