#!/bin/bash
mkdir -p gpurun_out
for r in 0 1; do
  LDM_HALO_TWO_CTAS=$r timeout 300 python -m pytest tests/test_pix.py -q -m gpu --timeout=300 -p no:cacheprovider 2>&1 | tail -1
  LDM_HALO_TWO_CTAS=$r timeout 300 python tools/pix_profile.py --batch 64 --steps 50 2>&1 | tail -1
  LDM_HALO_TWO_CTAS=$r timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/halo_two_$r.csv python tools/pix_profile.py --batch 64 --steps 1 --reps 1 --no-graph > gpurun_out/ncu_dbg.log 2>&1
  echo "two=$r: $(grep conv_halo gpurun_out/halo_two_$r.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
