#!/bin/bash
# Evidence for the round: plain bench, then the ncu launch list of the SAME command, then --set full captures of the
# dominant kernel (chain_kernel) and of the first decoder convolutions.  Usage: bash tools/gpu_profile_round.sh <tag>
# Afterwards, here: python tools/ncu_traffic.py gpurun_out/<tag>_chain_full.ncu-rep profiles/<tag>_chain_traffic.json
TAG=${1:-rXX}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary"
$CMD > gpurun_out/${TAG}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_bench_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_bench_plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -f -o gpurun_out/${TAG}_chain_full $CMD > gpurun_out/ncu2.log 2>&1
echo "chain full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 2 -f -o gpurun_out/${TAG}_conv_full $CMD > gpurun_out/ncu3.log 2>&1
echo "conv full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"final_gn_conv3|convt_halo|sa_map_gate" -c 5 -f -o gpurun_out/${TAG}_dec_full $CMD > gpurun_out/ncu4.log 2>&1
echo "decoder kernels full rc=$?"
ls -la gpurun_out/*.ncu-rep
