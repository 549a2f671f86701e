#!/bin/bash
# Round evidence r01_k: full GPU suite, smoke, bench lines (v2 / v3 / v4), ncu launch lists and --set full captures.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/t_all.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/smoke.log
for wl in v2 v3 v4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r01_k_bench_$wl.log 2>&1; echo "bench $wl rc=$?"
  tail -n 1 gpurun_out/r01_k_bench_$wl.log | cut -c1-200
done
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r01_k_bench_ref.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/r01_k_bench_ref.log | cut -c1-200
echo "== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_k_launches_bench_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1; echo "v2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_k_launches_v4_step.csv python tools/pix_profile.py --batch 64 --steps 2 --reps 1 --no-graph > gpurun_out/ncu2.log 2>&1; echo "v4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_k_launches_v3_step.csv python tools/profile_step.py --v3 --batch 128 --steps 2 --no-decode > gpurun_out/ncu3.log 2>&1; echo "v3 rc=$?"
echo "== full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_halo_kernel" -c 17 -f -o gpurun_out/r01_k_pix_full python tools/pix_profile.py --batch 64 --steps 1 --reps 1 --no-graph > gpurun_out/ncu4.log 2>&1; echo "pix full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -c 4 -f -o gpurun_out/r01_k_attn_full python tools/profile_step.py --v3 --batch 128 --steps 1 --no-decode > gpurun_out/ncu5.log 2>&1; echo "attn full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
