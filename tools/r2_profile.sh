#!/bin/bash
# Final evidence pass (one gpurun call; all ncu runs count as one): plain runs first (must exit 0), then
#  1. the launch list of the default bench step (gpu__time_duration.sum per launch),
#  2. ncu --set full of every kernel of one step: staged path on C2x64 (traffic per launch) and C3 (RGB variants),
#     the fused encoder and the one-stream decoder on C2x8.
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-e2e --no-cpu --no-secondary"
$B --steps 2 --workload C2x64 > gpurun_out/prof_plain_c2x64.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_c2x64.csv $B --steps 2 --workload C2x64 > gpurun_out/prof_ncu1.log 2>&1; tail -1 gpurun_out/prof_ncu1.log | cut -c1-200
# staged path: warm-up = 3 steps x 6 kernels + 1 copy/flush kernels? select by name, skip the warm-up launches of each
ncu --set full --import-source on --clock-control none -k regex:"k_histograms|k_tables|k_slots|k_pack|k_finalize|k_decode" -s 18 -c 6 -o gpurun_out/r02_full_c2x64 $B --steps 1 --workload C2x64 > gpurun_out/prof_ncu2.log 2>&1; tail -1 gpurun_out/prof_ncu2.log | cut -c1-200
$B --steps 1 --workload C3 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"k_histograms|k_tables|k_slots|k_pack|k_finalize|k_decode" -s 18 -c 6 -o gpurun_out/r02_full_c3 $B --steps 1 --workload C3 > gpurun_out/prof_ncu3.log 2>&1; tail -1 gpurun_out/prof_ncu3.log | cut -c1-200
$B --steps 1 --workload C2x8 --encoder fused --flags 0x21 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"k_encode|k_decode_one" -s 6 -c 2 -o gpurun_out/r02_full_fused_one_c2x8 $B --steps 1 --workload C2x8 --encoder fused --flags 0x21 > gpurun_out/prof_ncu4.log 2>&1; tail -1 gpurun_out/prof_ncu4.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -5
