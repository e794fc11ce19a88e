#!/bin/bash
# round 2, second session: ncu evidence of the final build.  The .ncu-rep files (20 MB each with source) are digested on
# the box — gpurun_out/ carries at most 64 MiB back — and only the text summaries and ONE report are kept.
mkdir -p gpurun_out /tmp/rep
B="python bench.py --warmup 3 --no-e2e --no-cpu --no-secondary"
$B --steps 2 --workload C2x64 > gpurun_out/prof_plain_c2x64.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_ncu_launches_c2x64.csv $B --steps 2 --workload C2x64 > gpurun_out/prof_ncu1.log 2>&1; tail -1 gpurun_out/prof_ncu1.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:"k_histograms|k_tables|k_slots|k_pack|k_finalize|k_decode" -s 18 -c 6 -f -o /tmp/rep/r02b_full_c2x64 $B --steps 1 --workload C2x64 > gpurun_out/prof_ncu2.log 2>&1; tail -1 gpurun_out/prof_ncu2.log | cut -c1-200
python tools/ncu_summary.py /tmp/rep/r02b_full_c2x64.ncu-rep > gpurun_out/r02b_ncu_full_c2x64.txt
RAW=$(python -c "import json;d=json.load(open('gpurun_out/prof_plain_c2x64.json'));print(d['config']['raw_bytes_per_gpu_step'], d['compressed_ratio'])")
python tools/ncu_traffic.py /tmp/rep/r02b_full_c2x64.ncu-rep C2x64 $RAW > gpurun_out/r02b_traffic_c2x64.json
python tools/ncu_lsu.py /tmp/rep/r02b_full_c2x64.ncu-rep > gpurun_out/r02b_ncu_lsu_c2x64.txt
cp /tmp/rep/r02b_full_c2x64.ncu-rep gpurun_out/
$B --steps 1 --workload C3 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"k_histograms|k_tables|k_slots|k_pack|k_finalize|k_decode" -s 18 -c 6 -f -o /tmp/rep/r02b_full_c3 $B --steps 1 --workload C3 > gpurun_out/prof_ncu3.log 2>&1; tail -1 gpurun_out/prof_ncu3.log | cut -c1-200
python tools/ncu_summary.py /tmp/rep/r02b_full_c3.ncu-rep > gpurun_out/r02b_ncu_full_c3.txt
python tools/ncu_lsu.py /tmp/rep/r02b_full_c3.ncu-rep > gpurun_out/r02b_ncu_lsu_c3.txt
$B --steps 1 --workload C2x8 --flags 0x41 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"k_pack" -s 3 -c 1 -f -o /tmp/rep/r02b_full_exact $B --steps 1 --workload C2x8 --flags 0x41 > gpurun_out/prof_ncu4.log 2>&1
python tools/ncu_summary.py /tmp/rep/r02b_full_exact.ncu-rep > gpurun_out/r02b_ncu_full_pack_exact_c2x8.txt
$B --steps 1 --workload C2x8 --flags 0x21 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"k_pack|k_decode_one" -s 6 -c 2 -f -o /tmp/rep/r02b_full_one $B --steps 1 --workload C2x8 --flags 0x21 > gpurun_out/prof_ncu5.log 2>&1
python tools/ncu_summary.py /tmp/rep/r02b_full_one.ncu-rep > gpurun_out/r02b_ncu_full_one_stream_c2x8.txt
du -sh gpurun_out; ls -la gpurun_out | tail -20
