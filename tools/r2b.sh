#!/bin/bash
mkdir -p gpurun_out
python tools/dbg1.py > gpurun_out/r2b_dbg.log 2>&1; tail -60 gpurun_out/r2b_dbg.log
for w in C2x8 C5 C3; do python tools/phase.py $w 2>&1 | tail -10; done
python tools/phase.py C2x8 0x41 2>&1 | tail -10
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_encode -c 1 -o gpurun_out/r2b_enc python bench.py --workload C2x8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_ncu.log 2>&1; tail -3 gpurun_out/r2b_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_decode_one -c 1 -o gpurun_out/r2b_one python bench.py --workload C2x8 --flags 0x21 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_ncu2.log 2>&1; tail -3 gpurun_out/r2b_ncu2.log
