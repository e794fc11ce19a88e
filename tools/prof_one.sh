#!/bin/bash
# ncu --set full (+ source counters) of one step's kernels on a small workload: tools/prof_one.sh TAG [WORKLOAD] [KERNEL_REGEX] [extra bench flags]
TAG=$1; W=${2:-C2x8}; K=${3:-"k_histograms|k_tables|k_pack|k_decode"}; shift 3
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-e2e --no-cpu --no-secondary --workload $W $@"
$B --steps 1 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
N=$(echo "$K" | tr '|' '\n' | wc -l)
ncu --set full --import-source on --clock-control none -k regex:"$K" -s $((3 * N)) -c $N -f -o gpurun_out/${TAG} $B --steps 1 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ls -la gpurun_out/${TAG}.ncu-rep
