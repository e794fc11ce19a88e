#!/bin/bash
# usage: tools/r2_multi.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r02_bench_default_${N}gpu.raw 2> gpurun_out/r02_bench_default_${N}gpu.err; echo "exit $?"; tail -2 gpurun_out/r02_bench_default_${N}gpu.err | cut -c1-200
grep '^{' gpurun_out/r02_bench_default_${N}gpu.raw > gpurun_out/r02_bench_default_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_default_${N}gpu.json"))
print("N", d["n_gpus"], "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"],2), d["config"]["host_numa"])
for k,v in d["secondary"].items(): print(" ", k, v["value_GBps"], v.get("scaling"))
PY
nvidia-smi topo -m 2>/dev/null | head -12; nproc; lscpu | grep -i "numa node" | head -4
