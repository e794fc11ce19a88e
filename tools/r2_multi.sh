#!/bin/bash
# usage: tools/r2_multi.sh N [TAG]   -> default bench.py under torchrun on N GPUs of one box
N=$1; TAG=${2:-r02b}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/${TAG}_bench_default_${N}gpu.raw 2> gpurun_out/${TAG}_bench_default_${N}gpu.err; echo "exit $?"; tail -2 gpurun_out/${TAG}_bench_default_${N}gpu.err | cut -c1-200
grep '^{' gpurun_out/${TAG}_bench_default_${N}gpu.raw > gpurun_out/${TAG}_bench_default_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_default_${N}gpu.json"))
print("N", d["n_gpus"], "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"],2), d["config"]["host_numa"])
for k,v in d["secondary"].items(): print(" ", k, v["value_GBps"], v.get("encode_GBps"), v.get("decode_GBps"), v.get("scaling"), v.get("transport"), v.get("peer_memory_unavailable"))
PY
nvidia-smi topo -m 2>/dev/null | head -12; nproc
