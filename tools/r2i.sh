#!/bin/bash
# 2-GPU pass: C4 split alone (small then full), then the default bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --workload C4 --c4-height 2048 --steps 3 > gpurun_out/r2i_c4small.json 2> gpurun_out/r2i_c4small.err; tail -3 gpurun_out/r2i_c4small.err; cut -c1-900 gpurun_out/r2i_c4small.json
timeout 300 $TR bench.py --gpus 2 --workload C4 --steps 5 > gpurun_out/r2i_c4.json 2> gpurun_out/r2i_c4.err; tail -3 gpurun_out/r2i_c4.err; cut -c1-1200 gpurun_out/r2i_c4.json
( time timeout 600 $TR bench.py --gpus 2 --steps 5 > gpurun_out/r2i_default2.json 2> gpurun_out/r2i_default2.err ) 2>&1 | tail -3; tail -3 gpurun_out/r2i_default2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2i_default2.json"))
    print("value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"], "n", d["n_gpus"])
    print("e2e", d["e2e"])
    for k,v in (d.get("secondary") or {}).items(): print(" ", k, v)
except Exception as e:
    print("default bench FAILED", e)
PY
