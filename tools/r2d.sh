#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
for w in C2x8 C5; do python tools/phase.py $w 2>&1 | tail -10; done
python tools/phase.py C2x8 0x41 2>&1 | tail -10
( time python bench.py --steps 5 > gpurun_out/r2d_default.json 2> gpurun_out/r2d_default.err ) 2>&1 | tail -3; tail -3 gpurun_out/r2d_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2d_default.json"))
    print("value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"])
    print("e2e", d["e2e"])
    for k,v in (d.get("secondary") or {}).items(): print(" ", k, v)
    print("cpu_model", d.get("cpu_model"))
    print("numa", d["config"]["host_numa"], "clocks", d["clocks"])
except Exception as e:
    print("default bench FAILED", e)
PY
python bench.py --impl reference --steps 2 --no-cpu | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_decode_one -c 1 -o gpurun_out/r2d_one python bench.py --workload C2x8 --flags 0x21 --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log | cut -c1-200
nvidia-smi topo -m 2>&1 | head -20; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; free -g | head -2
