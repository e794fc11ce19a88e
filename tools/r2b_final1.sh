#!/bin/bash
# round 2, second session: single-GPU bench lines of the final build
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r02b_bench_default_1gpu.json 2> gpurun_out/r02b_bench_default_1gpu.err; tail -2 gpurun_out/r02b_bench_default_1gpu.err
python bench.py --impl reference --steps 2 --no-cpu > gpurun_out/r02b_bench_reference_arm.json 2>/dev/null
for w in C3 C5 C2Ax64 C2 C1; do python bench.py --workload $w --no-e2e --no-cpu --no-secondary > gpurun_out/r02b_bench_${w}_device.json 2>/dev/null; done
python bench.py --workload C2x64 --encoder fused --no-e2e --no-cpu --no-secondary > gpurun_out/r02b_bench_c2x64_fused_device.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02b_bench_default_1gpu.json"))
print("value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["decode_path_frac"])
print("e2e", d["e2e"]["value"], "cpu_model", d["cpu_model"]["value"])
for k,v in d["secondary"].items(): print(" ", k, v["value_GBps"], v.get("encode_GBps"), v.get("decode_GBps"), v.get("encode_path_frac"), v.get("decode_path_frac"))
PY
