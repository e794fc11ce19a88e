#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/r2k_tests.log 2>&1; tail -4 gpurun_out/r2k_tests.log
run() {  # tag, env, args...
  local tag=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary "$@" > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err || tail -5 gpurun_out/r2k_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2k_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64 A=1 --workload C2x64
run c3 A=1 --workload C3
run c3_nortma FLIC_NO_TMA_RGB=1 --workload C3
run c5 A=1 --workload C5
run c2a A=1 --workload C2Ax64
for b in 2 4 16; do run c2x64_bpw$b FLIC_TAB_BPW=$b --workload C2x64; done
run c5_bpw4 FLIC_TAB_BPW=4 --workload C5
