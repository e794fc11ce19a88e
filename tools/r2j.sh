#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/r2j_tests.log 2>&1; tail -4 gpurun_out/r2j_tests.log
run() {  # tag, env, args...
  local tag=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary "$@" > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err || tail -5 gpurun_out/r2j_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64_tma A=1 --workload C2x64
run c2x64_notma FLIC_NO_TMA_LOAD=1 --workload C2x64
run c3_tma A=1 --workload C3
run c3_notma FLIC_NO_TMA_LOAD=1 --workload C3
run c5_tma A=1 --workload C5
run c5_notma FLIC_NO_TMA_LOAD=1 --workload C5
run c2a_tma A=1 --workload C2Ax64
run c2a_notma FLIC_NO_TMA_LOAD=1 --workload C2Ax64
