#!/bin/bash
# round-2 first GPU pass: parity tests, then short device-only bench lines for both encoders and the layouts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
tail -40 gpurun_out/r2a_tests.log
run() {  # tag, args...
  local tag=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/r2a_$tag.json 2> gpurun_out/r2a_$tag.err || tail -5 gpurun_out/r2a_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "r", d["compressed_ratio"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64_fused --workload C2x64
run c2x64_staged --workload C2x64 --encoder staged
run c2x64_exact --workload C2x64 --flags 0x41
run c2x64_one --workload C2x64 --flags 0x21
run c5_fused --workload C5
run c5_staged --workload C5 --encoder staged
run c5_one --workload C5 --flags 0x21
run c3_fused --workload C3
run c3_staged --workload C3 --encoder staged
run c2a_fused --workload C2Ax64
