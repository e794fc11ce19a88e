#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1
tail -15 gpurun_out/r2c_tests.log
python tools/dbg1.py > gpurun_out/r2c_dbg.log 2>&1; tail -5 gpurun_out/r2c_dbg.log
for w in C2x8 C5; do python tools/phase.py $w 2>&1 | tail -10; done
run() {  # tag, args...
  local tag=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary "$@" > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err || tail -5 gpurun_out/r2c_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "r", d["compressed_ratio"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64_fused --workload C2x64 --encoder fused
run c2x64_exact --workload C2x64 --flags 0x41
run c2x64_one --workload C2x64 --flags 0x21
run c5_one --workload C5 --flags 0x21
run c3_one --workload C3 --flags 0x21
run c1_fused --workload C1 --encoder fused
run c1_staged --workload C1 --encoder staged
run c2_fused --workload C2 --encoder fused
run c2_staged --workload C2 --encoder staged
run c2x8_fused --workload C2x8 --encoder fused
run c2x8_staged --workload C2x8 --encoder staged
python __graft_entry__.py smoke 2>&1 | tail -3
