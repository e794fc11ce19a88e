#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -k "host_pipeline or submit or pageable or batch or one or layout or random or errors or mixed" 2>&1 | tail -4
python tools/e2e_probe.py 2>&1 | tail -9
run() {  # tag, args...
  local tag=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary "$@" > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err || tail -5 gpurun_out/r2g_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "r", d["compressed_ratio"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64_one --workload C2x64 --flags 0x21
run c5_one --workload C5 --flags 0x21
run c3_one --workload C3 --flags 0x21
python tools/phase.py C2x8 0x21 2>&1 | tail -9
python tools/phase.py C5x8 0x21 2>&1 | tail -9
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2g_e2e.json 2> gpurun_out/r2g_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_e2e.json')); print('e2e', d['e2e'])"
