#!/bin/bash
# parity tests (with a per-test timeout) + short device-only bench lines: tools/quick2.sh TAG [workload ...]
TAG=$1; shift
WL=${@:-C2x64}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
for w in $WL; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err || tail -5 gpurun_out/${TAG}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$w.json"))
    print("$w", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$w FAILED", e)
PY
done
