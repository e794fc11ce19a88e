#!/bin/bash
# usage (on the GPU box): tools/quick.sh TAG [workloads...]   -> parity tests + short device-only bench lines
TAG=$1; shift
WL=${@:-C2x64}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in $WL; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err || tail -5 gpurun_out/${TAG}_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$w.json"))
print("$w", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "r", d["compressed_ratio"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
done
