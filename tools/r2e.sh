#!/bin/bash
mkdir -p gpurun_out
python tools/pcie_bw.py
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -k "one or layout or random or corrupt or golden" 2>&1 | tail -4
run() {  # tag, args...
  local tag=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary "$@" > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err || tail -5 gpurun_out/r2e_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2e_$tag.json"))
    print("$tag", "value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "r", d["compressed_ratio"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2x64_one --workload C2x64 --flags 0x21
run c5_one --workload C5 --flags 0x21
run c3_one --workload C3 --flags 0x21
run c2a_one --workload C2Ax64 --flags 0x21
SMALL='not offsets_beyond and not c3_batch and not full_size and not c2_4k and not alpha_variants and not c5_ and not c4_ and not submit_wait and not pageable'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$SMALL" > gpurun_out/r2e_memcheck.log 2>&1; echo "memcheck exit $?"; tail -12 gpurun_out/r2e_memcheck.log
