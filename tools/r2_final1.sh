#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err; tail -2 gpurun_out/r02_bench_default_1gpu.err
python bench.py --impl reference --steps 2 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for w in C3 C5 C2Ax64 C2 C1; do python bench.py --workload $w --no-e2e --no-cpu --no-secondary > gpurun_out/r02_bench_${w}_device.json 2>/dev/null; done
python bench.py --workload C2x64 --encoder fused --no-e2e --no-cpu --no-secondary > gpurun_out/r02_bench_c2x64_fused_device.json 2>/dev/null
python tools/pcie_bw.py > gpurun_out/r02_link_probe.txt; python tools/pcie_bw2.py >> gpurun_out/r02_link_probe.txt; python tools/pcie_bw3.py >> gpurun_out/r02_link_probe.txt; python tools/e2e_probe.py >> gpurun_out/r02_link_probe.txt
python tools/phase.py C2x8 > gpurun_out/r02_phase_clocks.txt 2>&1; python tools/phase.py C2x8 0x41 >> gpurun_out/r02_phase_clocks.txt 2>&1; python tools/phase.py C2x8 0x21 >> gpurun_out/r02_phase_clocks.txt 2>&1; python tools/phase.py C5x8 0x21 >> gpurun_out/r02_phase_clocks.txt 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_default_1gpu.json"))
print("value", d["value"], "enc", d["encode_GBps"], "dec", d["decode_GBps"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["decode_path_frac"])
print("e2e", d["e2e"]["value"], "cpu_model", d["cpu_model"]["value"])
for k,v in d["secondary"].items(): print(" ", k, v["value_GBps"], v.get("encode_path_frac"), v.get("decode_path_frac"))
PY
