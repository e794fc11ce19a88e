#!/bin/bash
mkdir -p gpurun_out
python tools/phase.py C2x8 0x21 2>&1 | tail -22
python tools/phase.py C5x8 0x21 2>&1 | tail -11
e2e() {  # tag, chunk bytes
  FLIC_CHUNK_BYTES=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2f_e2e_$1.json 2> gpurun_out/r2f_e2e_$1.err || tail -3 gpurun_out/r2f_e2e_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r2f_e2e_$1.json')); print('chunk $2', 'e2e', round(d['e2e']['value'],2), 'per-dir', d['e2e']['pcie_GBps_per_direction_per_gpu'])"
}
e2e 33m 34000000
e2e 67m 67000000
e2e 134m 134000000
e2e 268m 268000000
e2e 540m 540000000
