python tools/e2e_probe.py 2>&1 | tail -9
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider --timeout=120 --timeout-method=thread -k "host_pipeline or submit or pageable or batch or mixed or errors" 2>&1 | tail -3
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2m_e2e.json 2> gpurun_out/r2m_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/r2m_e2e.json')); print('value', d['value'], 'e2e', d['e2e'])"
