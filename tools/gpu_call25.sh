#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call25_tests.log
tail -5 gpurun_out/r02_call25_tests.log
timeout 300 python tools/r02b_probe.py --quick 2>&1 | cut -c1-900 | tee gpurun_out/r02_call25_probe.log
timeout 300 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/r02_call25_e2e.log
timeout 600 python bench.py --steps 30 > gpurun_out/r02_call25_bench.json 2> gpurun_out/r02_call25_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_call25_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_spread'], 'parity', d['parity']['mismatches'])
PY
