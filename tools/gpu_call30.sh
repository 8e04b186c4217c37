#!/bin/bash
# final validation of the round: every GPU test on one B200, smoke, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r02_call30_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call30_tests.log
tail -5 gpurun_out/r02_call30_tests.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_call30_bench.json 2> gpurun_out/r02_call30_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_call30_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_spread'], 'parity', d['parity']['mismatches'], 'launches', d['gpu_launches'], 'steps', d['steps'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
