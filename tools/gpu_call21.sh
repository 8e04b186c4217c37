#!/bin/bash
# round 2, second half: overlapped restore (factor dependency event); ncu capture of the head kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "restore or screening or bench" > gpurun_out/r02_call21_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call21_tests.log
tail -6 gpurun_out/r02_call21_tests.log
timeout 600 python bench.py --steps 30 > gpurun_out/r02_call21_bench.json 2> gpurun_out/r02_call21_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_call21_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity'])
print('roofline', {k: d['roofline'].get(k) for k in ('kernel_ms', 'frac', 'stage_ms', 'stage1', 'exp_frac')})
print('clocks', d['clocks'])
PY
tail -3 gpurun_out/r02_call21_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filter_head' -s 4 -c 1 -o gpurun_out/r02b_head -f python tools/profile_sweep.py --filtered > gpurun_out/r02_call21_ncu.log 2>&1
echo "ncu rc=$?"
