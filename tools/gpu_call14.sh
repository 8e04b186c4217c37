#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call14_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call14_tests.log
tail -4 gpurun_out/r02_call14_tests.log
timeout 200 python tools/e2e_breakdown.py > gpurun_out/r02_call14_e2e.json 2> gpurun_out/r02_call14_e2e.err; cat gpurun_out/r02_call14_e2e.json; tail -2 gpurun_out/r02_call14_e2e.err
SLB200_GRAPHS=1 timeout 200 python tools/e2e_breakdown.py > gpurun_out/r02_call14_e2e_graphs.json 2>> gpurun_out/r02_call14_e2e.err; cat gpurun_out/r02_call14_e2e_graphs.json
for g in 0 1; do
SLB200_GRAPHS=$g timeout 600 python bench.py > gpurun_out/r02_call14_bench_g$g.json 2> gpurun_out/r02_call14_bench_g$g.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call14_bench_g$g.json').read().strip().splitlines()[-1])
print('graphs=$g', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline'].get('stage_ms'), d['parity']['mismatches'])"
done
