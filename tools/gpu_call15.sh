#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call15_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call15_tests.log
tail -4 gpurun_out/r02_call15_tests.log
for sm in 8 4 2 1; do
SLB200_SPLIT_MAX=$sm timeout 300 python bench.py --steps 20 > gpurun_out/r02_call15_bench_s$sm.json 2> gpurun_out/r02_call15_bench_s$sm.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call15_bench_s$sm.json').read().strip().splitlines()[-1])
print('split_max=$sm', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'].get('stage_ms'), d['parity']['mismatches'])"
done
