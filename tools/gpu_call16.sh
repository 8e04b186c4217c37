#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call16_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call16_tests.log
tail -4 gpurun_out/r02_call16_tests.log
timeout 600 python tools/bench_extra.py bellman argmax > gpurun_out/r02_call16_extra.jsonl 2> gpurun_out/r02_call16_extra.err
cat gpurun_out/r02_call16_extra.jsonl; tail -3 gpurun_out/r02_call16_extra.err
timeout 300 python bench.py --steps 20 > gpurun_out/r02_call16_bench.json 2> gpurun_out/r02_call16_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call16_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'].get('stage_ms'), d['parity']['mismatches'])"
