#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call12_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call12_tests.log
tail -4 gpurun_out/r02_call12_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call12_probe.jsonl 2> gpurun_out/r02_call12_probe.err
cut -c1-330 gpurun_out/r02_call12_probe.jsonl
timeout 600 python bench.py --scaling strong > gpurun_out/r02_call12_bench_strong.json 2> gpurun_out/r02_call12_bench_strong.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call12_bench_strong.json').read().strip().splitlines()[-1])
print('n1 strong', d['value'], d['ms_per_step'], d['roofline'].get('stage_ms'))"
timeout 600 python bench.py > gpurun_out/r02_call12_bench.json 2> gpurun_out/r02_call12_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call12_bench.json').read().strip().splitlines()[-1])
print('n1 weak', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'), d['roofline']['frac'])"
