#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call13_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call13_tests.log
tail -4 gpurun_out/r02_call13_tests.log
timeout 300 python tools/mean_floor_probe.py > gpurun_out/r02_call13_floor.jsonl 2> gpurun_out/r02_call13_floor.err
cat gpurun_out/r02_call13_floor.jsonl; tail -2 gpurun_out/r02_call13_floor.err
timeout 600 python bench.py > gpurun_out/r02_call13_bench.json 2> gpurun_out/r02_call13_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call13_bench.json').read().strip().splitlines()[-1])
print('n1 weak', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'), d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filter_head_kernel' -s 5 -c 1 -o gpurun_out/r02_filter_head -f python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call13_ncu_head.log 2>&1
echo "ncu head rc=$?"
