#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call10_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call10_tests.log
tail -4 gpurun_out/r02_call10_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call10_probe.jsonl 2> gpurun_out/r02_call10_probe.err
cut -c1-330 gpurun_out/r02_call10_probe.jsonl
timeout 300 python tools/mean_floor_probe.py > gpurun_out/r02_call10_floor.jsonl 2> gpurun_out/r02_call10_floor.err
cat gpurun_out/r02_call10_floor.jsonl; tail -2 gpurun_out/r02_call10_floor.err
timeout 600 python bench.py --scaling strong > gpurun_out/r02_call10_bench_strong.json 2> gpurun_out/r02_call10_bench_strong.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call10_bench_strong.json').read().strip().splitlines()[-1])
print('n1 strong', d['value'], d['ms_per_step'], d['roofline'].get('stage_ms'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -c 100 --csv --log-file gpurun_out/r02_call10_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call10_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
