#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call17_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call17_tests.log
tail -4 gpurun_out/r02_call17_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call17_probe.jsonl 2> gpurun_out/r02_call17_probe.err
cut -c1-330 gpurun_out/r02_call17_probe.jsonl
for mode in weak strong; do
timeout 300 python bench.py --steps 20 --scaling $mode > gpurun_out/r02_call17_bench_$mode.json 2> gpurun_out/r02_call17_bench_$mode.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call17_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'].get('stage_ms'), d['parity'] and d['parity']['mismatches'])"
done
