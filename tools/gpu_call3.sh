#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call3_tests.log
tail -5 gpurun_out/r02_call3_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call3_probe.jsonl 2> gpurun_out/r02_call3_probe.err
cut -c1-400 gpurun_out/r02_call3_probe.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_call3_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call3_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
