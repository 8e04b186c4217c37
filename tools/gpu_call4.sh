#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call4_tests.log
tail -5 gpurun_out/r02_call4_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call4_probe.jsonl 2> gpurun_out/r02_call4_probe.err
cut -c1-330 gpurun_out/r02_call4_probe.jsonl
timeout 600 python bench.py > gpurun_out/r02_call4_bench.json 2> gpurun_out/r02_call4_bench.err
tail -c 600 gpurun_out/r02_call4_bench.json; tail -3 gpurun_out/r02_call4_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -c 300 --csv --log-file gpurun_out/r02_call4_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call4_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_mean_kernel -s 5 -c 1 -o gpurun_out/r02_filter_mean python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call4_ncu_full.log 2>&1
echo "ncu full rc=$?"
