#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call7_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call7_tests.log
tail -5 gpurun_out/r02_call7_tests.log
timeout 600 python tools/bench_extra.py det_linear det > gpurun_out/r02_call7_extra.jsonl 2> gpurun_out/r02_call7_extra.err
cut -c1-400 gpurun_out/r02_call7_extra.jsonl; tail -3 gpurun_out/r02_call7_extra.err
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call7_probe.jsonl 2> gpurun_out/r02_call7_probe.err
cut -c1-330 gpurun_out/r02_call7_probe.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -c 100 --csv --log-file gpurun_out/r02_call7_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call7_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bellman_argmax_tile' -s 1 -c 1 -o gpurun_out/r02_argmax_tile python tools/argmax_profile.py > gpurun_out/r02_call7_ncu_argmax.log 2>&1
echo "ncu argmax rc=$?"
