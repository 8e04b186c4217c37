#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call6_tests.log
tail -5 gpurun_out/r02_call6_tests.log
timeout 600 python tools/mean_variants.py > gpurun_out/r02_call6_variants.txt 2>&1
cat gpurun_out/r02_call6_variants.txt
timeout 600 python tools/bench_extra.py argmax bellman > gpurun_out/r02_call6_extra.jsonl 2> gpurun_out/r02_call6_extra.err
cat gpurun_out/r02_call6_extra.jsonl; tail -3 gpurun_out/r02_call6_extra.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -c 100 --csv --log-file gpurun_out/r02_call6_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call6_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
