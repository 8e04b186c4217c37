#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call5_tests.log
tail -5 gpurun_out/r02_call5_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call5_probe.jsonl 2> gpurun_out/r02_call5_probe.err
cut -c1-330 gpurun_out/r02_call5_probe.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -c 200 --csv --log-file gpurun_out/r02_call5_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call5_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gp_tile_kernel' -s 30 -c 3 -o gpurun_out/r02_refine_tile python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call5_ncu_full.log 2>&1
echo "ncu full rc=$?"
