#!/bin/bash
# round-2 GPU call 2: all parity tests (no -x), probe, bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_call2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_call2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call2_tests.log
tail -5 gpurun_out/r02_call2_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call2_probe.jsonl 2> gpurun_out/r02_call2_probe.err
timeout 600 python bench.py > gpurun_out/r02_call2_bench.json 2> gpurun_out/r02_call2_bench.err
tail -c 1500 gpurun_out/r02_call2_bench.json
tail -3 gpurun_out/r02_call2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_call2_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_call2_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
