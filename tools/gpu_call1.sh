#!/bin/bash
# round-2 GPU call 1: parity tests, probe, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_call1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call1_tests.log
tail -5 gpurun_out/r02_call1_tests.log
timeout 300 python tools/r02_probe.py > gpurun_out/r02_call1_probe.jsonl 2> gpurun_out/r02_call1_probe.err
tail -5 gpurun_out/r02_call1_probe.jsonl
timeout 600 python bench.py > gpurun_out/r02_call1_bench.json 2> gpurun_out/r02_call1_bench.err
tail -c 3000 gpurun_out/r02_call1_bench.json
tail -3 gpurun_out/r02_call1_bench.err
