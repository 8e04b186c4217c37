#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "restore" > gpurun_out/r02_call22_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call22_tests.log
tail -4 gpurun_out/r02_call22_tests.log
timeout 600 python tools/e2e_breakdown.py --profile > gpurun_out/r02_call22_e2e.log 2>&1
head -c 6000 gpurun_out/r02_call22_e2e.log
