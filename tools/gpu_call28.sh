#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "other_lipschitz_forms" > gpurun_out/r02_call28_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call28_tests.log
tail -25 gpurun_out/r02_call28_tests.log
