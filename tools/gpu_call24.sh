#!/bin/bash
# programmatic dependent launch through the sweep's kernel chain: full GPU tests, probe with PDL on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call24_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call24_tests.log
tail -5 gpurun_out/r02_call24_tests.log
echo "--- PDL on"
timeout 300 python tools/r02b_probe.py --quick 2>&1 | cut -c1-900 | tee gpurun_out/r02_call24_probe_pdl.log
echo "--- PDL off"
SLB200_PDL=0 timeout 300 python tools/r02b_probe.py --quick 2>&1 | cut -c1-900 | tee gpurun_out/r02_call24_probe_nopdl.log
