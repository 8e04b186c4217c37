#!/bin/bash
# round 2, second half: fp32 screening stage -- parity subset, probe of stage times and refine clocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or screening or refine_pass" > gpurun_out/r02_call18_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call18_tests.log
tail -6 gpurun_out/r02_call18_tests.log
timeout 600 python tools/r02b_probe.py > gpurun_out/r02_call18_probe.log 2> gpurun_out/r02_call18_probe.err
cut -c1-1200 gpurun_out/r02_call18_probe.log
tail -5 gpurun_out/r02_call18_probe.err
echo "--- SLB200_SPLIT_FACTORS=0"
SLB200_SPLIT_FACTORS=0 timeout 300 python tools/r02b_probe.py --quick 2>&1 | cut -c1-1200 | tee gpurun_out/r02_call18_probe_nofs.log
echo "--- SLB200_HEAD_PREFETCH=0"
SLB200_HEAD_PREFETCH=0 timeout 300 python tools/r02b_probe.py --quick 2>&1 | cut -c1-1200 | tee gpurun_out/r02_call18_probe_nopf.log
