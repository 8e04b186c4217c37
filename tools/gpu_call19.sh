#!/bin/bash
# round 2, second half: fp32 screening + cooperative head means + factor-split refine: parity subset,
# stage times, ncu captures (full set + source) of the three stages of the default C2 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "filter or screening or refine_pass" > gpurun_out/r02_call19_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call19_tests.log
tail -6 gpurun_out/r02_call19_tests.log
timeout 600 python tools/r02b_probe.py > gpurun_out/r02_call19_probe.log 2> gpurun_out/r02_call19_probe.err
cut -c1-1500 gpurun_out/r02_call19_probe.log
tail -5 gpurun_out/r02_call19_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filter_mean32|filter_head|gp_tile' -s 16 -c 4 -o gpurun_out/r02b_stages -f python tools/profile_sweep.py --filtered > gpurun_out/r02_call19_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02_call19_ncu.log
ls -la gpurun_out/*.ncu-rep
