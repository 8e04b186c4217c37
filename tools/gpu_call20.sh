#!/bin/bash
# round 2, second half: head stage decides from the screened box first; persistent unsplit refine launch;
# operand prefetches.  Full GPU test suite (1 GPU), stage probe, launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r02_call20_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call20_tests.log
tail -6 gpurun_out/r02_call20_tests.log
timeout 600 python tools/r02b_probe.py > gpurun_out/r02_call20_probe.log 2> gpurun_out/r02_call20_probe.err
cut -c1-1500 gpurun_out/r02_call20_probe.log
tail -5 gpurun_out/r02_call20_probe.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|gp_tile|first_fail|apply_prefix' -s 28 -c 14 --csv --log-file gpurun_out/r02_call20_launches.csv python tools/profile_sweep.py --filtered > gpurun_out/r02_call20_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_call20_launches.csv | cut -d, -f5,14- | cut -c1-200 | tail -14
