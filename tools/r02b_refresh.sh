#!/bin/bash
# Final artefacts of the second half of round 2 on ONE B200 (run under gpurun): tests, smoke, bench (both
# arms), launch list, ncu captures of the three stages + the full-posterior kernel, secondary measurements.
# Outputs: gpurun_out/r02b_*  (tools/summarize_profiles_r02.py with SLB_TAG=r02b copies them to profiles/)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>> gpurun_out/r02b_bench.err
timeout 400 python bench.py --scaling strong --steps 10 > gpurun_out/r02b_bench_n1_strong.json 2>> gpurun_out/r02b_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02b_b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'filter_mean32|filter_head|gp_tile' -s 16 -c 3 -o gpurun_out/r02b_stages3 -f python tools/profile_sweep.py --filtered > gpurun_out/r02b_ncu_stages.log 2>&1
# the full-posterior tile kernel over the whole grid (filter off): the kernel of `roofline_full_posterior`
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gp_tile -s 3 -c 1 -o gpurun_out/r02b_gp_tile_full -f python tools/profile_sweep.py > gpurun_out/r02b_ncu_full.log 2>&1
timeout 900 python tools/bench_extra.py bellman argmax det det_linear c5 shared c4 nb > gpurun_out/r02b_bench_extra.jsonl 2> gpurun_out/r02b_extra.err; cut -c1-220 gpurun_out/r02b_bench_extra.jsonl
timeout 300 python tools/r02b_probe.py > gpurun_out/r02b_probe.log 2>> gpurun_out/r02b_extra.err
timeout 200 python tools/e2e_breakdown.py > gpurun_out/r02b_e2e_breakdown.json 2>> gpurun_out/r02b_extra.err
timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke 2>&1 | tail -3 > gpurun_out/r02b_racecheck.txt; cat gpurun_out/r02b_racecheck.txt
ls -la gpurun_out/*.ncu-rep
