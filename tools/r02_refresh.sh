#!/bin/bash
# Final round-2 artefacts on ONE B200 (run under gpurun): tests, smoke, bench (both arms), launch list,
# ncu captures of the three stages, secondary measurements.  Outputs: gpurun_out/r02_*
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
timeout 400 python bench.py --scaling strong --steps 10 > gpurun_out/r02_bench_n1_strong.json 2>> gpurun_out/r02_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_b_ncu.log 2>&1
for k in filter_mean_kernel filter_head_kernel gp_tile_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -o gpurun_out/r02_$k -f python bench.py --steps 2 --warmup 1 > gpurun_out/r02_ncu_$k.log 2>&1
done
# the full-posterior tile kernel over the whole grid (filter off): the kernel of `roofline_full_posterior`
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gp_tile -s 3 -c 1 -o gpurun_out/r02_gp_tile_full -f python tools/profile_sweep.py > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 python tools/bench_extra.py bellman argmax det det_linear c5 shared c4 nb > gpurun_out/r02_bench_extra.jsonl 2> gpurun_out/r02_extra.err; cut -c1-220 gpurun_out/r02_bench_extra.jsonl
timeout 200 python tools/r02_probe.py > gpurun_out/r02_filter_probe.jsonl 2>> gpurun_out/r02_extra.err
timeout 200 python tools/mean_floor_probe.py > gpurun_out/r02_mean_floor.jsonl 2>> gpurun_out/r02_extra.err
timeout 200 python tools/e2e_breakdown.py > gpurun_out/r02_e2e_breakdown.json 2>> gpurun_out/r02_extra.err
timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke 2>&1 | tail -3 > gpurun_out/r02_racecheck.txt; cat gpurun_out/r02_racecheck.txt
