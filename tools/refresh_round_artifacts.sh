set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r01_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r01_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gp_tile -s 3 -c 1 -o gpurun_out/r01_gp_tile -f python tools/profile_sweep.py > gpurun_out/ncu_full.log 2>&1
timeout 600 python tools/bench_extra.py > gpurun_out/r01_bench_extra.jsonl 2> gpurun_out/extra.err; cat gpurun_out/r01_bench_extra.jsonl | cut -c1-200
