#!/bin/bash
# 8-GPU call: weak + strong scaling lines, C4 at its stated size, C5 points
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "filtered_flags or refine_pass or pivoted" > gpurun_out/r02_call11_precheck.log 2>&1 || { tail -20 gpurun_out/r02_call11_precheck.log; echo PRECHECK FAILED; exit 1; }
tail -2 gpurun_out/r02_call11_precheck.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 bash -c "$(declare -f run); N=$N; run 29551 bench.py --gpus $N --steps 20 --warmup 3" > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
timeout 400 bash -c "$(declare -f run); N=$N; run 29552 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong" > gpurun_out/r02_bench_n${N}_strong.json 2> gpurun_out/r02_bench_n${N}_strong.err
timeout 900 bash -c "$(declare -f run); N=$N; run 29553 tools/bench_extra_dist.py c4 c5" > gpurun_out/r02_bench_extra_dist_n${N}.jsonl 2> gpurun_out/r02_bench_extra_dist_n${N}.err
python - <<PY
import json
for name in ("r02_bench_n$N.json", "r02_bench_n${N}_strong.json"):
    try:
        d = json.loads(open("gpurun_out/" + name).read().strip().splitlines()[-1])
        print(name, "%.4g pts/s %.4f ms e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["parity"], d["exchange"][:30])
    except Exception as e:
        print(name, "failed", e)
PY
cut -c1-300 gpurun_out/r02_bench_extra_dist_n${N}.jsonl
tail -3 gpurun_out/r02_bench_n${N}.err gpurun_out/r02_bench_extra_dist_n${N}.err | cut -c1-300
