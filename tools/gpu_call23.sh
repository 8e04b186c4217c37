#!/bin/bash
# 2 GPUs: sharded parity test + bench (weak) with the screened pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_call23_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call23_tests.log
tail -4 gpurun_out/r02_call23_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 > gpurun_out/r02_call23_bench_n2.json 2> gpurun_out/r02_call23_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_call23_bench_n2.json').read().strip().splitlines()[-1])
print('n2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity'] and d['parity']['mismatches'], d['exchange'][:40])
PY
tail -3 gpurun_out/r02_call23_bench_n2.err
