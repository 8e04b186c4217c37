#!/bin/bash
# 8 GPUs: weak-scaling bench line of the final code (the driver's SCALE run does N=1,2,4,8 itself)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 > gpurun_out/r02b_bench_n8.json 2> gpurun_out/r02b_bench_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_n8.json').read().strip().splitlines()[-1])
print('n8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity'] and d['parity']['mismatches'], d['exchange'][:40])
print(d['filter'])
print(d['roofline'].get('stage_ms'))
PY
tail -3 gpurun_out/r02b_bench_n8.err
