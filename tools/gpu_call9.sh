#!/bin/bash
# 2-GPU call: full GPU suite incl. the sharded test, bench at N=2 (weak, strong), N=1 for comparison
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_call9_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_call9_tests.log
tail -5 gpurun_out/r02_call9_tests.log
for mode in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --scaling $mode > gpurun_out/r02_call9_bench_n2_$mode.json 2> gpurun_out/r02_call9_bench_n2_$mode.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r02_call9_bench_n2_$mode.json').read().strip().splitlines()[-1])
    print('$mode', d['value'], d['ms_per_step'], d['parity'], d['exchange'][:60], d['e2e']['value'])
except Exception as e:
    print('bench n2 $mode failed', e)
PY
tail -3 gpurun_out/r02_call9_bench_n2_$mode.err
done
timeout 600 python bench.py --scaling strong > gpurun_out/r02_call9_bench_n1_strong.json 2> gpurun_out/r02_call9_bench_n1_strong.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_call9_bench_n1_strong.json').read().strip().splitlines()[-1])
print('n1 strong', d['value'], d['ms_per_step'], d['filter']['refined_by_full_posterior'])"
