"""Time the C2 decision pass with alternative builds of the filter (tools/build_variants.sh): one
subprocess per library (SLB200_LIB), CUDA-event medians of the filtered compute_negative."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
import bench_workloads as W
par = W.make_pendulum(num_points=256, M=500, shared_hypers=False)
lyap = W.build_product(par)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, steps=30, warm=5, cold=False):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(steps):
        if cold: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([x.elapsed_time(y) for x, y in ev]))
print(json.dumps({"decision_ms_warm": timed(lyap.compute_negative), "decision_ms_cold": timed(lyap.compute_negative, cold=True),
                  "step_ms_cold": timed(lyap.update_safe_set, cold=True)}))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "safe_learning_b200", "variants", "libslb200_*.so"))):
    env = dict(os.environ, SLB200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
    print(os.path.basename(lib), line)
