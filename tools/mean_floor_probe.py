"""How much of the filter's mean stage is per-point prologue / epilogue (x, policy, V(x), threshold,
V(mu), L_V(mu), decision, list append) rather than the M-row loop: the stage timed alone on the C2
grid with M = 500, 100, 8 and 0 training points."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W  # noqa: E402
from safe_learning_b200 import _native as nat  # noqa: E402

lib = nat.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, steps=30, warm=5, cold=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(steps):
        if cold:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([x.elapsed_time(y) for x, y in ev]))


for M in (500, 100, 8, 0):
    par = W.make_pendulum(num_points=256, M=max(M, 1))
    if M == 0:
        par["X"], par["Y"] = par["X"][:0], par["Y"][:0]
    lyap = W.build_product(par)
    lyap.filter = True
    lib.slb_debug_filter_stages(0)
    out = {"M": M, "mean_stage_ms_warm": timed(lyap.compute_negative),
           "mean_stage_ms_cold": timed(lyap.compute_negative, cold=True)}
    lib.slb_debug_filter_stages(3)
    print(json.dumps(out))
