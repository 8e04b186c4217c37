"""Round-2 probe: timings of the default (filtered) sweep, its two kernels and the full posterior
on the C2 workload, and how the filter's fractions move with the discretisation constant tau."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W  # noqa: E402


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(t)), float(np.min(t))


for M in (500,):
    for tau_scale in (1.0, 0.25, 1 / 16., 1 / 64., 0.0):
        par = W.make_pendulum(num_points=256, M=M, shared_hypers=False, tau_scale=tau_scale)
        lyap = W.build_product(par)
        lyap.reset_filter_stats()
        lyap.compute_negative()
        fs = lyap.filter_stats
        med, best = timed(lyap.compute_negative)
        step, step_best = timed(lyap.update_safe_set)
        lyap.filter = False
        full, full_best = timed(lyap.compute_negative)
        step_full, _ = timed(lyap.update_safe_set)
        n = lyap.discretization.nindex
        print(json.dumps({"probe": "filter", "grid": "256x256", "M": M, "tau_scale": tau_scale,
                          "negative_frac": float(lyap.compute_negative().float().mean().item()),
                          "filter": {k: v / max(fs["points"], 1) for k, v in fs.items() if k != "points"},
                          "decision_ms": med, "decision_ms_best": best, "step_ms": step,
                          "full_decision_ms": full, "full_step_ms": step_full,
                          "points_per_s": n / (step * 1e-3)}))
