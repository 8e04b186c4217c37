"""Where the end-to-end step of bench.py spends its time: wall clock of each segment with a device
synchronisation after it (so host enqueue cost + device work of that segment), C2 workload."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W  # noqa: E402

par = W.make_pendulum(num_points=256, M=500, shared_hypers=False)
lyap = W.build_product(par)
tables = lyap.dynamics.export_cache(pinned=True)
mask = np.zeros(lyap.discretization.nindex, dtype=bool)
mask[par["initial"]] = True
for _ in range(5):
    lyap.dynamics.import_cache(tables); lyap.initial_safe_set = mask; lyap.update_safe_set(); lyap.safe_set; lyap.feed_dict[lyap.c_max]
seg = {k: [] for k in ("import_cache", "set_initial", "update_safe_set_enqueue", "update_safe_set_device",
                        "safe_set_read", "c_max_read", "whole_step_no_syncs")}
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); lyap.dynamics.import_cache(tables); torch.cuda.synchronize(); t1 = time.perf_counter()
    lyap.initial_safe_set = mask; t2 = time.perf_counter()
    lyap.update_safe_set(); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    s = lyap.safe_set; t5 = time.perf_counter()
    c = lyap.feed_dict[lyap.c_max]; t6 = time.perf_counter()
    for k, v in zip(seg, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        seg[k].append(v)
    t0 = time.perf_counter()
    lyap.dynamics.import_cache(tables); lyap.initial_safe_set = mask; lyap.update_safe_set(); s = lyap.safe_set; c = lyap.feed_dict[lyap.c_max]
    seg["whole_step_no_syncs"].append(time.perf_counter() - t0)
print(json.dumps({k: round(1e6 * float(np.median(v)), 1) for k, v in seg.items()} | {"unit": "us (median of 30)"}))

# host-only cost of each call (no synchronisation in between: what the enqueueing thread spends)
host = {k: [] for k in ("import_cache", "set_initial", "update_safe_set", "safe_set", "c_max")}
for _ in range(200):
    t0 = time.perf_counter(); lyap.dynamics.import_cache(tables)
    t1 = time.perf_counter(); lyap.initial_safe_set = mask
    t2 = time.perf_counter(); lyap.update_safe_set()
    t3 = time.perf_counter(); s = lyap.safe_set
    t4 = time.perf_counter(); c = lyap.feed_dict[lyap.c_max]
    t5 = time.perf_counter()
    for k, v in zip(host, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        host[k].append(v)
print(json.dumps({"host_only_us": {k: round(1e6 * float(np.median(v)), 1) for k, v in host.items()}}))
if "--profile" in sys.argv:
    import cProfile, pstats
    def steps(n):
        for _ in range(n):
            lyap.dynamics.import_cache(tables); lyap.initial_safe_set = mask; lyap.update_safe_set()
            s = lyap.safe_set; c = lyap.feed_dict[lyap.c_max]
    pr = cProfile.Profile(); pr.enable(); steps(500); pr.disable()
    st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
