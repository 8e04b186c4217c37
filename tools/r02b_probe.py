"""Round-2 (second half) probe for one gpurun call: stage times of the default C2 step with the fp32
screening stage and with the fp64 mean stage, filter statistics of both, parity of both against the
full posterior, and the per-phase clocks of the refine pass's CTAs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench_workloads as W
from safe_learning_b200 import _native as nat

lib = nat.load()
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, steps=30):
    ev = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in ev)
    return per[len(per) // 2]


def run(num, M, label, tau_scale=1.0):
    par = W.make_pendulum(num_points=num, M=M, shared_hypers=False, tau_scale=tau_scale)
    lyap = W.build_product(par)
    lyap.filter = False
    full = lyap.compute_negative().cpu().numpy().copy()
    lyap.filter = "auto"
    res = {}
    for name, base in (("fp32", 0), ("fp64", 4)):
        lib.slb_debug_filter_stages(3 | base)
        lyap.reset_filter_stats()
        flags = lyap.compute_negative().cpu().numpy().copy()
        st = lyap.filter_stats
        r = {"mismatches_vs_full": int((flags != full).sum()), "stats": st}
        lyap.__dict__["_sweep_graph"] = None          # the cached CUDA graph belongs to the other stage 1
        lyap.__dict__["_sweep_graph_seen"] = None
        for _ in range(4):
            lyap.update_safe_set()
        r["update_safe_set_ms"] = timed(lyap.update_safe_set)
        for stage, mask in (("mean", 0), ("mean_head", 1), ("all", 3)):
            lib.slb_debug_filter_stages(mask | base)
            for _ in range(3):
                lyap.compute_negative()
            r[stage + "_ms"] = timed(lyap.compute_negative)
        res[name] = r
    lib.slb_debug_filter_stages(3)
    out[label] = res
    print(label, json.dumps(res), flush=True)
    return lyap


lyap = run(256, 500, "c2_256_M500")
if "--quick" in sys.argv:
    sys.exit(0)
run(256, 500, "c2_tau_64x_finer", tau_scale=1 / 64.)
run(512, 500, "c5_512_M500")
run(256, 100, "c5_256_M100")
run(1024, 500, "c5_1024_M500")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/r02b_probe.json", "w") as f:
    json.dump(out, f, indent=1)
