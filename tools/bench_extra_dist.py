"""Multi-GPU secondary measurements (launch with torchrun, one rank per GPU; also runs on 1 GPU):

  c4     cart-pole 64^4 = 16.8 M grid points, four RBF GPs on 5-D inputs (M=2000, four Cholesky
         factors), V = LyapunovNetwork(4, [64, 64, 64], tanh): one full update_safe_set, the grid
         sharded by contiguous index range over the ranks (BASELINE.json config C4)
  c5     resolution x M points of the pendulum sweep (BASELINE.json config C5) on the same ranks

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29531 tools/bench_extra_dist.py [c4] [c5]

Rank 0 prints one JSON line per measurement: device time (CUDA events, max over ranks) of the whole
sweep incl. the first-fail reduction and the inter-rank key exchange, points/s of the whole job,
the filter's fractions, and the safe-set size (identical on every N by construction -- compare the
lines of different N)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import bench_workloads as W  # noqa: E402
import safe_learning_b200 as sl  # noqa: E402,F401
from bench import algorithmic_flops_per_point  # noqa: E402


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    barrier()
    per = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
        barrier()
    ms = float(np.median(per))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def filter_fractions(lyap):
    lyap.reset_filter_stats()
    lyap.compute_negative()
    fs = lyap.filter_stats
    t = torch.tensor([fs["prior"], fs["head"], fs["refined"], fs["points"]], dtype=torch.int64,
                     device="cuda")
    if world > 1:
        dist.all_reduce(t)
    a, b, c, n = (int(v) for v in t.cpu())
    n = max(n, 1)
    return {"prior": a / n, "head": b / n, "refined": c / n}


def report(name, lyap, M, factors, outputs, din, steps, warmup, extra=None):
    n = lyap.discretization.nindex
    ms = timed(lyap.update_safe_set, steps, warmup)
    fr = filter_fractions(lyap)
    n_safe = int(lyap.last_sweep.get("n_safe", -1))
    lyap.filter = False
    ms_full = timed(lyap.update_safe_set, 1, 0) if extra and extra.get("time_full") else None
    lyap.filter = "auto"
    if rank == 0:
        fl = algorithmic_flops_per_point(M, din, factors, outputs)
        line = {"bench": name, "n_gpus": world, "grid_points": n, "M": M, "factors": factors,
                "ms_per_update_safe_set": ms, "points_per_s": n / (ms * 1e-3),
                "algorithmic_tflops_equivalent": fl * n / (ms * 1e-3) * 1e-12,
                "filter": fr, "safe_points": n_safe}
        if ms_full is not None:
            line["full_posterior_ms"] = ms_full
            line["full_posterior_tflops"] = fl * n / (ms_full * 1e-3) * 1e-12
        print(json.dumps(line), flush=True)


def c4():
    par = W.make_cartpole(num_points=64, M=2000, tau_scale=0.01, with_initial=False)
    lyap = W.build_product(par)
    report("c4_cartpole_64^4_update_safe_set", lyap, 2000, 4, 4, 5, steps=2, warmup=1)


def c5():
    for num, M, full in ((2048, 500, False), (4096, 500, False), (1024, 2000, False),
                         (512, 5000, False), (2048, 100, False)):
        par = W.make_pendulum(num_points=num, M=M)
        lyap = W.build_product(par)
        report("c5_pendulum_%dx%d" % (num, num), lyap, M, 2, 2, 3, steps=3, warmup=1)
        del lyap
        torch.cuda.empty_cache()


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c4", "c5"]):
        globals()[name]()
    if world > 1:
        dist.destroy_process_group()
