#!/usr/bin/env python
"""bench.py -- grid-point Lyapunov checks/sec on the 2-D inverted pendulum (BASELINE.json).

One "step" = one complete ``Lyapunov.update_safe_set()`` over a 256x256 GridWorld per GPU with
two stacked RBF GPs (M=500, distinct hyper-parameters => two Cholesky factors): fused sweep
kernel (GP posterior + decrease test for EVERY grid point, no early exit) + first-fail reduction
+ the per-sweep collective + prefix application.  With N GPUs the grid is (256 N) x 256 and each
rank owns one contiguous 256x256 slab (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  ``--impl reference`` times the reference algorithm's CPU path
(the numpy oracle, all host threads) on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point Lyapunov checks/sec (2D pendulum, M=500 GP)"
UNIT = "points/s"
GRID = 256
M_TRAIN = 500


def algorithmic_flops_per_point(M, d_in, n_factors, n_outputs):
    """sum over distinct factors of [M^2 + M (3 d_in + 6)] + D M E_exp
    (E_exp = 1) + F_small (~100)."""
    return n_factors * (M * M + M * (3 * d_in + 6)) + n_outputs * M * 1 + 100


def algorithmic_bytes_per_point(d):
    """8 d (coordinates, charged although generated) + 1 (flag) + 8 (V written)."""
    return 8 * d + 1 + 8


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if smax and v > 0.4 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU reference
def cpu_reference_rate(par, seconds_budget, steps=1, warmup=0):
    """Reference algorithm on the host (numpy/scipy oracle, BLAS threads = all cores): the
    per-batch graph of lyapunov.py:433-441 over `sample` grid points in 10 000-point batches
    (early exit disabled).  Returns (points/s, cores, sample description, per-step seconds)."""
    import bench_workloads as W
    import oracle as O
    # every host thread this process may use -- not the BLAS pools' current size, which torchrun
    # pins to 1 through OMP_NUM_THREADS (threadpool_limits below can raise it again)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        threads = os.cpu_count() or 1
    lyap = W.build_oracle(par)
    grid = lyap.discretization
    batch = O.config.gp_batch_size
    order = O.stable_value_order(lyap.values)
    # calibrate on one batch; BLAS with every hardware thread is often slower than with fewer on
    # these [M x 10 000] panels, so give the CPU its best thread count (reported as `cores`)
    limiter = None
    t_batch = None
    try:
        from threadpoolctl import threadpool_limits
        best = None
        for nt in sorted({threads, 32, 16, 8}, reverse=True):
            if nt > threads:
                continue
            with threadpool_limits(limits=nt):
                lyap.negative(grid.index_to_state(order[:batch]))
                t0 = time.perf_counter()
                lyap.negative(grid.index_to_state(order[:batch]))
                dt = time.perf_counter() - t0
            if best is None or dt < best[1]:
                best = (nt, dt)
        threads, t_batch = best
        limiter = threadpool_limits(limits=threads)
    except Exception:  # pragma: no cover
        pass
    if t_batch is None:
        t0 = time.perf_counter()
        lyap.negative(grid.index_to_state(order[:batch]))
        t_batch = time.perf_counter() - t0
    total_steps = max(1, steps + warmup)
    nb_max = -(-grid.nindex // batch)
    nb = int(max(1, min(nb_max, seconds_budget / total_steps / max(t_batch, 1e-6))))
    sample = order[:min(nb * batch, grid.nindex)]
    times = []
    for s in range(total_steps):
        t0 = time.perf_counter()
        for i, (idx,) in O.batchify((sample,), batch):
            lyap.negative(grid.index_to_state(idx))
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    if limiter is not None:
        limiter.restore_original_limits()
    rate = len(sample) / (sum(times) / len(times))
    desc = ("%d of %d grid points (V-sorted order, %d batches of %d, early exit disabled), "
            "numpy/scipy oracle, %d BLAS threads" % (len(sample), grid.nindex, nb, batch, threads))
    return rate, threads, desc, times


# --------------------------------------------------------------------------- main arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    import bench_workloads as W
    par = W.make_pendulum(num_points=GRID, M=M_TRAIN, shared_hypers=False)
    rate, cores, desc, times = cpu_reference_rate(par, seconds_budget=150.0, steps=args.steps,
                                                  warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(world):
    return {"workload": "inverted pendulum 2D, %dx%d GridWorld per GPU (global %dx%d), 2 stacked "
                        "RBF GPs on [x,u] (M=%d, distinct ARD hyper-parameters => 2 Cholesky "
                        "factors), linear prior mean, saturated LQR policy, quadratic V, "
                        "update_safe_set full-grid (no early exit)"
                        % (GRID, GRID, GRID * world, GRID, M_TRAIN),
            "grid_points_per_gpu": GRID * GRID, "M": M_TRAIN, "gp_outputs": 2, "gp_factors": 2,
            "parallelism": "grid sharded by contiguous index range, %d rank(s)" % world,
            "l2": "L2 flushed (256 MiB write) before every timed step"}


def run_ours(args, rank, world, local_rank):
    import torch
    import __graft_entry__
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "safe_learning_b200", "libslb200.so")):
        __graft_entry__.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (safe_learning_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import bench_workloads as W
    from safe_learning_b200 import _native as nat

    par = W.make_pendulum(num_points=[GRID * world, GRID], M=M_TRAIN, shared_hypers=False)
    lyap = W.build_product(par)
    n_local = lyap._end - lyap._begin
    n_total = lyap.discretization.nindex
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Sum of per-step CUDA-event times.  L2 is flushed (256 MiB write, outside the event
        pair) before each step; steps are enqueued without extra host synchronisation, so the
        GPU sees the same back-to-back cadence as a learning loop."""
        events = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            events.append((e0, e1))
        torch.cuda.synchronize()
        per = [a.elapsed_time(b) for a, b in events]
        return float(sum(per)), per

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: whole update_safe_set per step
    # The clock sampler (nvidia-smi, 100 ms period) runs from here to the end of the e2e arm; the
    # warm-up is stretched by 600 sweeps (~1 s) so that samples under load exist even though the
    # timed region itself lasts only K x ~1.5 ms.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # a FIXED count (identical on every rank -- each sweep contains a collective): W + 600 sweeps
    n_warm = max(args.warmup, 3) + 600
    for _ in range(n_warm):
        lyap.update_safe_set()
    barrier()
    launches0 = nat.launch_count()
    ms_total, _ = timed(lyap.update_safe_set, args.steps)
    launches = nat.launch_count() - launches0
    barrier()
    ms_total = max_over_ranks(ms_total)
    value = n_total * args.steps / (ms_total * 1e-3)

    # ---- dominant kernel alone (roofline): the fused sweep kernel
    for _ in range(3):
        lyap.compute_negative()
    torch.cuda.synchronize()
    k_total, k_per = timed(lyap.compute_negative, args.steps)
    kernel_ms = k_total / args.steps

    # ---- end-to-end arm: host buffers in, host buffers out, every step
    gps = [f.gaussian_process for f in lyap.dynamics.functions]
    host_in, dev_dst = [], []
    for gp in gps:
        gp._ensure()
        for t in (gp._factor.Xs, gp._factor.Wpack, gp._alpha_dev, gp._gamma_dev):
            host_in.append(t.cpu().pin_memory())
            dev_dst.append(t)
    init_mask = np.zeros(n_total, dtype=bool)
    init_mask[par["initial"]] = True
    init_host = torch.from_numpy(init_mask[lyap._begin:lyap._end].astype(np.uint8)).pin_memory()
    lyap._initial_device()
    safe_host = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host_in) + init_host.numel()
    d2h = safe_host.numel() + 32 + 32

    def e2e_step():
        for src, dst in zip(host_in, dev_dst):
            dst.copy_(src, non_blocking=True)
        lyap._initial_dev.copy_(init_host, non_blocking=True)
        lyap.update_safe_set()
        safe_host.copy_(lyap._safe_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return lyap.feed_dict[lyap.c_max]

    for _ in range(3):
        e2e_step()
    barrier()
    e_total, _ = timed(e2e_step, args.steps)
    barrier()
    e_total = max_over_ranks(e_total)
    e2e_value = n_total * args.steps / (e_total * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    flops_pt = algorithmic_flops_per_point(M_TRAIN, 3, 2, 2)
    achieved_tf = flops_pt * n_local / (kernel_ms * 1e-3) * 1e-12
    peak_tf, peak_src = 37.1, "fallback 37.1 (tools/fp64_peaks.cu on this pool, r01)"
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")) as fh:
            peak_tf = float(json.load(fh)["dmma_tflops_w8_acc8"])
            peak_src = "measured DMMA.8x8x4 peak, tools/fp64_peaks.cu (profiles/r01_fp64_peaks.json)"
    except Exception:
        pass
    hbm_peak = 6650.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gp_tile_kernel_ncu.json")) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        pass
    hbm_gbs = algorithmic_bytes_per_point(2) * n_local / (kernel_ms * 1e-3) * 1e-9
    roofline = {"bound": "tensor", "kernel": "gp_tile_kernel<3> (fp64 DMMA.8x8x4)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kernel_ms, "algorithmic_flops_per_point": flops_pt,
                "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_gbs / hbm_peak,
                        "note": "path is fp64-compute-bound (AI ~3e4 FLOP/B); HBM fraction "
                                "reported for completeness"}}

    # ---- CPU baseline, bounded sample, same run
    cpu_par = W.make_pendulum(num_points=GRID, M=M_TRAIN, shared_hypers=False)
    rate, cores, desc, _ = cpu_reference_rate(cpu_par, seconds_budget=20.0, steps=1, warmup=0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "warmup_sweeps_run": n_warm, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e_total / args.steps},
        "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": desc},
        "safe_points": int(lyap.last_sweep.get("n_safe", -1)),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
