"""Worker for the multi-GPU tests: one process per GPU (NCCL), grid sharded by index range.
Checks the sharded Lyapunov sweep and Bellman sweep against the single-process CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

local_rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()

import bench_workloads as W  # noqa: E402
import oracle as O  # noqa: E402
import safe_learning_b200 as sl  # noqa: E402

# --- Lyapunov sweep, grid size not divisible by the world size or the 64-point tile
for num, tau_scale in (([37, 23], 1 / 48.), ([48, 48], 1 / 48.), ([9, 7], 0.0)):
    par = W.make_pendulum(num_points=num, M=70, tau_scale=tau_scale)
    gpu = W.build_product(par)
    cpu = W.build_oracle(par)
    begin, end = gpu._begin, gpu._end
    assert (begin, end) == sl._device.shard_range(cpu.discretization.nindex, rank, world)
    gpu.update_safe_set()
    cpu.update_safe_set()
    assert np.array_equal(gpu.values, cpu.values)
    assert np.array_equal(gpu.safe_set, cpu.safe_set), (rank, num)
    assert gpu.feed_dict[gpu.c_max] == cpu.c_max, (rank, num, gpu.feed_dict[gpu.c_max], cpu.c_max)
    assert gpu.last_sweep["n_safe"] == int(cpu.safe_set.sum())

# --- covariance expressions (the notebook kernels) on sharded slabs
par = W.make_pendulum(num_points=[24, 19], M=90, tau_scale=1 / 150., with_prior_mean=True, seed=3)
par["kernel_specs"] = W.notebook_pendulum_kernels([[2e-3, 6e-3, 1.5e-3], [2.5e-2, 8e-3, 1.2e-2]])
gpu, cpu = W.build_product(par), W.build_oracle(par)
gpu.update_safe_set()
cpu.update_safe_set()
assert np.array_equal(gpu.safe_set, cpu.safe_set) and 50 < cpu.safe_set.sum() < cpu.safe_set.size
assert gpu.feed_dict[gpu.c_max] == cpu.c_max

# --- adaptive refinement on sharded slabs (several batches)
par = W.make_pendulum(num_points=[26, 21], M=90, tau_scale=1 / 30.)
old_batch = (sl.config.gp_batch_size, O.config.gp_batch_size)
sl.config.gp_batch_size = O.config.gp_batch_size = 64
pair = []
for ns, kind in ((sl, "product"), (O, "oracle")):
    grid, dyn = W._build(ns, par, kind)
    policy = ns.Saturation(ns.LinearSystem(-par["K"]), -1., 1.)
    pair.append(ns.Lyapunov(grid, ns.QuadraticFunction(par["P"]), dyn, par["L_dyn"],
                            ns.AbsFunction(ns.LinearSystem((2 * par["P"],))), par["tau"], policy,
                            initial_set=par["initial"], adaptive=True))
gpu, cpu = pair
for kwargs in (dict(), dict(max_refinement=8, safety_factor=2.0)):
    gpu.update_safe_set(**kwargs)
    cpu.update_safe_set(**kwargs)
    assert np.array_equal(gpu.safe_set, cpu.safe_set), (rank, kwargs)
    assert np.array_equal(gpu._refinement, cpu._refinement), (rank, kwargs)
    assert gpu.feed_dict[gpu.c_max] == cpu.c_max, (rank, kwargs)
assert cpu._refinement.max() > 1
sl.config.gp_batch_size, O.config.gp_batch_size = old_batch

# --- Bellman sweep: slabs all-gathered into the full vertex table every sweep
import scipy.linalg  # noqa: E402
par = W.make_pendulum(num_points=8, M=80)


def make_rl(ns, kind):
    grid = ns.GridWorld(par["limits"], [21, 17])
    _, dyn = W._build(ns, par, kind)
    policy = ns.Saturation(ns.LinearSystem(-par["K"]), -1., 1.)
    reward = ns.QuadraticFunction(-scipy.linalg.block_diag(np.diag([1., 2.]), 1.2 * np.eye(1)))
    value = ns.Triangulation(grid, np.zeros((grid.nindex, 1)), project=True)
    return ns.PolicyIteration(policy, dyn, reward, value, gamma=0.98), value


rl_g, v_g = make_rl(sl, "product")
rl_c, v_c = make_rl(O, "oracle")
for _ in range(3):
    rl_g.value_iteration()
    rl_c.value_iteration()
    np.testing.assert_allclose(v_g.parameters[0], v_c.parameters, rtol=1e-9, atol=1e-12)

dist.barrier()
if rank == 0:
    print("dist worker ok, world", world)
dist.destroy_process_group()
