"""Small discrete_policy_optimization run for ncu (factored path, 128 x 128 states, 101 actions)."""
import os, sys
import numpy as np, scipy.linalg, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as W
import safe_learning_b200 as sl
par = W.make_pendulum(num_points=8, M=500)
grid = sl.GridWorld(par["limits"], int(sys.argv[1]) if len(sys.argv) > 1 else 128)
_, dyn = W._build(sl, par, "product")
reward = sl.QuadraticFunction(-scipy.linalg.block_diag(np.diag([1., 2.]), 1.2 * np.eye(1)))
value = sl.Triangulation(grid, np.random.default_rng(0).normal(size=(grid.nindex, 1)), project=True)
policy = sl.Triangulation(grid, np.zeros((grid.nindex, 1)), project=True)
rl = sl.PolicyIteration(policy, dyn, reward, value, gamma=0.98)
actions = np.linspace(-1, 1, 101).reshape(-1, 1)
for _ in range(3):
    rl.discrete_policy_optimization(actions)
torch.cuda.synchronize()
print("ok")
