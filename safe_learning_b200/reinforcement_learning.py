"""``PolicyIteration`` (``safe_learning/reinforcement_learning.py:26-279``): the Bellman sweep.

``value_iteration`` and ``discrete_policy_optimization`` run as fused CUDA sweeps over the
value function's grid (``slb_bellman_sweep`` / ``slb_bellman_argmax``); ``future_values`` on
arbitrary state arrays is composed from the eager GPU evaluations of the function objects.
LP value optimisation (``:142-211``, cvxpy) and ``bellmann_error`` (``:116-133``, autodiff)
are outside this build's hot path.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _device as dev
from . import _native as nat
from .functions import (Function, FunctionStack, GaussianProcess, ScaledFunction, Triangulation,
                        UncertainFunction)

__all__ = ["PolicyIteration", "OptimizationError"]


class OptimizationError(Exception):
    """``reinforcement_learning.py:22-23``."""


def _triangulation_of(value_function):
    """The Triangulation under an optional ``-V`` / ``c * V`` wrapper."""
    base = value_function
    while isinstance(base, ScaledFunction):
        base = base.fun
    if not isinstance(base, Triangulation):
        raise TypeError("value_function must be a Triangulation (possibly scaled)")
    return base


class PolicyIteration(object):
    """See ``reinforcement_learning.py:26-63`` for the parameters."""

    def __init__(self, policy, dynamics, reward_function, value_function, gamma=0.98):
        self.dynamics = dynamics
        self.reward_function = reward_function
        self.value_function = value_function
        self.gamma = gamma
        self.policy = policy
        self.feed_dict = {}
        self._storage = {}
        self.factor_actions = True        # discrete_policy_optimization: see csrc/bellman_tile.cu
        self._grid = _triangulation_of(value_function).discretization
        self._begin, self._end = dev.shard_range(self._grid.nindex)

    @property
    def state_space(self):
        """``value_function.discretization.all_points`` (``:58-59``)."""
        return self._grid.all_points

    # ------------------------------------------------------------------ generic path
    def future_values(self, states, policy=None, actions=None, lyapunov=None,
                      lagrange_multiplier=1.):
        """``r(x, u) + gamma V(mean f(x, u))`` [``- lambda (decrease - threshold)``]
        (``:65-114``); numpy in, numpy [n, 1] out."""
        if isinstance(states, torch.Tensor) or isinstance(actions, torch.Tensor):
            return self._future_values_torch(states, policy, actions, lyapunov, lagrange_multiplier)
        states = np.atleast_2d(np.asarray(states, dtype=np.float64))
        if actions is None:
            actions = (policy or self.policy)(states)
        actions = np.broadcast_to(np.atleast_2d(actions), (states.shape[0],
                                                           np.atleast_2d(actions).shape[1]))
        next_states = self.dynamics(states, actions)
        rewards = self.reward_function(states, actions)
        var = None
        if isinstance(next_states, tuple):
            next_states, var = next_states
        updated = rewards + self.gamma * self.value_function(next_states)
        if lyapunov is not None:
            decrease = lyapunov.v_decrease_bound(states, (next_states, var))
            updated = updated - lagrange_multiplier * (decrease - lyapunov.threshold(states))
        return updated

    def _future_values_torch(self, states, policy, actions, lyapunov, lagrange_multiplier):
        """``future_values`` on device tensors as a differentiable torch expression (the reference
        differentiates this graph with ``tf.gradients`` to optimise a parametric policy under the
        Lyapunov penalty, ``examples/inverted_pendulum.ipynb`` cell 17): every fused function
        object is one autograd node (CUDA evaluation forward, its device Jacobian backward,
        ``Function.torch``), the GP mean / variance are torch operations on the cached Cholesky
        factor.  ``policy`` may be any callable on tensors (e.g. a ``torch.nn.Module``); gradients
        flow to ``actions`` / the policy's parameters and to ``states`` if they require them."""
        states = dev.to_device(states) if not isinstance(states, torch.Tensor) else states
        if actions is None:
            fn = policy or self.policy
            actions = fn.torch(states) if isinstance(fn, Function) else fn(states)
        elif not isinstance(actions, torch.Tensor):
            actions = dev.to_device(np.atleast_2d(np.asarray(actions, dtype=np.float64)))
        actions = actions.expand(states.shape[0], actions.shape[1])
        z = torch.cat((states, actions), dim=1)
        err = None
        if isinstance(self.dynamics, (FunctionStack, GaussianProcess)):
            mean, err = self.dynamics.torch(z)                                   # :92, :98-99
        else:
            mean = self.dynamics.torch(z)
        updated = self.reward_function.torch(z) + self.gamma * self.value_function.torch(mean)
        if lyapunov is not None:                                                 # :107-112
            v_fn, lv = lyapunov.lyapunov_function, lyapunov._lipschitz_lyapunov
            decrease = v_fn.torch(mean) - v_fn.torch(states)
            if err is not None:
                lv_mu = lv.torch(mean) if isinstance(lv, Function) else float(lv)
                decrease = decrease + (lv_mu * err).sum(dim=1, keepdim=True)
            threshold = dev.to_device(np.broadcast_to(
                lyapunov.threshold(states.detach().cpu().numpy()), (states.shape[0], 1)).copy())
            updated = updated - lagrange_multiplier * (decrease - threshold)
        return updated

    # ------------------------------------------------------------------ fused sweeps
    def bellman_descriptor(self, fixed_action=None):
        cfg = nat.SlbBellman()
        cfg.grid = self._grid.descriptor()
        if fixed_action is None:
            cfg.policy = self.policy.descriptor()
        else:
            action = np.atleast_1d(np.asarray(fixed_action, dtype=np.float64))
            cfg.fixed_action = 1
            cfg.policy.out_dim = len(action)
            for i, a in enumerate(action):
                cfg.action[i] = float(a)
        if isinstance(self.dynamics, (FunctionStack, GaussianProcess)):
            cfg.gp = self.dynamics.gp_stack()
        elif isinstance(self.dynamics, UncertainFunction) or not isinstance(self.dynamics, Function):
            raise TypeError("dynamics must be a fusable Function, GaussianProcess or FunctionStack")
        else:
            cfg.dynamics = self.dynamics.descriptor()
        cfg.reward = self.reward_function.descriptor()
        cfg.value = self.value_function.descriptor()
        cfg.gamma = float(self.gamma)
        return cfg

    def _sweep_device(self):
        """One Jacobi sweep over this rank's slab -> full new vertex table on the device."""
        lib = nat.load()
        cfg = self.bellman_descriptor()
        n = self._end - self._begin
        slab = dev.empty((n,))
        nat.check(lib.slb_bellman_sweep(dev.stream(), cfg, self._begin, self._end,
                                        slab.data_ptr()), "slb_bellman_sweep")
        return self._gather(slab)

    def _gather(self, slab):
        rank, world = dev.dist_info()
        if world == 1:
            return slab
        import torch.distributed as dist
        n = self._grid.nindex
        per = -(-n // world)
        padded = torch.zeros(per, dtype=slab.dtype, device=slab.device)
        padded[:slab.numel()] = slab
        out = torch.empty(per * world, dtype=slab.dtype, device=slab.device)
        dist.all_gather_into_tensor(out, padded)      # the one collective per sweep
        return out[:n]

    def value_iteration(self):
        """One synchronous value-iteration sweep (``:135-140``): every vertex is updated from
        the OLD table, then the table is replaced.  Returns ``max |V_new - V_old|`` (the
        convergence test user code applies, ``tests/test_rl.py:66-69``)."""
        lib = nat.load()
        tri = _triangulation_of(self.value_function)
        scale = 1.0
        base = self.value_function
        while isinstance(base, ScaledFunction):
            scale *= base.factor
            base = base.fun
        new = self._sweep_device()
        if scale != 1.0:
            new = new / scale            # the table stores the un-scaled vertex values
        old = tri._param_dev
        if old.shape[1] != 1:
            raise ValueError("value function must have one output")
        residual = dev.zeros((1,))
        nat.check(lib.slb_max_abs_diff(dev.stream(), new.data_ptr(), old.data_ptr(),
                                       new.numel(), residual.data_ptr()), "slb_max_abs_diff")
        tri._param_dev = new.reshape(-1, 1).contiguous()
        return float(residual.item())

    def discrete_policy_optimization(self, action_space, constraint=None):
        """Greedy policy over a discrete action set (``:213-279``): the piecewise-linear
        policy's vertex values become ``action_space[argmax_a future_values(x, a)]``."""
        lib = nat.load()
        policy_tri = _triangulation_of(self.policy)
        grid = policy_tri.discretization
        if grid.nindex != self._grid.nindex or np.any(grid.num_points != self._grid.num_points) \
                or np.any(grid.limits != self._grid.limits):
            raise NotImplementedError("policy and value function must share one discretization")
        actions = np.atleast_2d(np.asarray(action_space, dtype=np.float64))
        n_opt, m = actions.shape
        cfg = self.bellman_descriptor(fixed_action=actions[0])
        n = self._end - self._begin
        cons_dev = None
        if constraint is not None:
            n_states = grid.nindex
            rows = []
            for action in actions:
                arr = np.broadcast_to(action, (n_states, m))
                rows.append(np.asarray(constraint(arr), dtype=np.float64).reshape(-1)
                            [self._begin:self._end])
            cons_dev = dev.to_device(np.stack(rows))
        actions_dev = dev.to_device(actions)
        best = dev.empty((n,), torch.int32)
        # factored path (csrc/bellman_tile.cu): one kernel row per state + a tensor-core contraction
        # against the per-action table, when the library says it applies (workspace > 0)
        need = int(lib.slb_bellman_argmax_workspace(cfg, n_opt)) if self.factor_actions else 0
        scratch = dev.empty((need // 8 + 1,)) if need else None
        nat.check(lib.slb_bellman_argmax(dev.stream(), cfg, self._begin, self._end,
                                         actions_dev.data_ptr(), n_opt, dev.ptr(cons_dev),
                                         best.data_ptr(), None, dev.ptr(scratch)),
                  "slb_bellman_argmax")
        chosen = actions_dev[best.to(torch.int64)]            # [n, m]
        if m != 1 and dev.dist_info()[1] > 1:
            raise NotImplementedError("multi-GPU policy optimisation supports m == 1")
        full = self._gather(chosen[:, 0].contiguous()).reshape(-1, 1) if m == 1 else chosen
        policy_tri.parameters = full
        return full
