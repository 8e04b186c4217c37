"""torch-CPU (fp64) variant of the oracle's per-batch graph -- TEST / BENCH INFRASTRUCTURE ONLY.

BASELINE.md section 3.4 asks for two CPU timings of the reference algorithm and reports the
faster: the numpy/scipy restatement (``reference_path.Lyapunov.negative``) and this one, whose
dense kernels (``torch.cdist``-free RBF, ``torch.linalg.solve_triangular``, matmul) run on torch's
own CPU thread pool -- closer to the Eigen kernels TF1 would have used.  It evaluates the same
graph node ``tf_negative`` (``lyapunov.py:436-441``) for the object types of the C2 workload:
saturated linear policy (``functions.py:349-354, 1583``), stacked RBF ``GPRCached`` models with a
linear prior mean (``functions.py:417-458, 507-515, 278-291``), quadratic V (``:1537-1539``) and
``|2 P x|`` as per-dimension Lipschitz function (``lyapunov.py:284-288, 344-347``).
Results agree with the numpy oracle to rounding (checked in ``tests/test_oracle_reference_vectors``).
"""

from __future__ import annotations

import numpy as np
import torch

from . import reference_path as R


class TorchPendulumGraph(object):
    """Pre-converted tensors of an oracle ``Lyapunov`` object of the pendulum workloads."""

    def __init__(self, lyap):
        pol = lyap.policy
        if not (isinstance(pol, R.Saturation) and isinstance(pol.fun, R.LinearSystem)):
            raise TypeError("torch baseline: policy must be Saturation(LinearSystem)")
        if not isinstance(lyap.lyapunov_function, R.QuadraticFunction):
            raise TypeError("torch baseline: V must be a QuadraticFunction")
        lv = lyap._lipschitz_lyapunov
        if not (isinstance(lv, R.AbsFunction) and isinstance(lv.fun, R.LinearSystem)):
            raise TypeError("torch baseline: L_V must be AbsFunction(LinearSystem)")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))  # noqa: E731
        self.K = t(pol.fun.matrix)
        self.lo, self.hi = float(pol.lower), float(pol.upper)
        self.P = t(lyap.lyapunov_function.matrix)
        self.LV = t(lv.fun.matrix)
        self.lf = float(lyap._lipschitz_dynamics)
        self.tau = float(lyap.tau)
        self.gps = []
        for f in lyap.dynamics.functions:
            gp = f.gaussian_process
            kern = gp.kern
            if not isinstance(kern, R.RBF):
                raise TypeError("torch baseline: RBF kernels only")
            ls = np.broadcast_to(np.asarray(kern.lengthscales, dtype=np.float64), (gp.X.shape[1],))
            row = None if gp.mean_function is None else t(gp.mean_function.row)
            self.gps.append(dict(Xs=t(gp.X / ls), ls=t(ls.copy()), var=float(kern.variance),
                                 L=t(gp.cholesky), alpha=t(gp.alpha), s=float(gp._scale),
                                 beta=float(f.beta), row=row))

    def negative(self, states):
        x = torch.from_numpy(np.ascontiguousarray(states, dtype=np.float64))
        u = torch.clamp(x @ self.K.T, self.lo, self.hi)                      # :349-354, :1583
        z = torch.cat((x, u), dim=1)                                         # utilities.py:143
        means, errs = [], []
        for g in self.gps:
            zs = z / g["ls"]
            # gpflow square_dist: -2 X X2^T + |X|^2 + |X2|^2
            d2 = -2.0 * (g["Xs"] @ zs.T) + (g["Xs"] ** 2).sum(1)[:, None] + (zs ** 2).sum(1)[None, :]
            Kx = (g["s"] ** 2) * (g["var"] * torch.exp(-d2 / 2))             # functions.py:438
            a = torch.linalg.solve_triangular(g["L"], Kx, upper=False)       # :441
            mx = 0.0 if g["row"] is None else g["s"] * (z @ g["row"].T)      # :439
            mean = (a.T @ g["alpha"] + mx) / g["s"]                          # :442, :455
            var = ((g["s"] ** 2) * g["var"] - (a * a).sum(0)) / (g["s"] ** 2)   # :450-451, :456
            means.append(mean)
            errs.append(g["beta"] * torch.sqrt(var)[:, None])               # :514
        mu, err = torch.cat(means, dim=1), torch.cat(errs, dim=1)
        v = lambda p: ((p @ self.P) * p).sum(1, keepdim=True)  # noqa: E731   :1537-1539
        bound = (torch.abs(mu @ self.LV.T) * err).sum(1, keepdim=True)       # lyapunov.py:344-347
        decrease = v(mu) - v(x) + bound                                      # :351-352, :376
        lvx = torch.abs(x @ self.LV.T).sum(1, keepdim=True)                  # :284-286
        threshold = -lvx * (1.0 + self.lf) * self.tau                        # :288
        return (decrease < threshold)[:, 0].numpy()                          # :441
