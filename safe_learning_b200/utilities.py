"""Helpers kept from ``safe_learning/utilities.py``: ``batchify`` (``:224-249``, defines the
reference's batch semantics), ``dlqr`` / ``lqr`` (``:300-356``, fixture generation)."""

import numpy as np
import scipy.linalg

__all__ = ["batchify", "dlqr", "lqr", "concatenate_inputs"]

from .functions import concatenate_inputs  # noqa: E402,F401


def batchify(arrays, batch_size):
    """Yield ``(start, [views])`` in order; the last batch may be short."""
    if not isinstance(arrays, (list, tuple)):
        arrays = (arrays,)
    start = 0
    while True:
        views = [arr[start:start + batch_size] for arr in arrays]
        if views[0].size == 0:
            return
        yield start, views
        start += batch_size


def dlqr(a, b, q, r):
    """Discrete-time LQR: returns (k, p) with u = -k x."""
    a, b, q, r = (np.atleast_2d(m) for m in (a, b, q, r))
    p = scipy.linalg.solve_discrete_are(a, b, q, r)
    btp = b.T.dot(p)
    return np.linalg.solve(btp.dot(b) + r, btp.dot(a)), p


def lqr(a, b, q, r):
    """Continuous-time LQR: returns (k, p)."""
    a, b, q, r = (np.atleast_2d(m) for m in (a, b, q, r))
    p = scipy.linalg.solve_continuous_are(a, b, q, r)
    return np.linalg.solve(r, b.T.dot(p)), p
