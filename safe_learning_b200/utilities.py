"""Host helpers kept from ``safe_learning/utilities.py``: ``batchify`` (``:224-249``, defines the
reference's batch semantics), ``dlqr`` / ``lqr`` (``:300-356``), and the small array builders the
callers of the path use: ``combinations`` / ``linearly_spaced_combinations`` (``:252-296``, the
action set of ``discrete_policy_optimization``) and ``unique_rows`` (``:496-516``)."""

import numpy as np
import scipy.linalg

__all__ = ["batchify", "dlqr", "lqr", "concatenate_inputs", "combinations",
           "linearly_spaced_combinations", "unique_rows"]

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


def combinations(arrays):
    """All combinations of the entries of ``arrays`` as rows (last array varies fastest)."""
    return np.array(np.meshgrid(*arrays)).T.reshape(-1, len(arrays))


def linearly_spaced_combinations(bounds, num_samples):
    """Rows of all combinations of ``num_samples`` linearly spaced values within ``bounds``
    (``[(lo, hi), ...]``; ``num_samples`` an integer or one per variable)."""
    bounds = np.atleast_2d(bounds)
    num_samples = np.broadcast_to(num_samples, len(bounds))
    return combinations([np.linspace(b[0], b[1], n) for b, n in zip(bounds, num_samples)])


def unique_rows(array):
    """Unique rows in the order of ``np.unique`` on the raw row bytes (what
    ``perturb_actions`` relies on, ``lyapunov.py:645-649``)."""
    array = np.ascontiguousarray(array)
    dtype = np.dtype((np.void, array.dtype.itemsize * array.shape[1]))
    _, idx = np.unique(array.view(dtype=dtype), return_index=True)
    return array[idx]
