"""Configuration singleton (reference: ``safe_learning/configuration.py:8-32``).

``dtype`` / ``np_dtype`` are fp64 like the reference.  ``gp_batch_size`` no longer
bounds a Session.run: the sweep runs as one fused kernel over the whole index range.
It is kept because the batch size leaks into reference results in two places that
are reproduced bit for bit: the ``can_shrink=False`` clearing rule and the ``c_max``
index quirk (``lyapunov.py:583-595``).
"""

import numpy as np


class Configuration(object):
    """Dtype and batch-size knobs."""

    def __init__(self):
        self.dtype = np.float64
        self.gp_batch_size = 10000

    @property
    def np_dtype(self):
        return np.float64

    def __repr__(self):
        lines = ["Configuration parameters:", ""]
        lines += ["{}: {!r}".format(k, v) for k, v in self.__dict__.items()]
        return "\n".join(lines)
