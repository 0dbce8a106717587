"""Data holder mirroring VectorDistributions/QaryMemorylessVectorDistribution.py:9-24."""
import math

import numpy as np


class QaryMemorylessVectorDistribution:
    def __init__(self, q, length, use_log=False):
        assert q > 1
        assert length > 0
        self.q = q
        self.probs = np.empty((length, q), dtype=np.float64)
        self.probs[:] = np.nan
        self.length = length
        self.use_log = use_log
        self.default_marginal_probs = [-math.log(q)] * q if use_log else [1 / q] * q

    def __len__(self):
        return self.length
