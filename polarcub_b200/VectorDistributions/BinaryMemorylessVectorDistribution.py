"""Data holder mirroring VectorDistributions/BinaryMemorylessVectorDistribution.py:6-13.

Only the container survives on the host: `probs[i][x]` = probability of x at time i.  The transforms
(minusTransform / plusTransform / normalize, reference :15-87) run inside the CUDA decoder.
"""
import numpy as np


class BinaryMemorylessVectorDistribution:
    def __init__(self, length):
        assert length > 0
        self.probs = np.empty((length, 2))
        self.probs[:] = np.nan
        self.length = length

    def __len__(self):
        return self.length
