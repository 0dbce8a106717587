"""Host-side pieces of VectorDistributions/BinaryTrellis.py that the reference's drivers call directly.

The trellis arithmetic itself (construction :309-438, transforms :206-258, normalisation :280-306) runs on the GPU
(csrc/trellis.cu); this module only mirrors the channel simulator."""
import random


def deletionChannelSimulation(codeword, p, seed, randomNumberGenerator=None):
    """VectorDistributions/BinaryTrellis.py:441-461: i.i.d. deletions with probability p, one RNG draw per symbol."""
    if randomNumberGenerator is not None:
        assert seed is None
    else:
        if seed is None:
            seed = 200
        randomNumberGenerator = random.Random()
        randomNumberGenerator.seed(seed)
    return [codeword[i] for i in range(len(codeword)) if not randomNumberGenerator.random() < p]
