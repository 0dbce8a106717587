"""Deletion-channel input adapter -- mirror of buildCollectionOfBinaryTrellises_uniformInput_deletion
(VectorDistributions/CollectionOfBinaryTrellises.py:106-129).

The reference builds 2^(n-n0) BinaryTrellis objects on the host (dicts of vertices and edges) and hands the collection to
BinaryPolarEncoderDecoder.decode as the xyVectorDistribution.  Here the collection is a DESCRIPTOR: the trimmed sub-words
as fixed-width arrays plus the channel parameters; trellis construction (BinaryTrellis.py:309-438) and every trellis
transform run on the GPU inside pc_trellis_decode (csrc/trellis.cu).  `BinaryPolarEncoderDecoder.decode` accepts the
descriptor wherever the reference accepts a CollectionOfBinaryTrellises.
"""
import numpy as np

from . import Guardbands


class CollectionOfBinaryTrellises:
    """Descriptor of B collections (B = 1 for the reference's single-frame call)."""

    def __init__(self, sub_bits, sub_len, deletionProb, n, n0, ones):
        self.sub_bits = np.ascontiguousarray(sub_bits, dtype=np.uint8)  # [B, T, maxlen]
        self.sub_len = np.ascontiguousarray(sub_len, dtype=np.int32)    # [B, T]
        self.deletionProb = float(deletionProb)
        self.n, self.n0, self.ones = int(n), int(n0), int(ones)
        self.length = 1 << self.n
        self.numberOfTrellises = 1 << (self.n - self.n0)
        assert self.sub_bits.ndim == 3 and self.sub_bits.shape[:2] == self.sub_len.shape
        assert self.sub_len.shape[1] == self.numberOfTrellises

    def __len__(self):
        return self.length

    @property
    def frames(self):
        return self.sub_len.shape[0]


def buildCollectionOfBinaryTrellises_uniformInput_deletion(receivedWord, deletionProb, xi, n, n0,
                                                           numberOfOnesToAddAtBothEndsOfGuardbands, verbosity=0):
    """Same signature as the reference (`xi` only matters to the transmitter; the receiver trims zeros)."""
    assert 0 <= n0 <= n
    bits, lens = Guardbands.split_batch([receivedWord], n, n0)
    return CollectionOfBinaryTrellises(bits, lens, deletionProb, n, n0, numberOfOnesToAddAtBothEndsOfGuardbands)


def buildCollectionBatch_uniformInput_deletion(receivedWords, deletionProb, xi, n, n0,
                                               numberOfOnesToAddAtBothEndsOfGuardbands):
    """Batched form: one descriptor for a list of received words."""
    assert 0 <= n0 <= n
    bits, lens = Guardbands.split_batch(receivedWords, n, n0)
    return CollectionOfBinaryTrellises(bits, lens, deletionProb, n, n0, numberOfOnesToAddAtBothEndsOfGuardbands)
