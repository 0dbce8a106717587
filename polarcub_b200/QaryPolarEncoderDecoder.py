"""Drop-in mirror of the reference's QaryPolarEncoderDecoder (QaryPolarEncoderDecoder.py:26-401).

`encode` returns encodedVector int64 [N] (:65-88), `decode` returns only information int64 [k] (:90-116).
Frozen symbols are 0 (:351).  The recursion and the QaryMemorylessVectorDistribution arithmetic run in
polarcub_b200/csrc/sc_qary.cu / encode.cu.
"""
import random
from enum import Enum

import numpy as np
import torch

from . import engine
from ._lib import PolarcubError
from .BinaryPolarEncoderDecoder import _probs_of


class ProbResult(Enum):  # QaryPolarEncoderDecoder.py:18-24
    SuccessActualIsMax = 0
    SuccessActualSmallerThanMax = 1
    FailActualLargerThanMax = 2
    FailActualIsMax = 3
    FailActualWithinRange = 4
    FailActualSmallerThanMin = 5


class QaryPolarEncoderDecoder:
    def __init__(self, q, length, frozenSet, commonRandomnessSeed, use_log=False):
        self.q = q
        self.commonRandomnessSeed = commonRandomnessSeed
        self.frozenSet = sorted(frozenSet)
        fs = set(self.frozenSet)
        self.infoSet = sorted(i for i in range(length) if i not in fs)
        self.length = length
        n = int(length).bit_length() - 1
        assert length >= 1 and (1 << n) == length, "length must be a power of two"
        self.n = n
        self.k = length - len(self.frozenSet)
        mask = np.zeros(length, dtype=np.uint8)
        if self.frozenSet:
            mask[self.frozenSet] = 1
        self.frozenMask = mask
        self.randomlyGeneratedNumbers = self.initRandomlyGeneratedNumbers()
        self.use_log = use_log
        self.prob_list = None
        self.actual_prob = None
        self._plan = None

    def initRandomlyGeneratedNumbers(self):  # QaryPolarEncoderDecoder.py:54-59 (unused by decode)
        if self.commonRandomnessSeed != -1:
            rng = random.Random(self.commonRandomnessSeed)
            return np.array([rng.random() for _ in range(self.length)])
        return np.full(self.length, 1.0)

    def reinitRandomlyGeneratedNumbers(self, newSeed):
        self.commonRandomnessSeed = newSeed
        self.randomlyGeneratedNumbers = self.initRandomlyGeneratedNumbers()

    @property
    def plan(self):
        if self._plan is None:
            self._plan = engine.Plan(self.q, self.n, self.frozenMask, None)
        return self._plan

    def _require_linear(self):
        if self.use_log:
            raise PolarcubError("use_log=True (log-domain arithmetic) is not implemented in the CUDA path yet; "
                                "there is no CPU fallback")

    # ---- batched --------------------------------------------------------------------------------------
    def encode_batch(self, information):
        info = np.ascontiguousarray(information, dtype=np.uint8)
        assert info.ndim == 2 and info.shape[1] == self.k
        cw = engine.qsc_encode(self.plan, torch.from_numpy(info).to(self.plan.device).contiguous())
        return cw.cpu().numpy().astype(np.int64)

    def decode_batch(self, xyProbs, return_codeword=False):
        self._require_linear()
        xy = xyProbs if torch.is_tensor(xyProbs) else torch.from_numpy(np.ascontiguousarray(xyProbs, dtype=np.float64))
        assert xy.shape[1:] == (self.length, self.q)
        cw, info = engine.qsc_decode_probs(self.plan, xy.to(self.plan.device).contiguous())
        info = info.cpu().numpy().astype(np.int64)
        if return_codeword:
            return cw.cpu().numpy().astype(np.int64), info
        return info

    # ---- the reference's entry points -----------------------------------------------------------------
    def encode(self, xVectorDistribution, information):
        assert len(xVectorDistribution) == self.length
        assert len(information) == self.k
        return self.encode_batch(np.asarray(information, dtype=np.int64).reshape(1, self.k))[0]

    def decode(self, xVectorDistribution, xyVectorDistribution):
        assert len(xVectorDistribution) == len(xyVectorDistribution) == self.length
        xy = _probs_of(xyVectorDistribution, self.length, self.q).reshape(1, self.length, self.q)
        return self.decode_batch(xy)[0]


def polarTransformOfQudits(q, xvec):
    """QaryPolarEncoderDecoder.py:1136-1154 (x -> u).  Integer butterfly; computed with numpy on the host
    because callers use it on single short vectors (the batched inverse lives in the encoder kernel)."""
    x = np.asarray(xvec, dtype=np.int64)
    if x.shape[0] == 1:
        return x
    first = (x[0::2] + x[1::2]) % q
    second = (q - x[1::2]) % q
    return np.concatenate((polarTransformOfQudits(q, first), polarTransformOfQudits(q, second)))
