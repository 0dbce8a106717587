"""Drop-in mirror of the reference's BinaryPolarEncoderDecoder (BinaryPolarEncoderDecoder.py:15-325).

Same constructor, `encode` / `decode` signatures, return conventions and assertion behaviour; the
recursion and the BinaryMemorylessVectorDistribution arithmetic run in the sm_100a kernels behind the
C-ABI (polarcub_b200/csrc).  Batched entry points (`encode_batch`, `decode_batch`,
`decode_symbols_batch`) are the ones to use for throughput; the single-frame methods exist so the
reference's Monte-Carlo drivers (encodeDecodeSimulation, :328-387) run unchanged on top of this class.
"""
import random

import numpy as np
import torch

from . import engine
from ._lib import PolarcubError
from .CollectionOfBinaryTrellises import CollectionOfBinaryTrellises


def _probs_of(vd, length, cols):
    """Accept any VectorDistribution-like object exposing `.probs` [length, cols] (the reference's own
    BinaryMemorylessVectorDistribution works unchanged), or a plain array."""
    p = getattr(vd, "probs", vd)
    p = np.asarray(p, dtype=np.float64)
    assert p.shape == (length, cols)
    return p


def is_uniform_prior(xprobs):
    """A prior whose rows are (c, c) makes every leaf marginal exactly 0.5 in the reference's arithmetic
    (f: 2c^2 on both sides, max-normalised to (1,1); (0,0) after underflow also yields [0.5,0.5],
    BinaryMemorylessVectorDistribution.py:62-67), so frozen bits do not depend on the data."""
    return bool(np.all(xprobs[:, 0] == xprobs[:, 1]) and np.all(np.isfinite(xprobs)) and np.all(xprobs >= 0))


class BinaryPolarEncoderDecoder:
    def __init__(self, length, frozenSet, commonRandomnessSeed):
        self.commonRandomnessSeed = commonRandomnessSeed
        self.frozenSet = frozenSet
        self.length = length
        n = int(length).bit_length() - 1
        assert length >= 1 and (1 << n) == length, "length must be a power of two"
        self.n = n
        self.initializeFrozenOrInformationAndRandomlyGeneratedNumbers()

    # BinaryPolarEncoderDecoder.py:24-44
    def initializeFrozenOrInformationAndRandomlyGeneratedNumbers(self):
        mask = np.zeros(self.length, dtype=np.uint8)
        for i in range(self.length):
            if i in self.frozenSet:
                mask[i] = 1
        self.frozenMask = mask
        self.k = int(self.length - mask.sum())
        self.randomlyGeneratedNumbers = np.empty(self.length)
        if self.commonRandomnessSeed != -1:
            rng = random.Random()
            rng.seed(self.commonRandomnessSeed)  # CPython MT19937 -- must be the stdlib generator
            for i in range(self.length):
                self.randomlyGeneratedNumbers[i] = rng.random()
        else:
            self.randomlyGeneratedNumbers[:] = 1.0
        # uniform prior: marginal is exactly 0.5, frozen u_i = 0 iff 0.5 >= r_i (:258-262)
        self.frozenValues = np.where(0.5 >= self.randomlyGeneratedNumbers, 0, 1).astype(np.uint8)
        self._plan = None

    @property
    def plan(self):
        if self._plan is None:
            self._plan = engine.Plan(2, self.n, self.frozenMask, self.frozenValues)
        return self._plan

    def _require_uniform(self, xVectorDistribution):
        xp = _probs_of(xVectorDistribution, self.length, 2)
        if not is_uniform_prior(xp):
            raise PolarcubError("non-uniform a-priori distributions (data-dependent frozen bits) are not "
                                "implemented in the CUDA path yet; there is no CPU fallback")

    # ---- batched entry points ---------------------------------------------------------------------
    def encode_batch(self, information, xVectorDistribution=None):
        """information [B, k] of 0/1 (numpy) -> codewords int64 [B, N]."""
        if xVectorDistribution is not None:
            self._require_uniform(xVectorDistribution)
        info = np.asarray(information)
        assert info.ndim == 2 and info.shape[1] == self.k
        packed = engine.pack_bits(info) if self.k else np.zeros((info.shape[0], 0), dtype=np.uint32)
        dev = torch.from_numpy(packed.view(np.int32)).to(self.plan.device)
        cw = engine.encode_bits(self.plan, dev.contiguous())
        return engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64)

    def decode_batch(self, xyProbs, xVectorDistribution=None):
        """xyProbs float64 [B, N, 2] (numpy or device tensor) -> (codewords int64 [B, N], information int64 [B, k])."""
        if xVectorDistribution is not None:
            self._require_uniform(xVectorDistribution)
        xy = xyProbs if torch.is_tensor(xyProbs) else torch.from_numpy(np.ascontiguousarray(xyProbs, dtype=np.float64))
        assert xy.shape[1:] == (self.length, 2)
        cw, info = engine.sc_decode_probs(self.plan, xy.to(self.plan.device).contiguous())
        return (engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64),
                engine.unpack_bits(info.cpu().numpy(), self.k).astype(np.int64))

    def decode_symbols_batch(self, y, table):
        """y uint8 [B, N] channel output symbols, table [Y, 2] = BinaryMemorylessDistribution.probs."""
        yt = y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.uint8))
        cw, info = engine.sc_decode_symbols(self.plan, yt.to(self.plan.device).contiguous(), table)
        return (engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64),
                engine.unpack_bits(info.cpu().numpy(), self.k).astype(np.int64))

    def decode_trellis_batch(self, collection, want_collapse=False):
        """Deletion channel: `collection` is the descriptor built by
        CollectionOfBinaryTrellises.buildCollection[Batch]_uniformInput_deletion -> (codewords int64 [B, N],
        information int64 [B, k]) (+ the first collapsed vector float64 [B, T, 2] with want_collapse)."""
        assert isinstance(collection, CollectionOfBinaryTrellises) and len(collection) == self.length
        if collection.n0 == 0:  # trellises of one symbol are not a reference use (main_deletion.py:74 takes n0 >= 1)
            raise PolarcubError("n0 = 0 is not supported by the CUDA trellis path")
        dev = self.plan.device
        bits = torch.from_numpy(collection.sub_bits).to(dev)
        lens = torch.from_numpy(collection.sub_len).to(dev)
        out = engine.trellis_decode(self.plan, collection.n0, collection.deletionProb, collection.ones, bits, lens,
                                    want_collapse=want_collapse)
        res = (engine.unpack_bits(out[0].cpu().numpy(), self.length).astype(np.int64),
               engine.unpack_bits(out[1].cpu().numpy(), self.k).astype(np.int64))
        return res + (out[2].cpu().numpy(),) if want_collapse else res

    # ---- genie runs (BinaryPolarEncoderDecoder.py:101-221): all indices frozen, per-leaf probabilities captured -----------
    def _genie_u(self, seeds):
        """u vectors of genie runs: every index frozen, u_i = 0 iff 0.5 >= r_i with r from random.Random(seed) (:24-44,
        :258-262; uniform prior)."""
        u = np.empty((len(seeds), self.length), dtype=np.uint8)
        for t, seed in enumerate(seeds):
            if seed == -1:
                u[t] = 1
            else:
                rng = random.Random()
                rng.seed(seed)
                u[t] = [0 if 0.5 >= rng.random() else 1 for _ in range(self.length)]
        return u

    def genie_encode_batch(self, xVectorDistribution, seeds):
        """genieSingleEncodeSimulatioan for a list of seeds -> (encodedVectors int64 [B, N], TV [B, N], H [B, N]).
        With the uniform prior every captured marginal is exactly [0.5, 0.5]: TV = 0, H = eta(0.5) + eta(0.5) = 1."""
        self._require_uniform(xVectorDistribution)
        u = self._genie_u(seeds)
        packed = torch.from_numpy(engine.pack_bits(u).view(np.int32)).to(self.plan.device)
        cw = engine.polar_transform_bits(self.n, packed.contiguous())  # x = u B_N F^(x)n; the map is an involution
        enc = engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64)
        B = len(seeds)
        return enc, np.zeros((B, self.length)), np.ones((B, self.length))

    def genie_decode_batch(self, xVectorDistribution, xy, seeds, trustXYProbs=True, return_marginals=False):
        """genieSingleDecodeSimulatioan for a batch: `xy` float64 [B, N, 2] or a trellis-collection descriptor, one genie
        seed per frame -> (decodedVectors int64 [B, N], Pe [B, N], H [B, N] or None)."""
        from .simulation import eta
        self._require_uniform(xVectorDistribution)
        u = self._genie_u(seeds)
        dev = self.plan.device
        up = torch.from_numpy(engine.pack_bits(u).view(np.int32)).to(dev).contiguous()
        if isinstance(xy, CollectionOfBinaryTrellises):
            assert xy.frames == len(seeds) and len(xy) == self.length
            if xy.n0 == 0:
                raise PolarcubError("n0 = 0 is not supported by the CUDA trellis path")
            if xy.n == xy.n0:
                raise PolarcubError("genie runs over a single trellis (n == n0) are not supported by the CUDA path")
            cw, marg = engine.trellis_genie(self.plan, xy.n0, xy.deletionProb, xy.ones, torch.from_numpy(xy.sub_bits).to(dev),
                                            torch.from_numpy(xy.sub_len).to(dev), up)
        else:
            x = xy if torch.is_tensor(xy) else torch.from_numpy(np.ascontiguousarray(xy, dtype=np.float64))
            assert x.shape == (len(seeds), self.length, 2)
            cw, marg = engine.sc_genie_probs(self.plan, x.to(dev).contiguous(), up)
        dec = engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64)
        marg = marg.cpu().numpy()
        if trustXYProbs:  # :153-157
            Pe = np.minimum(marg[..., 0], marg[..., 1])
            flat = marg.reshape(-1, 2)
            H = np.array([eta(a) + eta(b) for a, b in flat.tolist()]).reshape(marg.shape[:2])  # math.log2, as the reference
        else:  # :158-171: compare the probability of the actual u_i (polarTransformOfBits of the decoded vector) with the other
            d = u.astype(np.int64)  # the decoded vector is the codeword of u
            pd = np.take_along_axis(marg, d[..., None], axis=2)[..., 0]
            po = np.take_along_axis(marg, 1 - d[..., None], axis=2)[..., 0]
            Pe = np.where(pd > po, 0.0, np.where(pd == po, 0.5, 1.0))
            H = None
        return (dec, Pe, H, marg) if return_marginals else (dec, Pe, H)

    def genieSingleDecodeSimulatioan(self, xVectorDistribution, xyVectorDistribution, genieSingleRunSeed, trustXYProbs):
        """BinaryPolarEncoderDecoder.py:114-178 -> (decodedVector, Pevec, Hvec)."""
        assert len(xVectorDistribution) == self.length
        xy = xyVectorDistribution
        if not isinstance(xy, CollectionOfBinaryTrellises):
            xy = _probs_of(xy, self.length, 2).reshape(1, self.length, 2)
        dec, Pe, H = self.genie_decode_batch(xVectorDistribution, xy, [genieSingleRunSeed], trustXYProbs)
        return dec[0], list(Pe[0]), (list(H[0]) if trustXYProbs else [])

    def genieSingleEncodeSimulatioan(self, xVectorDistribution, genieSingleRunSeed):
        """BinaryPolarEncoderDecoder.py:180-221 -> (encodedVector, TVvec, Hvec)."""
        assert len(xVectorDistribution) == self.length
        enc, TV, H = self.genie_encode_batch(xVectorDistribution, [genieSingleRunSeed])
        return enc[0], list(TV[0]), list(H[0])

    # ---- the reference's entry points -----------------------------------------------------------------
    def encode(self, xVectorDistribution, information):
        """BinaryPolarEncoderDecoder.py:46-69 -> encodedVector int64 [N]."""
        assert len(xVectorDistribution) == self.length
        assert len(information) == self.k
        self._require_uniform(xVectorDistribution)
        info = np.asarray(information, dtype=np.int64).reshape(1, self.k)
        return self.encode_batch(info)[0]

    def decode(self, xVectorDistribution, xyVectorDistribution):
        """BinaryPolarEncoderDecoder.py:71-99 -> (encodedVector int64 [N], information int64 [k])."""
        assert len(xVectorDistribution) == len(xyVectorDistribution) == self.length
        self._require_uniform(xVectorDistribution)
        if isinstance(xyVectorDistribution, CollectionOfBinaryTrellises):  # main_deletion.py:52-59
            assert xyVectorDistribution.frames == 1
            cw, info = self.decode_trellis_batch(xyVectorDistribution)
            return cw[0], info[0]
        xy = _probs_of(xyVectorDistribution, self.length, 2).reshape(1, self.length, 2)
        cw, info = self.decode_batch(xy)
        return cw[0], info[0]


def polarTransformOfBits(xvec):
    """BinaryPolarEncoderDecoder.py:494-516 (x -> u), on the GPU."""
    x = np.asarray(xvec, dtype=np.int64)
    N = x.shape[0]
    n = int(N).bit_length() - 1
    assert (1 << n) == N
    packed = torch.from_numpy(engine.pack_bits(x.reshape(1, N)).view(np.int32)).cuda()
    u = engine.polar_transform_bits(n, packed.contiguous())
    return list(int(b) for b in engine.unpack_bits(u.cpu().numpy(), N)[0])
