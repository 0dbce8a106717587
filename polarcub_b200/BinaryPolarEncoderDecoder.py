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
    if not hasattr(vd, "probs") and (hasattr(vd, "trellises") or hasattr(vd, "layers") or type(vd).__name__ in (
            "CollectionOfBinaryTrellises", "BinaryTrellis")):
        # the reference's own CollectionOfBinaryTrellises / BinaryTrellis objects (dict-of-objects trellises) cannot be read
        # by the CUDA path: the collection has to come from this package's builder, which keeps the trimmed sub-words
        raise PolarcubError(
            "trellis inputs must be built with polarcub_b200.CollectionOfBinaryTrellises."
            "buildCollectionOfBinaryTrellises_uniformInput_deletion (the reference-built %s object holds Python trellis "
            "graphs the CUDA decoder cannot ingest); swap the builder import as INTEGRATION.md shows" % type(vd).__name__)
    p = getattr(vd, "probs", vd)
    p = np.asarray(p, dtype=np.float64)
    if p.shape != (length, cols):
        raise AssertionError("vector distribution of shape %s where (%d, %d) is expected" % (p.shape, length, cols))
    return p


def _check_n0(n0):
    """The CUDA trellis path keeps trellises of 2^n0 <= 16 symbols (n0 in [1, 4]); the reference's default n0 = n // 3
    (main_deletion.py:100) is outside that range for n < 3 and n >= 15."""
    if not 1 <= n0 <= 4:
        raise PolarcubError("n0 = %d: the CUDA trellis path supports n0 in [1, 4] (trellises of 2..16 symbols); "
                            "there is no CPU fallback" % n0)


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
            raise PolarcubError("non-uniform a-priori distributions are not supported on this path (trellis inputs); "
                                "there is no CPU fallback")

    def _prior(self, xVectorDistribution):
        """None for a uniform prior (frozen bits do not depend on the data: the fast kernels apply), else its probs [N, 2]."""
        if xVectorDistribution is None:
            return None
        xp = _probs_of(xVectorDistribution, self.length, 2)
        return None if is_uniform_prior(xp) else xp

    @property
    def genie_plan(self):
        """All indices frozen (geniePreSteps, BinaryPolarEncoderDecoder.py:101-107)."""
        if getattr(self, "_gplan", None) is None:
            self._gplan = engine.Plan(2, self.n, np.ones(self.length, dtype=np.uint8), np.zeros(self.length, dtype=np.uint8))
        return self._gplan

    def _rnd_dev(self):
        return torch.from_numpy(np.ascontiguousarray(self.randomlyGeneratedNumbers, dtype=np.float64)).to(self.plan.device)

    # ---- batched entry points ---------------------------------------------------------------------
    def encode_batch(self, information, xVectorDistribution=None):
        """information [B, k] of 0/1 (numpy) -> codewords int64 [B, N]."""
        info = np.asarray(information)
        assert info.ndim == 2 and info.shape[1] == self.k
        xp = self._prior(xVectorDistribution)
        if xp is not None:  # data-dependent frozen bits: walk the a-priori tree (BinaryPolarEncoderDecoder.py:258-262)
            dev = self.plan.device
            u = np.zeros((info.shape[0], self.length), dtype=np.uint8)
            u[:, self.frozenMask == 0] = info
            up = torch.from_numpy(engine.pack_bits(u).view(np.int32)).to(dev).contiguous()
            x = torch.from_numpy(xp).to(dev).expand(info.shape[0], self.length, 2).contiguous()
            cw = engine.sc_encode_prior(self.plan, x, up, self._rnd_dev())
            return engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64)
        packed = engine.pack_bits(info) if self.k else np.zeros((info.shape[0], 0), dtype=np.uint32)
        dev = torch.from_numpy(packed.view(np.int32)).to(self.plan.device)
        cw = engine.encode_bits(self.plan, dev.contiguous())
        return engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64)

    def decode_batch(self, xyProbs, xVectorDistribution=None):
        """xyProbs float64 [B, N, 2] (numpy or device tensor) -> (codewords int64 [B, N], information int64 [B, k])."""
        xy = xyProbs if torch.is_tensor(xyProbs) else torch.from_numpy(np.ascontiguousarray(xyProbs, dtype=np.float64))
        assert xy.shape[1:] == (self.length, 2)
        xp = self._prior(xVectorDistribution)
        if xp is not None:  # a-posteriori and a-priori trees in lock step (BinaryPolarEncoderDecoder.py:277-317)
            dev = self.plan.device
            cw, info = engine.sc_decode_probs_prior(self.plan, xy.to(dev).contiguous(), torch.from_numpy(xp).to(dev), self._rnd_dev())
        else:
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
        _check_n0(collection.n0)
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

    def _genie_rnd(self, seeds):
        """randomlyGeneratedNumbers of each genie run [B, N] (BinaryPolarEncoderDecoder.py:33-44 with the run's seed)."""
        r = np.empty((len(seeds), self.length), dtype=np.float64)
        for t, seed in enumerate(seeds):
            if seed == -1:
                r[t] = 1.0
            else:
                rng = random.Random()
                rng.seed(seed)
                r[t] = [rng.random() for _ in range(self.length)]
        return r

    def genie_encode_batch(self, xVectorDistribution, seeds):
        """genieSingleEncodeSimulatioan for a list of seeds -> (encodedVectors int64 [B, N], TV [B, N], H [B, N]).
        With the uniform prior every captured marginal is exactly [0.5, 0.5]: TV = 0, H = eta(0.5) + eta(0.5) = 1."""
        xp = self._prior(xVectorDistribution)
        if xp is not None:
            from .simulation import eta
            dev = self.plan.device
            B = len(seeds)
            rnd = torch.from_numpy(self._genie_rnd(seeds)).to(dev)
            x = torch.from_numpy(xp).to(dev).expand(B, self.length, 2).contiguous()
            up = torch.zeros((B, self.plan.Nw), dtype=torch.int32, device=dev)
            cw, marg = engine.sc_encode_prior(self.genie_plan, x, up, rnd, want_marg=True)
            marg = marg.cpu().numpy()
            TV = np.abs(marg[..., 0] - marg[..., 1])
            H = np.array([eta(a) + eta(b) for a, b in marg.reshape(-1, 2).tolist()]).reshape(marg.shape[:2])
            return engine.unpack_bits(cw.cpu().numpy(), self.length).astype(np.int64), TV, H
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
        dev = self.plan.device
        xp = self._prior(xVectorDistribution)
        if xp is not None:
            if isinstance(xy, CollectionOfBinaryTrellises):
                self._require_uniform(xVectorDistribution)
            x = xy if torch.is_tensor(xy) else torch.from_numpy(np.ascontiguousarray(xy, dtype=np.float64))
            assert x.shape == (len(seeds), self.length, 2)
            rnd = torch.from_numpy(self._genie_rnd(seeds)).to(dev)
            cw, _, marg, _ = engine.sc_decode_probs_prior(self.genie_plan, x.to(dev).contiguous(), torch.from_numpy(xp).to(dev), rnd,
                                                           want_marg=True)
            u = None
        else:
            u = self._genie_u(seeds)
            up = torch.from_numpy(engine.pack_bits(u).view(np.int32)).to(dev).contiguous()
        if xp is not None:
            pass
        elif isinstance(xy, CollectionOfBinaryTrellises):
            assert xy.frames == len(seeds) and len(xy) == self.length
            _check_n0(xy.n0)
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
        if u is None:  # the u vector behind the decoded codeword (polarTransformOfBits, :159)
            u = engine.unpack_bits(engine.polar_transform_bits(self.n, cw.contiguous()).cpu().numpy(), self.length)
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
        info = np.asarray(information, dtype=np.int64).reshape(1, self.k)
        return self.encode_batch(info, xVectorDistribution)[0]

    def decode(self, xVectorDistribution, xyVectorDistribution):
        """BinaryPolarEncoderDecoder.py:71-99 -> (encodedVector int64 [N], information int64 [k])."""
        assert len(xVectorDistribution) == len(xyVectorDistribution) == self.length
        if isinstance(xyVectorDistribution, CollectionOfBinaryTrellises):  # main_deletion.py:52-59
            self._require_uniform(xVectorDistribution)
            assert xyVectorDistribution.frames == 1
            cw, info = self.decode_trellis_batch(xyVectorDistribution)
            return cw[0], info[0]
        xy = _probs_of(xyVectorDistribution, self.length, 2).reshape(1, self.length, 2)
        cw, info = self.decode_batch(xy, xVectorDistribution)
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
