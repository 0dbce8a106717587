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

    def _require_linear(self, what):
        if self.use_log:
            raise PolarcubError(what + " is a linear-domain entry point; with use_log=True pass log-probabilities to "
                                "decode_batch / listDecode_batch (there is no CPU fallback)")

    # ---- batched --------------------------------------------------------------------------------------
    def encode_batch(self, information):
        info = np.ascontiguousarray(information, dtype=np.uint8)
        assert info.ndim == 2 and info.shape[1] == self.k
        cw = engine.qsc_encode(self.plan, torch.from_numpy(info).to(self.plan.device).contiguous())
        return cw.cpu().numpy().astype(np.int64)

    def decode_batch(self, xyProbs, return_codeword=False):
        """xyProbs [B, N, q] float64: probabilities, or with use_log=True their natural logarithms (-inf for 0), the values
        QaryMemorylessVectorDistribution.probs holds in that mode (QaryMemorylessDistribution.py:764-765)."""
        xy = xyProbs if torch.is_tensor(xyProbs) else torch.from_numpy(np.ascontiguousarray(xyProbs, dtype=np.float64))
        assert xy.shape[1:] == (self.length, self.q)
        cw, info = engine.qsc_decode_probs(self.plan, xy.to(self.plan.device).contiguous(), use_log=self.use_log)
        info = info.cpu().numpy().astype(np.int64)
        if return_codeword:
            return cw.cpu().numpy().astype(np.int64), info
        return info

    def decode_symbols_batch(self, y, table, return_codeword=False):
        """y uint8 [B, N] channel output symbols, table [Y, q] = QaryMemorylessDistribution.probs: decode fused with
        makeQaryMemorylessVectorDistribution(length, yvec, use_log) (QaryMemorylessDistribution.py:757-766).  `table` is
        always the LINEAR table; with use_log=True its logarithm is taken here with math.log as the reference does."""
        yt = y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.uint8))
        if self.use_log:
            import math
            lin = np.asarray(table, dtype=np.float64)
            table = np.array([[math.log(v) if v != 0 else -math.inf for v in row] for row in lin], dtype=np.float64)
        cw, info = engine.qsc_decode_symbols(self.plan, yt.to(self.plan.device).contiguous(), table, use_log=self.use_log)
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


    # ---- SC-list decoding --------------------------------------------------------------------------
    def listDecode_batch(self, xyProbs, frozenValues, maxListSize, actualInformation, return_list=False):
        """Batched listDecode with genie selection.  xyProbs [B,N,q] float64, frozenValues [B,N-k],
        actualInformation [B,k].  Returns (information int64 [B,k], ProbResult values int32 [B]) and, with
        return_list, a dict with the final list (sizes, normalised metrics, genie metric, per-path information).
        With use_log=True xyProbs and the returned metrics are natural logarithms (pc_scl_decode_logprobs)."""
        dev = self.plan.device
        xy = xyProbs if torch.is_tensor(xyProbs) else torch.from_numpy(np.ascontiguousarray(xyProbs, dtype=np.float64))
        B = xy.shape[0]
        if self.q == 2 and 1 <= self.n <= 13 and not self.use_log:
            return self._list_decode_packed(B, frozenValues, maxListSize, actualInformation, return_list,
                                            xy=xy.to(dev).contiguous())
        fv = torch.from_numpy(np.ascontiguousarray(frozenValues, dtype=np.uint8).reshape(B, self.length - self.k))
        ai = torch.from_numpy(np.ascontiguousarray(actualInformation, dtype=np.uint8).reshape(B, self.k))
        out = engine.scl_decode_probs(self.plan, int(maxListSize), xy.to(dev).contiguous(), fv.to(dev).contiguous(),
                                      ai.to(dev).contiguous(), want_list=return_list, want_list_info=return_list,
                                      use_log=self.use_log)
        info = out["info"].cpu().numpy().astype(np.int64)
        res = out["prob_result"].cpu().numpy()
        if not return_list:
            return info, res
        lst = {"list_size": out["list_size"].cpu().numpy(), "list_prob": out["list_prob"].cpu().numpy(),
               "actual_prob": out["actual_prob"].cpu().numpy(),
               "list_info": out["list_info"].cpu().numpy().astype(np.int64)}
        return info, res, lst

    def listDecode_symbols_batch(self, y, table, frozenValues, maxListSize, actualInformation, return_list=False):
        """Binary listDecode of channel OUTPUT SYMBOLS: y uint8 [B,N] with table [Y,2] = the rows
        makeQaryMemorylessVectorDistribution(length, y) copies into probs (QaryMemorylessDistribution.py:757-776), fused
        into the decoder (pc_scl_decode_symbols).  Same returns as listDecode_batch."""
        self._require_linear("listDecode_symbols_batch")
        if self.q != 2:
            raise PolarcubError("listDecode_symbols_batch is binary (q = 2)")
        yt = y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.uint8))
        table = np.ascontiguousarray(table, dtype=np.float64)
        if not torch.is_tensor(y) and yt.numel() and int(yt.max()) >= table.shape[0]:
            raise PolarcubError("channel symbol %d outside the %d-row table" % (int(yt.max()), table.shape[0]))
        return self._list_decode_packed(yt.shape[0], frozenValues, maxListSize, actualInformation, return_list,
                                        y=yt.to(self.plan.device).contiguous(), table=table)

    def _list_decode_packed(self, B, frozenValues, maxListSize, actualInformation, return_list, **chan):
        """pc_scl_decode_packed: bit-packed side buffers; all-zero frozen values travel as NULL."""
        dev = self.plan.device
        nf = self.length - self.k
        fvn = np.ascontiguousarray(frozenValues, dtype=np.uint8).reshape(B, nf)
        fvp = None
        if nf and fvn.any():
            fvp = torch.from_numpy(engine.pack_bits(fvn).view(np.int32)).to(dev)
        ain = np.ascontiguousarray(actualInformation, dtype=np.uint8).reshape(B, self.k)
        aip = torch.from_numpy(engine.pack_bits(ain).view(np.int32).reshape(B, -1)).to(dev)
        if aip.shape[1] == 0:
            aip = torch.zeros((B, 1), dtype=torch.int32, device=dev)
        out = engine.scl_decode_packed(self.plan, int(maxListSize), aip, frozen_packed=fvp, want_list=return_list,
                                       want_list_info=return_list, **chan)
        info = engine.unpack_bits(out["info_packed"].cpu().numpy(), self.k).astype(np.int64)
        res = out["prob_result"].cpu().numpy()
        if (res < 0).any():
            raise PolarcubError("list decoder: a node value left the range of the compressed level format")
        if not return_list:
            return info, res
        lst = {"list_size": out["list_size"].cpu().numpy(), "list_prob": out["list_prob"].cpu().numpy(),
               "actual_prob": out["actual_prob"].cpu().numpy(),
               "list_info": engine.unpack_bits(out["list_info_packed"].cpu().numpy(), self.k).astype(np.int64)}
        return info, res, lst

    def listDecode(self, xyVectorDistribution, frozenValues, maxListSize, check_matrix, check_value,
                   actualInformation=None, verbosity=0):
        """QaryPolarEncoderDecoder.py:118-227 -> (information int64 [k], ProbResult or None).

        With actualInformation (the form ir() uses, :856) the genie selection of :172-213 applies.  Without it the
        reference crashes inside the fast nodes (:569, :608 index self.actualInformation), so that form is not
        reproduced; the first list entry passing `info @ check_matrix % q == check_value` (:215-227) is selected
        from the final list instead, with a dummy genie path."""
        assert len(xyVectorDistribution) == self.length
        xy = _probs_of(xyVectorDistribution, self.length, self.q).reshape(1, self.length, self.q)
        fv = np.asarray(frozenValues, dtype=np.int64).reshape(1, -1)
        if actualInformation is not None:
            info, res = self.listDecode_batch(xy, fv, maxListSize, np.asarray(actualInformation).reshape(1, self.k))
            return info[0], ProbResult(int(res[0]))
        _, _, lst = self.listDecode_batch(xy, fv, maxListSize, np.zeros((1, self.k), dtype=np.int64), return_list=True)
        cands = lst["list_info"][0][:int(lst["list_size"][0])]
        cm, cv = np.asarray(check_matrix), np.asarray(check_value)
        for row in cands:
            if np.array_equal(np.matmul(row, cm) % self.q, cv):
                return row, None
        return cands[0], None

    def calculate_syndrome_and_complement(self, u_message):  # QaryPolarEncoderDecoder.py:822-833
        y = np.asarray(polarTransformOfQudits(self.q, u_message), dtype=np.int64).copy()
        w = np.copy(y)
        w[list(self.infoSet)] = 0
        w[list(self.frozenSet)] *= self.q - 1
        w[list(self.frozenSet)] %= self.q
        u = y
        u[list(self.frozenSet)] = 0
        return w, u

    def get_message_info_bits(self, u_message):
        return u_message[list(self.infoSet)]

    def get_message_frozen_bits(self, u_message):
        return u_message[list(self.frozenSet)]

    def ir(self, a, b, make_xyVectorDistribution, list_size=1, check_size=0, verbosity=0):
        """Information reconciliation wrapper, QaryPolarEncoderDecoder.py:841-858."""
        w, u = self.calculate_syndrome_and_complement(a)
        a_key = self.get_message_info_bits(u)
        frozen_values = (self.get_message_frozen_bits(w) * (self.q - 1)) % self.q
        check_matrix = np.random.choice(range(self.q), (self.k, check_size))
        check_value = np.matmul(a_key, check_matrix) % self.q
        b_key, prob_result = self.listDecode(make_xyVectorDistribution(b), frozenValues=frozen_values,
                                             maxListSize=list_size, check_matrix=check_matrix, check_value=check_value,
                                             actualInformation=a_key, verbosity=verbosity)
        return a_key, b_key, prob_result


def polarTransformOfQudits(q, xvec):
    """QaryPolarEncoderDecoder.py:1136-1154 (x -> u).  Integer butterfly; computed with numpy on the host
    because callers use it on single short vectors (the batched inverse lives in the encoder kernel)."""
    x = np.asarray(xvec, dtype=np.int64)
    if x.shape[0] == 1:
        return x
    first = (x[0::2] + x[1::2]) % q
    second = (q - x[1::2]) % q
    return np.concatenate((polarTransformOfQudits(q, first), polarTransformOfQudits(q, second)))


def hamming(x, y):  # QaryPolarEncoderDecoder.py:932-933
    return sum(x != y)


def irSimulation(q, length, simulateChannel, make_xyVectorDistribution, numberOfTrials, frozenSet, maxListSize=1, checkSize=0,
                 commonRandomnessSeed=1, randomInformationSeed=1, use_log=False, verbosity=0, ir_version=1):
    """QaryPolarEncoderDecoder.irSimulation (QaryPolarEncoderDecoder.py:887-930), same arguments and returns
    (frame_error_prob, symbol_error_prob, rate, probResultList), with the `numberOfTrials` calls of ir() -> listDecode run as
    ONE batch on the GPU.  The three random streams are consumed trial by trial in the reference's order (information:
    random.Random(randomInformationSeed).choices; channel: whatever simulateChannel draws from; check matrix:
    np.random.choice), and make_xyVectorDistribution is called once per trial in order, so seeded runs reproduce the
    reference's inputs exactly.  ir_version=2 calls ir2, which passes listDecode's arguments in the wrong order in the
    reference (:864) and cannot run there either."""
    import math
    if ir_version != 1:
        raise PolarcubError("ir_version=2 (ir2) does not run in the reference either (QaryPolarEncoderDecoder.py:864)")
    encDec = QaryPolarEncoderDecoder(q, length, frozenSet, commonRandomnessSeed, use_log=use_log)
    informationRNG = random.Random(randomInformationSeed)
    k = encDec.k
    xy = np.empty((numberOfTrials, length, q), dtype=np.float64)
    fvs = np.empty((numberOfTrials, length - k), dtype=np.int64)
    a_keys = np.empty((numberOfTrials, k), dtype=np.int64)
    checks = []
    for t in range(numberOfTrials):
        a = informationRNG.choices(range(0, q), k=encDec.length)
        b = simulateChannel(a)
        w, u = encDec.calculate_syndrome_and_complement(a)  # ir(), :841-858
        a_keys[t] = encDec.get_message_info_bits(u)
        fvs[t] = (encDec.get_message_frozen_bits(w) * (q - 1)) % q
        check_matrix = np.random.choice(range(q), (k, checkSize))
        checks.append((check_matrix, np.matmul(a_keys[t], check_matrix) % q))
        xy[t] = _probs_of(make_xyVectorDistribution(b), length, q)
    # with actualInformation the selection never looks at the check matrix (:172-213); it is drawn to keep numpy's stream in step
    b_keys, res = encDec.listDecode_batch(xy, fvs, maxListSize, a_keys) if numberOfTrials else (a_keys, np.zeros(0, dtype=np.int32))
    probResultList = [ProbResult(int(r)) for r in res]
    bad = (a_keys != b_keys)
    badKeys = int(bad.any(axis=1).sum())
    badSymbols = int(bad[bad.any(axis=1)].sum())
    assert k == length - len(frozenSet)
    rate = (math.log2(q) * k - math.log2(maxListSize)) / length
    frame_error_prob = badKeys / numberOfTrials
    symbol_error_prob = badSymbols / (numberOfTrials * encDec.length)
    if verbosity:
        print("Rate: ", rate)
        print("Frame error probability = ", badKeys, "/", numberOfTrials, " = ", frame_error_prob)
        print("Symbol error probability = ", badSymbols, "/ (", numberOfTrials, " * ", encDec.length, ") = ", symbol_error_prob)
    return frame_error_prob, symbol_error_prob, rate, probResultList


def encodeDecodeSimulation(q, length, make_xVectorDistribution, make_codeword, simulateChannel, make_xyVectorDistribution,
                           numberOfTrials, frozenSet, commonRandomnessSeed=1, randomInformationSeed=1, verbosity=0):
    """QaryPolarEncoderDecoder.encodeDecodeSimulation (QaryPolarEncoderDecoder.py:935-982; SC, not SCL): all trials encoded
    in one batch, the channel simulated trial by trial in the reference's order, all received words decoded in one batch.
    Prints the reference's summary line; additionally returns the number of misdecoded words (the reference returns None)."""
    xVectorDistribution = make_xVectorDistribution()
    assert len(xVectorDistribution) == length
    encDec = QaryPolarEncoderDecoder(q, length, frozenSet, commonRandomnessSeed)
    informationRNG = random.Random(randomInformationSeed)
    info = np.array([informationRNG.choices(range(0, q), k=encDec.k) for _ in range(numberOfTrials)],
                    dtype=np.int64).reshape(numberOfTrials, encDec.k)
    encoded = encDec.encode_batch(info) if numberOfTrials else np.zeros((0, length), dtype=np.int64)
    xy = np.empty((numberOfTrials, length, q), dtype=np.float64)
    for t in range(numberOfTrials):
        receivedWord = simulateChannel(make_codeword(encoded[t]))
        xy[t] = _probs_of(make_xyVectorDistribution(receivedWord), length, q)
    decoded = encDec.decode_batch(xy) if numberOfTrials else info
    misdecodedWords = int((decoded != info).any(axis=1).sum())
    print("Error probability = ", misdecodedWords, "/", numberOfTrials, " = ", misdecodedWords / numberOfTrials)
    return misdecodedWords
