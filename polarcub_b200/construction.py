"""Frozen-set selection from a per-index error-probability vector (host-side, offline).

Mirrors BinaryPolarEncoderDecoder.frozenSetFromTVAndPe (BinaryPolarEncoderDecoder.py:519-548): indices are
sorted by TV+Pe ascending with Python's stable sort (:525); with a uniform prior TV = 0.  `K` fixes the
number of information indices explicitly (the epsilon rule gives K=361 at N=1024, the
BASELINE config names K=512).
"""
import os

import numpy as np

_CONS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "constructions")


def frozen_set_from_pe(pe, K):
    pe = np.asarray(pe, dtype=np.float64)
    order = sorted(range(len(pe)), key=lambda i: pe[i])
    return set(int(i) for i in order[K:])


def frozen_set_from_tv_and_pe(tv, pe, error_upper_bound):
    """The epsilon rule of BinaryPolarEncoderDecoder.py:519-548."""
    s = [float(a) + float(b) for a, b in zip(tv, pe)]
    order = sorted(range(len(s)), key=lambda i: s[i])
    err, idx = 0.0, -1
    while err < error_upper_bound and idx + 1 < len(s):
        i = order[idx + 1]
        if s[i] + err <= error_upper_bound:
            err += s[i]
            idx += 1
        else:
            break
    return set(int(order[j]) for j in range(idx + 1, len(s)))


def bec_pe(n, eps):
    """Closed-form BEC recursion z -> (2z - z^2, z^2) in the reference's MSB-first minus/plus order; Pe_i = z_i / 2."""
    z = [float(eps)]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return 0.5 * np.array(z)


def load_pe(name):
    """Pe vectors produced by the live reference's degrade pass (oracle/gen_constructions.py), committed under
    tests/golden/constructions/."""
    return np.load(os.path.join(_CONS, name))


def ga_awgn_pe(n, sigma):
    """Gaussian-approximation construction for BI-AWGN (Chung et al. phi function, Trifonov's polar use): the mean LLR
    m evolves as minus: phi^-1(1 - (1 - phi(m))^2), plus: 2m; Pe_i = Q(sqrt(m_i / 2)).  NOT a reference algorithm (the
    reference's makeAWGN is an empty stub, QaryMemorylessDistribution.py:800-804): offered as an explicit alternative to
    the degrade-pass vectors under tests/golden/constructions/ (bench.py --construction ga)."""
    import math

    def phi(x):
        if x < 1e-12:
            return 1.0
        if x < 10.0:
            return math.exp(-0.4527 * x ** 0.86 + 0.0218)
        return math.sqrt(math.pi / x) * math.exp(-x / 4.0) * (1.0 - 10.0 / (7.0 * x))

    def phi_inv(y):
        if y >= 1.0:
            return 0.0
        lo, hi = 0.0, 1.0
        while phi(hi) > y and hi < 1e6:
            hi *= 2.0
        for _ in range(80):
            mid = 0.5 * (lo + hi)
            if phi(mid) > y:
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    m = [2.0 / (sigma * sigma)]
    for _ in range(n):
        nxt = []
        for v in m:
            nxt.append(phi_inv(1.0 - (1.0 - phi(v)) ** 2))
            nxt.append(2.0 * v)
        m = nxt
    return np.array([0.5 * math.erfc(math.sqrt(v / 2.0) / math.sqrt(2.0)) for v in m])


def tal_vardy_pe(n, L, xy_probs, threads=None):
    """Pe vector of the reference's degrading construction for a binary-input memoryless channel with a uniform input:
    `dist.minusTransform().degrade(L)` / `dist.plusTransform().degrade(L)` per level and `errorProb()` per leaf
    (ScalarDistributions/BinaryMemorylessDistribution.py:657-677), computed by the native host routine
    `pc_tv_degrade_pe` (csrc/tv_construct.cu) -- float64-identical to the reference, seconds instead of the reference's
    ~11 minutes at N = 1024, L = 100.  xy_probs: [Y, 2] joint probabilities p(y, x) (`BinaryMemorylessDistribution.probs`)."""
    import ctypes
    from . import _lib
    t = np.ascontiguousarray(getattr(xy_probs, "probs", xy_probs), dtype=np.float64)
    assert t.ndim == 2 and t.shape[1] == 2 and t.shape[0] >= 1
    pe = np.empty(1 << n, dtype=np.float64)
    _lib.check(_lib.lib().pc_tv_degrade_pe(int(n), int(L), t.ctypes.data_as(ctypes.c_void_p), int(t.shape[0]),
                                           pe.ctypes.data_as(ctypes.c_void_p), int(threads or os.cpu_count() or 1)),
               "pc_tv_degrade_pe")
    return pe


def calcFrozenSet_degradingUpgrading(n, L, upperBoundOnErrorProbability, xDistribution, xyDistribution, threads=None):
    """calcFrozenSet_degradingUpgrading (ScalarDistributions/BinaryMemorylessDistribution.py:620-680) for the uniform-input
    form (xDistribution=None) -- the only form that runs in the reference (the other branch calls a
    method that does not exist)."""
    assert n >= 0 and L > 0 and upperBoundOnErrorProbability > 0 and xyDistribution is not None
    if xDistribution is not None:
        from ._lib import PolarcubError
        raise PolarcubError("calcFrozenSet_degradingUpgrading: only the uniform-input form (xDistribution=None) is offered")
    pe = tal_vardy_pe(n, L, xyDistribution, threads)
    return frozen_set_from_tv_and_pe([0.0] * len(pe), list(pe), upperBoundOnErrorProbability)


def make_bsc(p):
    """makeBSC (BinaryMemorylessDistribution.py:478-483) as a [2, 2] table."""
    return np.array([[0.5 * (1.0 - p), 0.5 * p], [0.5 * p, 0.5 * (1.0 - p)]])


def make_bec(p):
    """makeBEC (BinaryMemorylessDistribution.py:486-492) as a [3, 2] table."""
    return np.array([[0.5 * (1.0 - p), 0.0], [0.0, 0.5 * (1.0 - p)], [0.5 * p, 0.5 * p]])


def tal_vardy_pe_qary(q, n, L, xy_probs, threads=None):
    """Pe vector of the q-ary degrading construction: `calcTVAndPe_degradingUpgrading(n, L, None, xyDistribution)`
    (ScalarDistributions/QaryMemorylessDistribution.py:934-990, `degrade_dynamic` :218-260) by the native host routine
    `pc_tv_degrade_pe_qary`; float64-identical to the reference.  xy_probs: [Y, q] (`QaryMemorylessDistribution.probs`)."""
    import ctypes
    from . import _lib
    t = np.ascontiguousarray(getattr(xy_probs, "probs", xy_probs), dtype=np.float64)
    assert t.ndim == 2 and t.shape[1] == q and t.shape[0] >= 1
    pe = np.empty(1 << n, dtype=np.float64)
    _lib.check(_lib.lib().pc_tv_degrade_pe_qary(int(q), int(n), int(L), t.ctypes.data_as(ctypes.c_void_p), int(t.shape[0]),
                                                pe.ctypes.data_as(ctypes.c_void_p), int(threads or os.cpu_count() or 1)),
               "pc_tv_degrade_pe_qary")
    return pe


def make_qsc(q, p):
    """makeQSC (QaryMemorylessDistribution.py:780-784): probs[y][x] = 1 - p if x == y else p / (q - 1)."""
    t = np.full((q, q), p / (q - 1))
    np.fill_diagonal(t, 1.0 - p)
    return t


def frozenSetFromTVAndPe_qary(TVvec, Pevec, errorUpperBoundForFrozenSet=None, numInfoIndices=None):
    """QaryPolarEncoderDecoder.frozenSetFromTVAndPe (QaryPolarEncoderDecoder.py:1156-1190): the epsilon rule, or a fixed
    number of information indices -- which, exactly as in the reference, leaves numInfoIndices + 1 of them (the slice starts
    at `numInfoIndices + 1`, :1173-1176); the single-parity fix-up loop that follows never fires there (`numpy.bool_ is
    False` is never true, :1180) and is therefore not reproduced."""
    s = np.add(TVvec, Pevec)
    N = len(s)
    order = sorted(range(N), key=lambda k: s[k])
    if numInfoIndices is None:
        err, idx = 0.0, -1
        while err < errorUpperBoundForFrozenSet and idx + 1 < N:
            i = order[idx + 1]
            if s[i] + err <= errorUpperBoundForFrozenSet:
                err += s[i]
                idx += 1
            else:
                break
    else:
        idx = numInfoIndices
    return set(int(i) for i in order[idx + 1:])


def calcFrozenSet_degradingUpgrading_qary(q, n, L, xDistribution, xyDistribution, upperBoundOnErrorProbability=None,
                                          numInfoIndices=None, threads=None):
    """QaryMemorylessDistribution.calcFrozenSet_degradingUpgrading (ScalarDistributions/QaryMemorylessDistribution.py:910-932)
    for a uniform input (xDistribution=None); no directory cache -- the native pass takes seconds."""
    assert n >= 0 and L > 0 and xyDistribution is not None
    assert upperBoundOnErrorProbability is None or upperBoundOnErrorProbability > 0
    if xDistribution is not None:
        from ._lib import PolarcubError
        raise PolarcubError("calcFrozenSet_degradingUpgrading: only the uniform-input form (xDistribution=None) is offered")
    pe = tal_vardy_pe_qary(q, n, L, xyDistribution, threads)
    return frozenSetFromTVAndPe_qary(np.zeros(len(pe)), pe, upperBoundOnErrorProbability, numInfoIndices)


def calcTVAndPe_degradingUpgrading(n, L, xDistribution, xyDistribution, directory_name=None, verbosity=False, q=None,
                                   threads=None):
    """QaryMemorylessDistribution.calcTVAndPe_degradingUpgrading (ScalarDistributions/QaryMemorylessDistribution.py:934-990)
    for a uniform input, with the reference's directory cache: `<directory_name>DegradingUpgrading_L=<L>_tv.npy` / `_pe.npy`
    are loaded when both exist and written otherwise, so caches produced by either side serve the other.  Unlike the
    reference (which returns None without a directory, :935) the vectors are also returned when directory_name is None."""
    if xDistribution is not None:
        from ._lib import PolarcubError
        raise PolarcubError("calcTVAndPe_degradingUpgrading: only the uniform-input form (xDistribution=None) is offered")
    t = np.ascontiguousarray(getattr(xyDistribution, "probs", xyDistribution), dtype=np.float64)
    q = int(q or getattr(xyDistribution, "q", t.shape[1]))
    tv_name = pe_name = None
    if directory_name is not None:
        tv_name = directory_name + "{}.npy".format("DegradingUpgrading_L=" + str(L) + "_tv")
        pe_name = directory_name + "{}.npy".format("DegradingUpgrading_L=" + str(L) + "_pe")
        if verbosity:
            print(tv_name)
            print(pe_name)
        if os.path.isfile(tv_name) and os.path.isfile(pe_name):
            return np.load(tv_name), np.load(pe_name)
    pe = tal_vardy_pe_qary(q, n, L, t, threads)
    tv = np.zeros(len(pe))
    if directory_name is not None:
        if not os.path.exists(directory_name):
            os.makedirs(directory_name)
        np.save(tv_name, tv)
        np.save(pe_name, pe)
    return tv, pe
