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
