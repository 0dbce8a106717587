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
