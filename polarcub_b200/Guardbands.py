"""Guard bands for the deletion channel -- host-side mirror of the reference's Guardbands.py:4-93.

Same functions and conventions (plain lists of 0/1 in, lists out) so `main_deletion.py`'s closures run unchanged on top
of this module; the work is done on numpy arrays with index ranges instead of list slicing, and
`split_batch` produces the fixed-width arrays the CUDA trellis decoder ingests.
"""
import math

import numpy as np


def guard_band_layout(n, n0, xi, ones=0):
    """Positions of the 2^(n-n0) sub-words (each 2^n0 + 2*ones symbols) inside the guarded codeword and its total
    length: between the two halves of a length-2^m block sits a band of floor(2^((1-xi)(m-1))) zeros (Guardbands.py:23)."""
    T, sub = 1 << max(n - n0, 0), (1 << min(n, n0)) + 2 * ones
    starts = np.zeros(T, dtype=np.int64)
    pos = 0
    for t in range(T):
        if t:
            m = n0 + 1 + ((t & -t).bit_length() - 1)  # the block whose two halves meet before sub-word t
            pos += math.floor(2 ** ((1 - xi) * (m - 1)))
        starts[t] = pos
        pos += sub
    return starts, pos


def addDeletionGuardBands(encodedVector, n, n0, xi, numberOfOnesToAddAtBothEndsOfGuardbands=0):
    """Guardbands.py:4-44: zeros between halves at every level above n0, optional runs of ones around each sub-word."""
    ones = numberOfOnesToAddAtBothEndsOfGuardbands
    enc = np.asarray(encodedVector, dtype=np.int64)
    if n <= n0 and ones == 0:
        return encodedVector
    assert n <= n0 or len(enc) % 2 == 0
    starts, total = guard_band_layout(n, n0, xi, ones)
    out = np.zeros(total, dtype=np.int64)
    sub = len(enc) // len(starts)
    for t, s in enumerate(starts):
        out[s:s + ones] = 1
        out[s + ones:s + ones + sub] = enc[t * sub:(t + 1) * sub]
        out[s + ones + sub:s + 2 * ones + sub] = 1
    return [int(v) for v in out]


def _trim(a, lo, hi):
    """[lo, hi) -> the range from the first 1 to the last 1 (empty if there is none), Guardbands.py:66-93."""
    nz = np.flatnonzero(a[lo:hi] == 1)
    if nz.size == 0:
        return lo, lo
    return lo + int(nz[0]), lo + int(nz[-1]) + 1


def split_ranges(a, n, n0):
    """Index ranges of the trimmed sub-words, in order (Guardbands.py:47-63)."""
    work = [(0, len(a), n)]
    out = []
    while work:
        lo, hi, m = work.pop()
        lo, hi = _trim(a, lo, hi)
        if m <= n0:
            out.append((lo, hi))
        else:
            mid = lo + (hi - lo) // 2
            work.append((mid, hi, m - 1))
            work.append((lo, mid, m - 1))
    return out


def trimZerosAtEdges(receivedWord):
    a = np.asarray(receivedWord, dtype=np.int64)
    lo, hi = _trim(a, 0, len(a))
    return [int(v) for v in a[lo:hi]]


def removeDeletionGuardBands(receivedWord, n, n0):
    a = np.asarray(receivedWord, dtype=np.int64)
    return [[int(v) for v in a[lo:hi]] for lo, hi in split_ranges(a, n, n0)]


def split_batch(receivedWords, n, n0, maxlen=None):
    """Batch form: list of received words -> (sub_bits uint8 [B, T, maxlen], sub_len int32 [B, T])."""
    T = 1 << max(n - n0, 0)
    rngs = []
    arrs = []
    for rw in receivedWords:
        a = np.asarray(rw, dtype=np.uint8)
        arrs.append(a)
        rngs.append(split_ranges(a, n, n0))
    longest = max((hi - lo for r in rngs for lo, hi in r), default=0)
    if maxlen is None:
        maxlen = max(longest, 1)
    assert longest <= maxlen, "sub-word of %d symbols exceeds maxlen %d" % (longest, maxlen)
    bits = np.zeros((len(arrs), T, maxlen), dtype=np.uint8)
    lens = np.zeros((len(arrs), T), dtype=np.int32)
    for b, (a, r) in enumerate(zip(arrs, rngs)):
        for t, (lo, hi) in enumerate(r):
            bits[b, t, :hi - lo] = a[lo:hi]
            lens[b, t] = hi - lo
    return bits, lens
