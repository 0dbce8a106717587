"""Host-side logic that needs no GPU: the mirror classes' constructor conventions (frozen mask, CPython MT19937 common
randomness, frozen-bit rule), the guard-band adapters against the live-reference goldens, the bit-packing convention of the
C-ABI buffers, the construction helpers, and the loud failure of every compute entry point when there is no CUDA device."""
import os
import random

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_binary_constructor_conventions():
    """SURVEY.md 8c known answer: N = 8, frozenSet = {0,1,2,4}, seed 1 and seed -1 (BinaryPolarEncoderDecoder.py:24-44, :258-262)."""
    import polarcub_b200 as pcb
    ed = pcb.BinaryPolarEncoderDecoder(8, {0, 1, 2, 4}, 1)
    assert ed.k == 4 and ed.length == 8
    np.testing.assert_array_equal(ed.frozenMask, [1, 1, 1, 0, 1, 0, 0, 0])
    np.testing.assert_allclose(ed.randomlyGeneratedNumbers,
                               [0.134364, 0.847434, 0.763775, 0.255069, 0.495435, 0.449491, 0.651593, 0.788723], atol=5e-7)
    r = random.Random()
    r.seed(1)
    assert [r.random() for _ in range(8)] == list(ed.randomlyGeneratedNumbers)  # the stdlib stream, bit for bit
    # u_i = 0 iff 0.5 >= r_i: the survey's u = [0,1,1,*,0,*,*,*] at the frozen positions
    np.testing.assert_array_equal(ed.frozenValues[[0, 1, 2, 4]], [0, 1, 1, 0])
    ed1 = pcb.BinaryPolarEncoderDecoder(8, {0, 1, 2, 4}, -1)
    assert np.all(ed1.randomlyGeneratedNumbers == 1.0) and np.all(ed1.frozenValues == 1)  # every frozen bit is 1
    with pytest.raises(AssertionError):
        pcb.BinaryPolarEncoderDecoder(12, set(), 1)  # not a power of two


def test_qary_constructor_conventions():
    """QaryPolarEncoderDecoder.py:27-63: sorted frozen / information sets, k, frozen symbols are 0."""
    import polarcub_b200 as pcb
    ed = pcb.QaryPolarEncoderDecoder(3, 8, {4, 0, 2, 1}, 1)
    assert ed.k == 4 and ed.length == 8 and ed.q == 3
    assert list(ed.frozenSet) == [0, 1, 2, 4] and list(ed.infoSet) == [3, 5, 6, 7]
    np.testing.assert_array_equal(ed.frozenMask, [1, 1, 1, 0, 1, 0, 0, 0])
    from polarcub_b200._lib import PolarcubError
    lg = pcb.QaryPolarEncoderDecoder(2, 8, {0}, 1, use_log=True)
    assert lg.use_log
    with pytest.raises(PolarcubError, match="use_log"):  # the symbol-input list decoder is linear-domain only: loud, not silent
        lg.listDecode_symbols_batch(np.zeros((1, 8), dtype=np.uint8), np.full((2, 2), 0.5), np.zeros((1, 1)), 2,
                                    np.zeros((1, 7)))


def test_guard_bands_match_reference_goldens():
    """addDeletionGuardBands (Guardbands.py:4-40) on every encoded vector of tests/golden/trellis.npz (produced by the live
    reference) and the split of the received words used by the decoder: the trimmed sub-words concatenate back to the
    received word without its outer zero runs only where the reference trims them (removeDeletionGuardBands, :43-93)."""
    from polarcub_b200 import Guardbands
    d = np.load(os.path.join(GOLD, "trellis.npz"), allow_pickle=True)
    for nm in d["names"]:
        n, n0, k, ones, seed, frames = (int(v) for v in d[nm + "/params"])
        delta, xi = (float(v) for v in d[nm + "/chan"])
        starts, total = Guardbands.guard_band_layout(n, n0, xi, ones)
        assert len(starts) == 1 << (n - n0)
        for f in range(frames):
            enc = [int(b) for b in d[nm + "/enc"][f]]
            want = d[nm + "/cwgb"][f][: int(d[nm + "/cwgb_len"][f])]
            got = Guardbands.addDeletionGuardBands(enc, n, n0, xi, ones)
            assert len(got) == total
            np.testing.assert_array_equal(np.asarray(got, dtype=np.uint8), want)
            sub = 1 << n0
            for t, s0 in enumerate(starts):  # the data bits sit at the layout's offsets, inside their runs of ones
                assert got[s0 + ones:s0 + ones + sub] == enc[t * sub:(t + 1) * sub]
                assert got[s0:s0 + ones] == [1] * ones and got[s0 + ones + sub:s0 + 2 * ones + sub] == [1] * ones
            rx = [int(b) for b in d[nm + "/rx"][f][: int(d[nm + "/rx_len"][f])]]
            parts = Guardbands.removeDeletionGuardBands(rx, n, n0)
            assert len(parts) == 1 << (n - n0)
            bits, lens = Guardbands.split_batch([rx], n, n0)
            assert bits.shape[0] == 1 and bits.shape[1] == 1 << (n - n0)
            for t, part in enumerate(parts):
                assert int(lens[0, t]) == len(part)
                np.testing.assert_array_equal(bits[0, t, : len(part)], np.asarray(part, dtype=np.uint8))


def test_frozen_set_from_tv_and_pe_matches_reference_goldens():
    """frozenSetFromTVAndPe (BinaryPolarEncoderDecoder.py:519-548) on the (TV, Pe) statistics the live reference's genie pass
    wrote to its frozen-bits file (tests/golden/genie.npz)."""
    import contextlib
    import io
    from polarcub_b200 import simulation
    g = np.load(os.path.join(GOLD, "genie.npz"), allow_pickle=True)
    for nm in g["names"]:
        stats = np.asarray(g[nm + "/stats"], dtype=np.float64)  # the file's "(TotalVariation + errorProbability) * trials" per index
        n, trials, eb = int(g[nm + "/params"][0]), int(g[nm + "/params"][3]), float(g[nm + "/chan"][2])
        assert stats.shape == (1 << n,)
        with contextlib.redirect_stdout(io.StringIO()):
            fs = simulation.frozenSetFromTVAndPe(list(stats / trials), [0.0] * len(stats), eb)
        assert sorted(fs) == sorted(int(i) for i in g[nm + "/frozen"])


def test_pack_unpack_bits_convention():
    """Bit i of a frame sits in word i / 32 at position i % 32 (include/polarcub_b200.h)."""
    from polarcub_b200 import engine
    rng = np.random.default_rng(3)
    for nbits in (1, 31, 32, 33, 100, 4096):
        b = rng.integers(0, 2, size=(5, nbits))
        w = engine.pack_bits(b)
        assert w.dtype == np.uint32 and w.shape == (5, (nbits + 31) // 32)
        np.testing.assert_array_equal(engine.unpack_bits(w, nbits), b)
        i = nbits - 1
        assert ((w[:, i // 32] >> (i % 32)) & 1).tolist() == b[:, i].tolist()
        if nbits % 32:
            assert np.all((w[:, -1] >> (nbits % 32)) == 0)  # padding bits are zero


def test_construction_helpers():
    from polarcub_b200 import construction
    pe = construction.bec_pe(2, 0.5)  # z -> (2z - z^2, z^2), minus first: [0.9375, 0.5625, 0.4375, 0.0625] / 2
    np.testing.assert_allclose(pe, 0.5 * np.array([0.9375, 0.5625, 0.4375, 0.0625]))
    assert construction.frozen_set_from_pe(pe, 1) == {0, 1, 2}
    assert construction.frozen_set_from_pe(np.array([0.1, 0.1, 0.3, 0.0]), 2) == {1, 2}  # stable order keeps index 0 before 1
    name = "bsc_p0.11_n10_L100_pe.npy"
    pe10 = construction.load_pe(name)
    assert pe10.shape == (1024,) and np.all((pe10 >= 0) & (pe10 <= 0.5))
    assert len(construction.frozen_set_from_pe(pe10, 512)) == 512


def test_compute_calls_fail_loudly_without_cuda():
    """No CPU fallback: on a host without a CUDA device every compute entry point raises PolarcubError."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import polarcub_b200 as pcb
    from polarcub_b200._lib import PolarcubError
    ed = pcb.BinaryPolarEncoderDecoder(8, {0, 1, 2, 4}, 1)
    with pytest.raises(PolarcubError):
        ed.encode_batch(np.zeros((2, ed.k), dtype=np.int64))
    with pytest.raises(PolarcubError):
        ed.decode_batch(np.full((2, 8, 2), 0.5))
    qd = pcb.QaryPolarEncoderDecoder(3, 8, {0, 1, 2, 4}, 1)
    with pytest.raises(PolarcubError):
        qd.decode_batch(np.full((2, 8, 3), 1.0 / 3))


def test_four_state_algebra_matches_the_reference_arithmetic():
    """The erasure-type fast path (csrc/sc_binary.cu: node01, f8 / g8) rests on one claim: on probability pairs that are hard
    knowledge, an erasure or the (0,0) contradiction, BinaryMemorylessVectorDistribution's minus / plus transforms followed by
    the max-normalisation (BinaryMemorylessVectorDistribution.py:15-47, :71-87) are exact in float64 and stay within those four
    states.  Checked here by running the reference's arithmetic on every operand combination -- normalised states and raw
    BEC table rows -- next to a Python copy of the kernels' bit logic (bit 0 side, bit 1 erasure, bit 2 contradiction)."""
    def normalise(p0, p1):
        m = max(p0, p1)
        return (p0 / m, p1 / m) if m > 0 else (p0, p1)

    def minus(a, b):
        return normalise(a[0] * b[0] + a[1] * b[1], a[0] * b[1] + a[1] * b[0])

    def plus(a, b, u):
        return normalise(a[0] * b[0], a[1] * b[1]) if u == 0 else normalise(a[1] * b[0], a[0] * b[1])

    def f8(a, b):
        o = a | b
        c = o & 4
        e = (o & 2) & ~(c >> 1)
        s = ((a ^ b) & 1) & ~(e >> 1) & ~(c >> 2)
        return c | e | s

    def g8(a, b, u):
        ea, eb = (a >> 1) & 1, (b >> 1) & 1
        sa, sb = (a ^ u) & 1, b & 1
        c = ((((a | b) >> 2) & 1) | (~(ea | eb) & (sa ^ sb))) & 1
        e = ea & eb & ~c & 1
        s = ((eb & sa) | (~eb & sb)) & 1 & ~c & ~e
        return (c << 2) | (e << 1) | s

    states = {0: (1.0, 0.0), 1: (0.0, 1.0), 2: (1.0, 1.0), 4: (0.0, 0.0)}
    code_of = {v: k for k, v in states.items()}
    for ca, pa in states.items():
        for cb, pb in states.items():
            assert code_of[minus(pa, pb)] == f8(ca, cb), (ca, cb)
            for u in (0, 1):
                assert code_of[plus(pa, pb, u)] == g8(ca, cb, u), (ca, cb, u)
    # raw channel rows of makeBEC (BinaryMemorylessDistribution.py:493-499) land in the same states after one transform,
    # for any erasure probability: the products and sums involve one non-zero term at most, or two equal ones
    for p in (0.1, 0.35, 0.5, 1e-3):
        rows = {0: (0.5 * (1 - p), 0.0), 1: (0.0, 0.5 * (1 - p)), 2: (0.5 * p, 0.5 * p)}
        for ca, ra in rows.items():
            for cb, rb in rows.items():
                assert code_of[minus(ra, rb)] == f8(ca, cb)
                for u in (0, 1):
                    assert code_of[plus(ra, rb, u)] == g8(ca, cb, u)
    # leaf rule (BinaryPolarEncoderDecoder.py:250-252): p0 >= p1 -> 0, so only "hard 1" decides 1 -- the side bit
    for c, (p0, p1) in states.items():
        assert (0 if p0 >= p1 else 1) == (c & 1)


def test_calculate_syndrome_and_complement_vs_live_reference_golden(golden_dir):
    """QaryPolarEncoderDecoder.calculate_syndrome_and_complement (QaryPolarEncoderDecoder.py:822-833), the host half of ir():
    (w, u) identical to the live reference's on the goldens of oracle/gen_golden_ir.py.  No GPU needed (no plan is built)."""
    from polarcub_b200.QaryPolarEncoderDecoder import QaryPolarEncoderDecoder
    g = np.load(os.path.join(golden_dir, "ir.npz"))
    for nm in [str(s) for s in g["names"]]:
        q, n = int(g[nm + "/q"]), int(g[nm + "/n"])
        fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
        ed = QaryPolarEncoderDecoder(q, 1 << n, fs, 1)
        for f in range(g[nm + "/a"].shape[0]):
            w, u = ed.calculate_syndrome_and_complement(np.copy(g[nm + "/a"][f]))
            np.testing.assert_array_equal(w, g[nm + "/w"][f])
            np.testing.assert_array_equal(u, g[nm + "/u"][f])
            np.testing.assert_array_equal(ed.get_message_info_bits(u), g[nm + "/a_key"][f])


def test_reference_built_trellis_objects_are_rejected_with_a_clear_error():
    """The reference's CollectionOfBinaryTrellises / BinaryTrellis hold Python trellis graphs the CUDA path cannot read: the
    drop-in needs the builder swapped too, and says so (a bare shape assertion used to fire)."""
    from polarcub_b200._lib import PolarcubError
    from polarcub_b200.BinaryPolarEncoderDecoder import _probs_of

    class CollectionOfBinaryTrellises:  # duck-type of the reference's class (VectorDistributions/CollectionOfBinaryTrellises.py:24-32)
        def __init__(self):
            self.length, self.numberOfTrellises, self.trellises = 8, 2, [object(), object()]

        def __len__(self):
            return self.length

    with pytest.raises(PolarcubError, match="buildCollectionOfBinaryTrellises_uniformInput_deletion"):
        _probs_of(CollectionOfBinaryTrellises(), 8, 2)
    with pytest.raises(AssertionError):
        _probs_of(np.zeros((4, 2)), 8, 2)
