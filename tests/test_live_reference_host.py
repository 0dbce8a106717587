"""Host-side mirrors against the LIVE reference on randomised inputs.  Runs only where the reference tree is present (the
build container); skipped elsewhere -- the committed goldens (tests/golden/*.npz) carry the parity bar on their own."""
import random

import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


def test_guard_bands_and_deletion_channel_randomised():
    """Guardbands.py:4-93 and BinaryTrellis.deletionChannelSimulation (:441-461): same lists out for the same lists in,
    including all-zero and empty received words, over random (n, n0, xi, ones, deletion probability)."""
    ref = refshim.load()
    from polarcub_b200 import Guardbands, BinaryTrellis
    rng = random.Random(99)
    for trial in range(300):
        n = rng.randint(1, 7)
        n0 = rng.randint(0, n + 1)
        xi = rng.choice([0.0, 0.1, 0.2, 0.5])
        ones = rng.choice([0, 0, 1, 2])
        enc = [rng.randint(0, 1) for _ in range(1 << n)]
        want = ref.Guardbands.addDeletionGuardBands(list(enc), n, n0, xi, ones)
        got = Guardbands.addDeletionGuardBands(list(enc), n, n0, xi, ones)
        assert list(got) == list(want), (trial, n, n0, xi, ones)
        delta = rng.choice([0.0, 0.05, 0.3, 0.9, 1.0])
        r1, r2 = random.Random(), random.Random()
        r1.seed(trial)
        r2.seed(trial)
        rx_want = ref.BT.deletionChannelSimulation(want, delta, seed=None, randomNumberGenerator=r1)
        rx_got = BinaryTrellis.deletionChannelSimulation(got, delta, seed=None, randomNumberGenerator=r2)
        assert rx_got == rx_want
        assert BinaryTrellis.deletionChannelSimulation(got, delta, 7) == ref.BT.deletionChannelSimulation(want, delta, 7)
        if trial % 7 == 0:
            rx_want = [0] * len(rx_want)  # nothing but zeros survives
        assert Guardbands.trimZerosAtEdges(rx_want) == ref.Guardbands.trimZerosAtEdges(rx_want)
        parts_want = ref.Guardbands.removeDeletionGuardBands(rx_want, n, n0)
        parts_got = Guardbands.removeDeletionGuardBands(rx_want, n, n0)
        assert [list(p) for p in parts_got] == [list(p) for p in parts_want], (trial, n, n0)
        bits, lens = Guardbands.split_batch([rx_want], n, n0)
        assert [list(bits[0, t, : lens[0, t]]) for t in range(bits.shape[1])] == [list(p) for p in parts_want]


def test_frozen_set_selection_randomised():
    """frozenSetFromTVAndPe of both classes (BinaryPolarEncoderDecoder.py:519-548, QaryPolarEncoderDecoder.py:1156-1190) on
    random (TV, Pe) vectors with ties."""
    import contextlib
    import io
    ref = refshim.load()
    from polarcub_b200 import simulation, construction
    rng = np.random.default_rng(4)
    for trial in range(60):
        N = 1 << int(rng.integers(1, 8))
        pe = np.round(rng.random(N) ** 3, 2)  # ties on purpose
        tv = np.round(rng.random(N) ** 4, 2) * (trial % 2)
        eb = float(rng.choice([0.05, 0.5, 2.0, 100.0]))
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref.BPED.frozenSetFromTVAndPe(list(tv), list(pe), eb)
            got = simulation.frozenSetFromTVAndPe(list(tv), list(pe), eb)
        assert got == want
        assert construction.frozen_set_from_tv_and_pe(tv, pe, eb) == want
        assert construction.frozenSetFromTVAndPe_qary(tv, pe, eb, None) == ref.QPED.frozenSetFromTVAndPe(tv, pe, eb, None)
        k = int(rng.integers(0, N))
        assert construction.frozenSetFromTVAndPe_qary(tv, pe, None, k) == ref.QPED.frozenSetFromTVAndPe(tv, pe, None, k)
