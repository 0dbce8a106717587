"""GPU parity: deletion-channel SC decoding over trellis collections (csrc/trellis.cu through pc_trellis_decode) against
(a) the golden vectors produced by the LIVE reference (tests/golden/trellis.npz) and (b) the CPU oracle on seeded random
batches (BASELINE config C5: N = 256, n0 = 2, guard bands, deletion probability 0.1).
Bar: BIT-EXACT decisions and float64-identical first collapsed vectors."""
import os
import random

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _bec_frozen(n, k):
    z = [0.5]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return set(int(i) for i in np.argsort(-np.array(z), kind="stable")[:(1 << n) - k])


def test_trellis_goldens(golden_dir):
    import polarcub_b200 as pcb
    from polarcub_b200 import CollectionOfBinaryTrellises as CBT
    g = np.load(os.path.join(golden_dir, "trellis.npz"))
    checked = 0
    for nm in [str(s) for s in g["names"]]:
        n, n0, k, ones, seed, frames = (int(v) for v in g[nm + "/params"])
        delta, xi = (float(v) for v in g[nm + "/chan"])
        N = 1 << n
        fs = set(int(i) for i in np.nonzero(g[nm + "/frozen"])[0])
        ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
        np.testing.assert_array_equal(ed.randomlyGeneratedNumbers, g[nm + "/r"])
        xvd = np.full((N, 2), 0.5)
        rxs = [list(int(v) for v in g[nm + "/rx"][f][:int(g[nm + "/rx_len"][f])]) for f in range(frames)]
        # single-frame drop-in call, as main_deletion.py:52-59 makes it
        coll = CBT.buildCollectionOfBinaryTrellises_uniformInput_deletion(rxs[0], delta, xi, n, n0, ones)
        cw, info = ed.decode(xvd, coll)
        np.testing.assert_array_equal(cw, g[nm + "/dec_cw"][0], err_msg=nm)
        np.testing.assert_array_equal(info, g[nm + "/dec_info"][0], err_msg=nm)
        # batched call
        collb = CBT.buildCollectionBatch_uniformInput_deletion(rxs, delta, xi, n, n0, ones)
        cwb, infob, col = ed.decode_trellis_batch(collb, want_collapse=True)
        np.testing.assert_array_equal(cwb, g[nm + "/dec_cw"], err_msg=nm)
        np.testing.assert_array_equal(infob, g[nm + "/dec_info"], err_msg=nm)
        assert np.array_equal(col, g[nm + "/collapse"]), nm
        # encode + guard bands through the host mirror
        from polarcub_b200 import Guardbands
        enc = ed.encode_batch(g[nm + "/info"])
        np.testing.assert_array_equal(enc, g[nm + "/enc"], err_msg=nm)
        for f in range(frames):
            cwgb = Guardbands.addDeletionGuardBands([int(b) for b in enc[f]], n, n0, xi, ones)
            np.testing.assert_array_equal(cwgb, g[nm + "/cwgb"][f][:int(g[nm + "/cwgb_len"][f])], err_msg=nm)
        checked += frames
    assert checked >= 70


@pytest.mark.parametrize("n,n0,k,delta,xi,ones,seed,frames", [
    (8, 2, 96, 0.1, 0.1, 0, 200, 512),     # BASELINE C5 shape (main_deletion.py defaults n=8, n0=2)
    (8, 3, 110, 0.05, 0.1, 0, 1, 300),
    (7, 2, 50, 0.1, 0.1, 1, -1, 300),
    (9, 4, 200, 0.03, 0.2, 0, 3, 64),
    (6, 1, 20, 0.15, 0.1, 2, 5, 300),
    (10, 2, 400, 0.02, 0.1, 0, 7, 40),
])
def test_trellis_vs_oracle(n, n0, k, delta, xi, ones, seed, frames):
    import polarcub_b200 as pcb
    from polarcub_b200 import CollectionOfBinaryTrellises as CBT, Guardbands
    N = 1 << n
    fs = _bec_frozen(n, k)
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
    rng = np.random.default_rng(1000 + n * 10 + n0)
    info = rng.integers(0, 2, size=(frames, k))
    enc = ed.encode_batch(info)
    chan = random.Random(100)
    rxs = []
    for f in range(frames):
        tx = Guardbands.addDeletionGuardBands([int(b) for b in enc[f]], n, n0, xi, ones)
        rxs.append([b for b in tx if not chan.random() < delta])  # BinaryTrellis.deletionChannelSimulation :441-461
    rxs[1] = []                      # everything deleted
    rxs[2] = [0] * 7                 # no ones at all
    rxs[3] = rxs[3] + [1, 0, 1, 1]   # a word longer than transmitted: some sub-words exceed the trellis length
    coll = CBT.buildCollectionBatch_uniformInput_deletion(rxs, delta, xi, n, n0, ones)
    cw, dinfo, col = ed.decode_trellis_batch(coll, want_collapse=True)
    maxlen = coll.sub_bits.shape[2]
    bad = 0
    for f in range(frames):
        sb, sl, ov = oracle.remove_guard_bands(np.array(rxs[f], dtype=np.uint8), n, n0, maxlen)
        assert not ov
        np.testing.assert_array_equal(sb, coll.sub_bits[f])
        np.testing.assert_array_equal(sl, coll.sub_len[f])
        ocw, oinfo, ocol = oracle.trellis_decode(n, n0, ed.frozenMask, ed.randomlyGeneratedNumbers, sb, sl, delta, ones,
                                                 want_collapse=True)
        if not (np.array_equal(cw[f], ocw) and np.array_equal(dinfo[f], oinfo) and np.array_equal(col[f], ocol[0])):
            bad += 1
    assert bad == 0, "%d of %d frames differ from the oracle" % (bad, frames)
