"""GPU parity: genie runs (marginalizedUProbs capture, BinaryPolarEncoderDecoder.py:101-221) and the batched Monte-Carlo
drivers (genieEncodeDecodeSimulation :390-491, encodeDecodeSimulation :328-387) against outputs of the LIVE reference
(tests/golden/genie.npz, oracle/gen_golden_genie.py).  Bar: decoded vectors, Pe, H, the frozen set, the per-index
(TV + Pe) * trials written to the frozen-bits file and the number of misdecoded words all IDENTICAL (float64-equal)."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_genie_single_decodes(golden_dir):
    import polarcub_b200 as pcb
    g = np.load(os.path.join(golden_dir, "genie.npz"))
    N = int(g["single/N"])
    ed = pcb.BinaryPolarEncoderDecoder(N, set(), 0)
    xvd = np.full((N, 2), 0.5)
    xy = g["single/table"][g["single/y"]]
    seeds = [int(s) for s in g["single/seeds"]]
    dec, pe, h = ed.genie_decode_batch(xvd, xy, seeds, True)
    np.testing.assert_array_equal(dec, g["single/dec"])
    assert np.array_equal(pe, g["single/pe"])
    assert np.array_equal(h, g["single/h"])
    # the reference's single-run call
    d1, p1, h1 = ed.genieSingleDecodeSimulatioan(xvd, xy[2], seeds[2], True)
    np.testing.assert_array_equal(d1, g["single/dec"][2])
    assert p1 == list(g["single/pe"][2]) and h1 == list(g["single/h"][2])
    e1, tv, he = ed.genieSingleEncodeSimulatioan(xvd, seeds[2])
    np.testing.assert_array_equal(e1, g["single/dec"][2])  # the genie decoder reproduces the encoded vector
    assert tv == [0.0] * N and he == [1.0] * N


def test_monte_carlo_drivers(golden_dir, tmp_path, capsys):
    import polarcub_b200 as pcb
    from polarcub_b200 import Guardbands, BinaryTrellis
    from polarcub_b200.CollectionOfBinaryTrellises import buildCollectionOfBinaryTrellises_uniformInput_deletion as build
    g = np.load(os.path.join(golden_dir, "genie.npz"))
    for nm in [str(s) for s in g["names"]]:
        n, n0, ones, gt, st, trust = (int(v) for v in g[nm + "/params"])
        prm, xi, eb = (float(v) for v in g[nm + "/chan"])
        N = 1 << n
        mk_x = lambda N=N: np.full((N, 2), 0.5)
        chan = random.Random()
        chan.seed(100)
        if nm.startswith("del"):
            mk_cw = lambda enc: Guardbands.addDeletionGuardBands([int(b) for b in enc], n, n0, xi, ones)
            sim = lambda cw: BinaryTrellis.deletionChannelSimulation(cw, prm, seed=None, randomNumberGenerator=chan)
            mk_xy = lambda rw: build(rw, prm, xi, n, n0, ones)
        else:
            tab = np.array([[0.5 * (1.0 - prm), 0.5 * prm], [0.5 * prm, 0.5 * (1.0 - prm)]])
            mk_cw = lambda enc: enc
            sim = lambda cw: [int(b) ^ (1 if chan.random() < prm else 0) for b in cw]
            mk_xy = lambda rw: tab[np.asarray(rw)]
        fn = str(tmp_path / (nm + ".txt"))
        fs, stats = pcb.genieEncodeDecodeSimulation(N, mk_x, mk_cw, sim, mk_xy, gt, eb, 300, trustXYProbs=bool(trust),
                                                    filename=fn, return_stats=True)
        assert sorted(fs) == g[nm + "/frozen"].tolist(), nm
        got = np.array([(stats["TV"][i] + stats["Pe"][i]) * gt for i in range(N)])
        assert np.array_equal(got, g[nm + "/stats"]), nm
        assert pcb.readFrozenSetFromFile(fn) == set(fs)
        errors = pcb.encodeDecodeSimulation(N, mk_x, mk_cw, sim, mk_xy, st, fs, commonRandomnessSeed=200, randomInformationSeed=400)
        assert errors == int(g[nm + "/errors"]), nm
