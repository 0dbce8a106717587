"""Golden outputs of the reference's Monte-Carlo drivers, produced by the LIVE reference (TEST INFRASTRUCTURE, build
container only).  Output: tests/golden/genie.npz.

  * genieSingleDecodeSimulatioan over a BSC (memoryless) for several genie seeds: channel outputs, decoded vector, Pe, H;
  * genieEncodeDecodeSimulation over the deletion channel (main_deletion.py closures :17-59, trustXYProbs False as
    main_deletion.py:128 picks for n > n0) and over a BSC (trustXYProbs True): the frozen set and the per-index
    (TV + Pe) * trials written to the frozen-bits file;
  * encodeDecodeSimulation with that frozen set: the number of misdecoded words the reference prints.
"""
import contextlib
import io
import os
import random
import re
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "genie.npz")


def parse_file(path):
    fs, stats = [], []
    for line in open(path):
        if line.startswith("*** "):
            _, i, v = line.split()
            stats.append(float(v))
        elif not line.startswith("*"):
            fs.append(int(line))
    return np.array(sorted(fs), dtype=np.int64), np.array(stats, dtype=np.float64)


def main():
    ref = refshim.load()
    BPED, BMD, BT, CBT, GB = ref.BPED, ref.BMD, ref.BT, ref.CBT, ref.Guardbands
    out = {}

    def uniform_x(N):
        def mk():
            d = BMD.BinaryMemorylessDistribution()
            d.probs.append([0.5, 0.5])
            return d.makeBinaryMemorylessVectorDistribution(N, None)
        return mk

    # ---- single genie decodes over a BSC -------------------------------------------------------------------------
    N, p = 32, 0.11
    bsc = BMD.makeBSC(p)
    ed = BPED.BinaryPolarEncoderDecoder(N, set(), 0)
    xvd = uniform_x(N)()
    rng = random.Random(77)
    ys, seeds, decs, pes, hs = [], [], [], [], []
    for t in range(8):
        seed = rng.randint(1, 1000000)
        enc, _, _ = ed.genieSingleEncodeSimulatioan(xvd, seed)
        y = [int(b) ^ (1 if rng.random() < p else 0) for b in enc]
        xy = bsc.makeBinaryMemorylessVectorDistribution(N, y)
        dec, pe, h = ed.genieSingleDecodeSimulatioan(xvd, xy, seed, True)
        ys.append(y), seeds.append(seed), decs.append(np.asarray(dec)), pes.append(pe), hs.append(h)
    out["single/N"] = np.int64(N)
    out["single/table"] = np.array(bsc.probs, dtype=np.float64)
    out["single/y"] = np.array(ys, dtype=np.uint8)
    out["single/seeds"] = np.array(seeds, dtype=np.int64)
    out["single/dec"] = np.array(decs, dtype=np.int64)
    out["single/pe"] = np.array(pes, dtype=np.float64)
    out["single/h"] = np.array(hs, dtype=np.float64)

    # ---- full drivers ---------------------------------------------------------------------------------------------------
    cases = [  # name, kind, n, n0, delta/p, xi, ones, genie trials, sim trials, error bound
        ("del_n5_n02", "deletion", 5, 2, 0.1, 0.1, 0, 40, 30, 0.3),
        ("del_n6_n02_ones1", "deletion", 6, 2, 0.05, 0.1, 1, 30, 20, 0.3),
        ("del_n7_n03", "deletion", 7, 3, 0.03, 0.1, 0, 24, 16, 0.2),
        ("bsc_n6", "bsc", 6, 0, 0.05, 0.0, 0, 40, 40, 0.1),
    ]
    out["names"] = np.array([c[0] for c in cases])
    for (nm, kind, n, n0, prm, xi, ones, gt, st, eb) in cases:
        N = 1 << n
        if kind == "deletion":
            mk_cw = lambda enc, n=n, n0=n0, xi=xi, ones=ones: GB.addDeletionGuardBands(enc, n, n0, xi, ones)
            chan_rng = random.Random()
            chan_rng.seed(100)
            sim = lambda cw, prm=prm, r=chan_rng: BT.deletionChannelSimulation(cw, prm, seed=None, randomNumberGenerator=r)
            mk_xy = lambda rw, prm=prm, xi=xi, n=n, n0=n0, ones=ones: CBT.buildCollectionOfBinaryTrellises_uniformInput_deletion(rw, prm, xi, n, n0, ones)
            trust = False if n > n0 else True
        else:
            ch = BMD.makeBSC(prm)
            mk_cw = lambda enc: enc
            chan_rng = random.Random()
            chan_rng.seed(100)
            sim = lambda cw, prm=prm, r=chan_rng: [int(b) ^ (1 if r.random() < prm else 0) for b in cw]
            mk_xy = lambda rw, ch=ch: ch.makeBinaryMemorylessVectorDistribution(len(rw), rw)
            trust = True
        tmp = tempfile.NamedTemporaryFile(suffix=".txt", delete=False).name
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            fs = BPED.genieEncodeDecodeSimulation(N, uniform_x(N), mk_cw, sim, mk_xy, gt, eb, 300, trustXYProbs=trust, filename=tmp)
        fsf, stats = parse_file(tmp)
        assert set(fsf.tolist()) == set(fs)
        buf2 = io.StringIO()
        with contextlib.redirect_stdout(buf2):
            BPED.encodeDecodeSimulation(N, uniform_x(N), mk_cw, sim, mk_xy, st, fs, commonRandomnessSeed=200, randomInformationSeed=400)
        m = re.search(r"Error probability =\s+(\d+) /", buf2.getvalue())
        out[nm + "/params"] = np.array([n, n0, ones, gt, st, 1 if trust else 0], dtype=np.int64)
        out[nm + "/chan"] = np.array([prm, xi, eb], dtype=np.float64)
        out[nm + "/frozen"] = fsf
        out[nm + "/stats"] = stats
        out[nm + "/errors"] = np.int64(int(m.group(1)))
        print(nm, "frozen", len(fsf), "of", N, "errors", int(m.group(1)), "/", st)
        os.unlink(tmp)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
