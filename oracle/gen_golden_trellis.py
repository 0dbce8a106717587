"""Golden vectors for the deletion-channel (trellis) path, produced by the LIVE reference.

TEST INFRASTRUCTURE, build container only.  Output: tests/golden/trellis.npz.  For each case: code (n, n0, frozen set,
seed), channel parameters, and per frame the information, the encoded vector, the codeword with guard bands, the received
word after the deletion channel (main_deletion.py closures :17-59), the decoded (codeword, information) pair returned by
BinaryPolarEncoderDecoder.decode over buildCollectionOfBinaryTrellises_uniformInput_deletion, and the first collapsed
(unnormalised) memoryless vector of the all-minus descent.
"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trellis.npz")


def bec_frozen(n, k):
    z = [0.5]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    order = np.argsort(-np.array(z), kind="stable")
    return set(int(i) for i in order[:(1 << n) - k])


def main():
    ref = refshim.load()
    BPED, BMD, BT, CBT, GB = ref.BPED, ref.BMD, ref.BT, ref.CBT, ref.Guardbands
    cases = [  # name, n, n0, k, delta, xi, ones, seed, frames
        ("n5_n02_d0.1", 5, 2, 12, 0.1, 0.1, 0, 200, 12),
        ("n6_n02_d0.1", 6, 2, 24, 0.1, 0.1, 0, 200, 10),
        ("n6_n03_d0.05", 6, 3, 20, 0.05, 0.1, 0, 7, 8),
        ("n8_n02_d0.1", 8, 2, 96, 0.1, 0.1, 0, 200, 6),
        ("n7_n03_d0.1", 7, 3, 40, 0.1, 0.2, 0, -1, 6),
        ("n6_n02_ones1", 6, 2, 24, 0.1, 0.1, 1, 200, 8),
        ("n6_n03_ones2", 6, 3, 20, 0.08, 0.1, 2, 5, 6),
        ("n4_n04_single", 4, 4, 6, 0.1, 0.1, 0, 200, 8),
        ("n5_n01_d0.2", 5, 1, 10, 0.2, 0.1, 0, 3, 8),
    ]
    out = {"names": np.array([c[0] for c in cases])}
    for (nm, n, n0, k, delta, xi, ones, seed, frames) in cases:
        N = 1 << n
        fs = bec_frozen(n, k)
        ed = BPED.BinaryPolarEncoderDecoder(N, fs, seed)
        xd = BMD.BinaryMemorylessDistribution()
        xd.probs.append([0.5, 0.5])
        xvd = xd.makeBinaryMemorylessVectorDistribution(N, None)
        rng_info = random.Random(400)
        rng_ch = random.Random(100)
        maxrx = 0
        rec = {"info": [], "enc": [], "cwgb": [], "rx": [], "dec_cw": [], "dec_info": [], "collapse": []}
        for _ in range(frames):
            info = [1 if rng_info.random() < 0.5 else 0 for _ in range(ed.k)]
            enc = ed.encode(xvd, info)
            cwgb = GB.addDeletionGuardBands(list(int(b) for b in enc), n, n0, xi, ones)
            rx = BT.deletionChannelSimulation(cwgb, delta, seed=None, randomNumberGenerator=rng_ch)
            coll = CBT.buildCollectionOfBinaryTrellises_uniformInput_deletion(rx, delta, xi, n, n0, ones)
            dcw, dinfo = ed.decode(xvd, coll)
            # first collapsed vector of the all-minus descent
            cur = CBT.buildCollectionOfBinaryTrellises_uniformInput_deletion(rx, delta, xi, n, n0, ones)
            for _t in range(n0):
                cur = cur.minusTransform()
                if _t < n0 - 1:
                    cur.normalize(cur.calcNormalizationVector())
            rec["info"].append(info)
            rec["enc"].append(np.asarray(enc))
            rec["cwgb"].append(cwgb)
            rec["rx"].append(rx)
            rec["dec_cw"].append(np.asarray(dcw))
            rec["dec_info"].append(np.asarray(dinfo))
            rec["collapse"].append(np.array(cur.probs, dtype=np.float64))
            maxrx = max(maxrx, len(rx), len(cwgb))
        fm = np.zeros(N, dtype=np.uint8)
        fm[list(fs)] = 1
        out[nm + "/params"] = np.array([n, n0, k, ones, seed, frames], dtype=np.int64)
        out[nm + "/chan"] = np.array([delta, xi], dtype=np.float64)
        out[nm + "/frozen"] = fm
        out[nm + "/r"] = np.array(ed.randomlyGeneratedNumbers, dtype=np.float64)
        out[nm + "/info"] = np.array(rec["info"], dtype=np.int64)
        out[nm + "/enc"] = np.array(rec["enc"], dtype=np.int64)
        out[nm + "/dec_cw"] = np.array(rec["dec_cw"], dtype=np.int64)
        out[nm + "/dec_info"] = np.array(rec["dec_info"], dtype=np.int64)
        out[nm + "/collapse"] = np.array(rec["collapse"], dtype=np.float64)
        cw_pad = np.full((frames, maxrx), 255, dtype=np.uint8)
        rx_pad = np.full((frames, maxrx), 255, dtype=np.uint8)
        for f in range(frames):
            cw_pad[f, :len(rec["cwgb"][f])] = rec["cwgb"][f]
            rx_pad[f, :len(rec["rx"][f])] = rec["rx"][f]
        out[nm + "/cwgb"] = cw_pad
        out[nm + "/rx"] = rx_pad
        out[nm + "/cwgb_len"] = np.array([len(x) for x in rec["cwgb"]], dtype=np.int64)
        out[nm + "/rx_len"] = np.array([len(x) for x in rec["rx"]], dtype=np.int64)
        errs = sum(int(not np.array_equal(a, b)) for a, b in zip(rec["info"], rec["dec_info"]))
        print(nm, "frames", frames, "frame errors", errs, "cw len", len(rec["cwgb"][0]))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
