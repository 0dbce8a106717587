"""Golden results of the q-ary Monte-Carlo drivers from the LIVE reference (build container only):
python oracle/gen_golden_sim.py -> tests/golden/qsim.npz (+ tests/golden/qsim_results.csv)

  * QaryPolarEncoderDecoder.irSimulation (QaryPolarEncoderDecoder.py:887-930), linear and use_log=True;
  * QaryPolarEncoderDecoder.encodeDecodeSimulation (:935-982), its printed error count;
  * test3.write_header / write_result (test3.py:282-312): the CSV text.

Channels: q-ary symmetric with test3.py's simulateChannel (:35-54, global `random`), make_xyVectorDistribution with a seeded
multiplicative jitter (tie-free list metrics, see oracle/polar_oracle_list.c).  The test re-creates the same closures from
the seeds stored here."""
import contextlib
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def bec_z_order(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.argsort(-np.array(z), kind="stable")


def closures(q, N, p, jitter_seed, use_log, QMVD=None):
    """simulateChannel / make_xyVectorDistribution shared by the generator (QMVD = the reference class) and the test
    (QMVD = None: plain arrays wrapped in a .probs holder)."""
    import math
    probs = [[1.0 - p if x == y else p / (q - 1) for x in range(q)] for y in range(q)]  # makeQSC, QaryMemorylessDistribution.py:780-784

    def simulateChannel(codeword):  # test3.py:35-54 with probXGivenY(x, y) = probs[y][x] / sum_y' probs[y'][x]
        received = []
        for x in codeword:
            rand = random.random()
            probSum = 0.0
            tot = sum(probs[yy][x] for yy in range(q))
            for y in range(q):
                if probSum + probs[y][x] / tot >= rand:
                    received.append(y)
                    break
                probSum += probs[y][x] / tot
        return received

    jrng = np.random.default_rng(jitter_seed)

    def make_xy(received):
        jit = 1.0 + 0.2 * jrng.random((N, q))
        arr = np.empty((N, q))
        for i in range(N):
            for x in range(q):
                v = probs[received[i]][x] * jit[i][x]
                arr[i][x] = (math.log(v) if v != 0 else -math.inf) if use_log else v
        if QMVD is None:
            class VD:
                def __init__(self, pr):
                    self.probs = pr

                def __len__(self):
                    return self.probs.shape[0]
            return VD(arr)
        vd = QMVD.QaryMemorylessVectorDistribution(q, N, use_log=use_log)
        vd.probs[:] = arr
        return vd

    return simulateChannel, make_xy


IR_CASES = (("ir_q2_n6_L4", 2, 6, 4, 0.11, 24, False), ("ir_q3_n5_L4", 3, 5, 4, 0.14, 16, False),
            ("ir_q2_n7_L8", 2, 7, 8, 0.09, 10, False), ("ir_q2_n6_L2_log", 2, 6, 2, 0.11, 16, True),
            ("ir_q3_n4_L4_log", 3, 4, 4, 0.14, 12, True))
ED_CASES = (("ed_q3_n6", 3, 6, 0.10, 20), ("ed_q2_n7", 2, 7, 0.08, 16), ("ed_q5_n4", 5, 4, 0.12, 12))


def main():
    ref = refshim.load()
    out, names = {}, []
    for name, q, n, L, p, trials, use_log in IR_CASES:
        N = 1 << n
        fs = set(int(i) for i in bec_z_order(n)[:N // 2])
        sim, mk = closures(q, N, p, 4242 + n, use_log, ref.QMVD)
        random.seed(1000 + n)
        np.random.seed(2000 + n)
        fer, ser, rate, prl = ref.QPED.irSimulation(q, N, sim, mk, trials, fs, maxListSize=L, checkSize=2, use_log=use_log)
        out[name + "/fer"], out[name + "/ser"], out[name + "/rate"] = np.float64(fer), np.float64(ser), np.float64(rate)
        out[name + "/pr"] = np.array([r.value for r in prl], dtype=np.int64)
        names.append(name)
        print(name, fer, ser, rate, [r.value for r in prl], flush=True)
    for name, q, n, p, trials in ED_CASES:
        N = 1 << n
        fs = set(int(i) for i in bec_z_order(n)[:N // 2])
        sim, mk = closures(q, N, p, 777 + n, False, ref.QMVD)

        def make_x():
            vd = ref.QMVD.QaryMemorylessVectorDistribution(q, N)
            vd.probs[:] = 1.0 / q
            return vd
        random.seed(3000 + n)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.QPED.encodeDecodeSimulation(q, N, make_x, lambda v: v, sim, mk, trials, fs)
        line = buf.getvalue().strip()
        out[name + "/line"] = np.array(line)
        names.append(name)
        print(name, line, flush=True)
    # CSV
    sys.path.insert(0, refshim.REFERENCE_ROOT)
    import test3
    path = os.path.join(ROOT, "tests", "golden", "qsim_results.csv")
    if os.path.exists(path):
        os.remove(path)
    test3.write_header(path)
    PR = ref.QPED.ProbResult
    test3.write_result(path, 2, 0.05, None, 0.7136, 6, 64, 100, "TalVardy", 32, 0.46875, 4, 0.125, 0.01, 0.4, 123.5, 24,
                       [PR.SuccessActualIsMax] * 20 + [PR.FailActualWithinRange] * 3 + [PR.SuccessActualSmallerThanMax])
    test3.write_result(path, 3, 0.02, None, 1.4, 5, 32, 100, "TalVardy", 16, 0.73, 9, 0.0, 0.0, 0.7, 1.25, 10,
                       [PR.SuccessActualIsMax] * 10)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "qsim.npz"), **out)
    print("wrote tests/golden/qsim.npz and qsim_results.csv")


if __name__ == "__main__":
    main()
