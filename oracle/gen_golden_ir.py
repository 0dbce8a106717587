"""Golden vectors for QaryPolarEncoderDecoder.ir / calculate_syndrome_and_complement from the LIVE reference
(QaryPolarEncoderDecoder.py:822-858; build container only): python oracle/gen_golden_ir.py -> tests/golden/ir.npz.

a is Alice's string, b Bob's correlated string (a through a q-ary symmetric channel); make_xyVectorDistribution(b) builds
the vector distribution from b with a seeded multiplicative jitter, so that the list metrics are tie-free (see
oracle/polar_oracle_list.c on the candidate order).  check_size = 0: np.random.choice draws an empty check matrix."""
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


def main():
    ref = refshim.load()
    out, names = {}, []
    for name, q, n, L, p, frames, seed in (("q2_n6_L4", 2, 6, 4, 0.16, 8, 11), ("q2_n8_L8", 2, 8, 8, 0.06, 4, 12),
                                           ("q3_n5_L4", 3, 5, 4, 0.16, 6, 13), ("q2_n7_L1", 2, 7, 1, 0.12, 6, 14)):
        N = 1 << n
        fs = set(int(i) for i in bec_z_order(n)[:N // 2])
        ed = ref.QPED.QaryPolarEncoderDecoder(q, N, fs, 1)
        rng = random.Random(seed)
        nrng = np.random.default_rng(seed)
        rec = {k: [] for k in ("a", "b", "jit", "w", "u", "a_key", "b_key", "pr")}
        for f in range(frames):
            a = np.array([rng.randrange(q) for _ in range(N)], dtype=np.int64)
            b = np.array([x if rng.random() > p else (x + rng.randrange(1, q)) % q for x in a], dtype=np.int64)
            jit = 1.0 + 0.2 * nrng.random((N, q))

            def make_xy(bv, jit=jit):
                vd = ref.QMVD.QaryMemorylessVectorDistribution(q, N)
                for i in range(N):
                    for x in range(q):
                        vd.probs[i][x] = (1.0 - p if x == bv[i] else p / (q - 1)) * jit[i][x]
                return vd

            w, u = ed.calculate_syndrome_and_complement(np.copy(a))
            np.random.seed(seed + f)
            a_key, b_key, pr = ed.ir(np.copy(a), np.copy(b), make_xy, list_size=L, check_size=0)
            for k, v in (("a", a), ("b", b), ("jit", jit), ("w", w), ("u", u), ("a_key", a_key), ("b_key", b_key), ("pr", pr.value)):
                rec[k].append(np.asarray(v))
        fm = np.zeros(N, dtype=np.uint8)
        fm[list(fs)] = 1
        out[name + "/q"], out[name + "/n"], out[name + "/L"], out[name + "/p"] = np.int64(q), np.int64(n), np.int64(L), np.float64(p)
        out[name + "/frozen"] = fm
        for k in rec:
            out[name + "/" + k] = np.array(rec[k])
        names.append(name)
        print(name, "ProbResult", [int(x) for x in rec["pr"]], flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ir.npz"), **out)
    print("wrote tests/golden/ir.npz")


if __name__ == "__main__":
    main()
