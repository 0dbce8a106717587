"""Leaf blocks of the binary SC decoder (sc_binary.cu: NODE_BLOCK / LeafBlock, pc_plan::sched_b) against the oracle and against the
leaf-by-leaf schedule (PC_SC_BLOCK=0): every block length 2 .. 4096, random frozen sets of several rates (blocks with every mix of
frozen and information leaves, all-frozen halves, non-zero frozen values through the common randomness), symbol and probability inputs."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2])
def test_leaf_blocks_vs_oracle_and_leaf_schedule(seed, monkeypatch):
    import polarcub_b200 as pcb
    rng = np.random.default_rng(9000 + seed)
    p = 0.09
    tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])
    for n in range(1, 13):
        N = 1 << n
        for rate in (0.1, 0.5, 0.9):
            k = min(N - 1, max(1, int(rate * N)))
            fs = set(int(i) for i in rng.permutation(N)[:N - k])
            ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
            B = 70 if n <= 10 else 40
            info = rng.integers(0, 2, size=(B, ed.k))
            cw = ed.encode_batch(info)
            y = (cw ^ (rng.random(cw.shape) < p)).astype(np.uint8)
            c1, i1 = ed.decode_symbols_batch(y, tab)
            c2, i2 = ed.decode_batch(tab[y])
            monkeypatch.setenv("PC_SC_BLOCK", "0")
            c0, i0 = ed.decode_symbols_batch(y, tab)
            monkeypatch.delenv("PC_SC_BLOCK")
            msg = "n=%d rate=%s" % (n, rate)
            np.testing.assert_array_equal(i1, i0, err_msg=msg)
            np.testing.assert_array_equal(c1, c0, err_msg=msg)
            np.testing.assert_array_equal(i2, i0, err_msg=msg)
            np.testing.assert_array_equal(c2, c0, err_msg=msg)
            if n <= 10:
                ocw, oinfo = oracle.bin_decode_batch(N, ed.frozenMask, ed.randomlyGeneratedNumbers, np.full((N, 2), 0.5), tab[y])
                np.testing.assert_array_equal(i1, oinfo, err_msg=msg)
                np.testing.assert_array_equal(c1, ocw, err_msg=msg)


def test_leaf_blocks_erasures_and_contradictions():
    """BEC-type inputs: erasures (ties decode to 0) and, with a wrong frozen value, the (0,0) contradiction state inside blocks."""
    import polarcub_b200 as pcb
    rng = np.random.default_rng(77)
    e = 0.3
    tab = np.array([[0.5 * (1 - e), 0.0], [0.0, 0.5 * (1 - e)], [0.5 * e, 0.5 * e]])
    for n in (5, 8, 10):
        N = 1 << n
        fs = set(int(i) for i in rng.permutation(N)[:N // 2])
        ed = pcb.BinaryPolarEncoderDecoder(N, fs, 3)
        info = rng.integers(0, 2, size=(64, ed.k))
        cw = ed.encode_batch(info)
        y = np.where(rng.random(cw.shape) < e, 2, cw).astype(np.uint8)
        y[:8] = np.where(rng.random(cw[:8].shape) < 0.05, 1 - cw[:8], y[:8])  # hard errors: contradictions with the frozen bits
        c1, i1 = ed.decode_symbols_batch(y, tab)
        ocw, oinfo = oracle.bin_decode_batch(N, ed.frozenMask, ed.randomlyGeneratedNumbers, np.full((N, 2), 0.5), tab[y])
        np.testing.assert_array_equal(i1, oinfo, err_msg="n=%d" % n)
        np.testing.assert_array_equal(c1, ocw, err_msg="n=%d" % n)
