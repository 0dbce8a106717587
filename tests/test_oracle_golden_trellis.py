"""Pins the trellis oracle (oracle/polar_oracle_trellis.c) against golden vectors produced by the LIVE reference
(oracle/gen_golden_trellis.py): guard-band insertion / removal, and BinaryPolarEncoderDecoder.decode over a
CollectionOfBinaryTrellises -- decisions and the first collapsed float64 vector, BIT-EXACT."""
import os

import numpy as np

import oracle


def test_trellis_decode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "trellis.npz"))
    names = [str(s) for s in g["names"]]
    assert len(names) >= 9
    frames_checked = 0
    for nm in names:
        n, n0, k, ones, seed, frames = (int(v) for v in g[nm + "/params"])
        delta, xi = (float(v) for v in g[nm + "/chan"])
        N = 1 << n
        fm, r = g[nm + "/frozen"], g[nm + "/r"]
        np.testing.assert_array_equal(r, oracle.common_randomness(N, seed), err_msg=nm)
        maxlen = (1 << n0) + 2 * ones + 8
        for f in range(frames):
            enc = g[nm + "/enc"][f]
            np.testing.assert_array_equal(oracle.bin_encode(N, fm, r, np.full((N, 2), 0.5), g[nm + "/info"][f]), enc, err_msg=nm)
            cwgb = g[nm + "/cwgb"][f][:int(g[nm + "/cwgb_len"][f])]
            np.testing.assert_array_equal(oracle.add_guard_bands(enc, n, n0, xi, ones), cwgb, err_msg=nm)
            rx = g[nm + "/rx"][f][:int(g[nm + "/rx_len"][f])]
            sub_bits, sub_len, overflow = oracle.remove_guard_bands(rx, n, n0, maxlen)
            assert not overflow, (nm, f)
            cw, info, col = oracle.trellis_decode(n, n0, fm, r, sub_bits, sub_len, delta, ones, want_collapse=True)
            np.testing.assert_array_equal(cw, g[nm + "/dec_cw"][f], err_msg="%s frame %d" % (nm, f))
            np.testing.assert_array_equal(info, g[nm + "/dec_info"][f], err_msg="%s frame %d" % (nm, f))
            assert np.array_equal(col[0], g[nm + "/collapse"][f]), (nm, f)
            frames_checked += 1
    assert frames_checked >= 70
