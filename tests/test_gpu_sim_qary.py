"""The q-ary Monte-Carlo drivers (irSimulation, encodeDecodeSimulation: QaryPolarEncoderDecoder.py:887-982) against results
of the LIVE reference on the same seeds (oracle/gen_golden_sim.py -> tests/golden/qsim.npz): the batched GPU drivers must
consume the random streams in the reference's order and return identical statistics and ProbResult lists."""
import contextlib
import io
import os
import random

import numpy as np
import pytest

from oracle.gen_golden_sim import ED_CASES, IR_CASES, bec_z_order, closures
import importlib

Q = importlib.import_module("polarcub_b200.QaryPolarEncoderDecoder")  # the module (the package re-exports the class under this name)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qsim.npz")


@pytest.mark.parametrize("case", IR_CASES, ids=[c[0] for c in IR_CASES])
def test_ir_simulation_vs_live_reference(case):
    name, q, n, L, p, trials, use_log = case
    g = np.load(GOLD)
    N = 1 << n
    fs = set(int(i) for i in bec_z_order(n)[:N // 2])
    sim, mk = closures(q, N, p, 4242 + n, use_log)
    random.seed(1000 + n)
    np.random.seed(2000 + n)
    fer, ser, rate, prl = Q.irSimulation(q, N, sim, mk, trials, fs, maxListSize=L, checkSize=2, use_log=use_log)
    assert fer == float(g[name + "/fer"]) and ser == float(g[name + "/ser"]) and rate == float(g[name + "/rate"])
    assert [r.value for r in prl] == [int(v) for v in g[name + "/pr"]]
    assert all(isinstance(r, Q.ProbResult) for r in prl)


@pytest.mark.parametrize("case", ED_CASES, ids=[c[0] for c in ED_CASES])
def test_encode_decode_simulation_vs_live_reference(case):
    name, q, n, p, trials = case
    g = np.load(GOLD)
    N = 1 << n
    fs = set(int(i) for i in bec_z_order(n)[:N // 2])
    sim, mk = closures(q, N, p, 777 + n, False)

    class X:
        probs = np.full((N, q), 1.0 / q)

        def __len__(self):
            return N
    random.seed(3000 + n)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Q.encodeDecodeSimulation(q, N, lambda: X(), lambda v: v, sim, mk, trials, fs)
    assert buf.getvalue().strip() == str(g[name + "/line"])


def test_ir_version_2_is_refused():
    from polarcub_b200 import PolarcubError
    with pytest.raises(PolarcubError):
        Q.irSimulation(2, 8, lambda a: a, lambda b: None, 1, {0, 1}, ir_version=2)
