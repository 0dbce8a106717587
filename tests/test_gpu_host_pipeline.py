"""GPU: the host-batch pipeline (engine.*_host: chunks alternate between two CUDA streams, copies overlap decoding) returns
exactly what the single-call device path returns, for chunk sizes that do and do not divide the batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bec_frozen(n, k):
    z = [0.5]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return set(int(i) for i in np.argsort(-np.array(z), kind="stable")[:(1 << n) - k])


def test_sc_symbols_host_pipeline():
    import torch
    import polarcub_b200 as pcb
    from polarcub_b200 import engine
    n, N, B = 9, 512, 5000
    ed = pcb.BinaryPolarEncoderDecoder(N, _bec_frozen(n, N // 2), 1)
    plan = ed.plan
    assert engine.sc_wave_frames(plan) > 0
    rng = np.random.default_rng(5)
    p = 0.08
    tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])
    y = torch.from_numpy(rng.integers(0, 2, size=(B, N)).astype(np.uint8))
    cw_ref, info_ref = engine.sc_decode_symbols(plan, y.cuda(), tab)
    torch.cuda.synchronize()
    for chunk in (None, 1024, 777):
        yh = y.clone().pin_memory()
        cwh = torch.zeros((B, plan.Nw), dtype=torch.int32).pin_memory()
        ih = torch.zeros((B, plan.Kw), dtype=torch.int32).pin_memory()
        engine.sc_decode_symbols_host(plan, yh, tab, cwh, ih, chunk=chunk)
        torch.cuda.synchronize()
        assert torch.equal(cwh, cw_ref.cpu()) and torch.equal(ih, info_ref.cpu()), chunk


def test_scl_host_pipeline():
    import torch
    import polarcub_b200 as pcb
    from polarcub_b200 import engine
    n, N, L, B = 8, 256, 8, 700
    k = N // 2
    ed = pcb.QaryPolarEncoderDecoder(2, N, _bec_frozen(n, k), 1)
    plan = ed.plan
    assert engine.scl_wave_frames(plan, L) > 0
    rng = np.random.default_rng(6)
    info = rng.integers(0, 2, size=(B, k)).astype(np.uint8)
    xy = rng.random((B, N, 2)) + 0.05
    fv = np.zeros((B, N - k), dtype=np.uint8)
    dev = plan.device
    ref = engine.scl_decode_probs(plan, L, torch.from_numpy(xy).to(dev), torch.from_numpy(fv).to(dev), torch.from_numpy(info).to(dev))
    torch.cuda.synchronize()
    for chunk in (None, 256, 333):
        xyh = torch.from_numpy(xy).pin_memory()
        fvh, aih = torch.from_numpy(fv).pin_memory(), torch.from_numpy(info).pin_memory()
        ih = torch.zeros((B, k), dtype=torch.uint8).pin_memory()
        rh = torch.zeros((B,), dtype=torch.int32).pin_memory()
        engine.scl_decode_probs_host(plan, L, xyh, fvh, aih, ih, rh, chunk=chunk)
        torch.cuda.synchronize()
        assert torch.equal(ih, ref["info"].cpu()) and torch.equal(rh, ref["prob_result"].cpu()), chunk
