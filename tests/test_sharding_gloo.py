"""Multi-rank host logic on CPU: world_size-2 `gloo` run of the frame sharding + the one counter all-reduce.

The decode itself is done by the oracle here (no GPU in this test); what is checked is the N>1 plumbing bench.py and
a Monte-Carlo driver rely on: contiguous disjoint shards, inputs keyed by the global frame index (rank-count
invariant), and the summed counters equal to the single-process result.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N, K, FRAMES, CHUNK, P = 64, 32, 203, 16, 0.08


def _code():
    z = [0.5]
    for _ in range(6):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    order = np.argsort(-np.array(z), kind="stable")
    fm = np.zeros(N, dtype=np.uint8)
    fm[order[:N - K]] = 1
    return fm


def _frames(begin, end):
    """Frames [begin, end) of the global synthetic stream: every CHUNK-aligned block has its own seed."""
    import oracle
    from polarcub_b200.sharding import chunk_seed
    fm = _code()
    r = oracle.common_randomness(N, 1)
    xp = np.full((N, 2), 0.5)
    infos, ys = [], []
    c0 = (begin // CHUNK) * CHUNK
    while c0 < end:
        rng = np.random.default_rng(chunk_seed(99, c0, CHUNK))
        info = rng.integers(0, 2, size=(CHUNK, K))
        cw = oracle.bin_encode_batch(N, fm, r, xp, info)
        y = cw ^ (rng.random((CHUNK, N)) < P)
        lo, hi = max(begin, c0) - c0, min(end, c0 + CHUNK) - c0
        infos.append(info[lo:hi])
        ys.append(y[lo:hi])
        c0 += CHUNK
    return fm, r, xp, np.concatenate(infos), np.concatenate(ys)


def _decode_count(begin, end):
    import oracle
    from polarcub_b200.sharding import Counters
    fm, r, xp, info, y = _frames(begin, end)
    tab = np.array([[0.5 * (1 - P), 0.5 * P], [0.5 * P, 0.5 * (1 - P)]])
    _, dinfo = oracle.bin_decode_batch(N, fm, r, xp, tab[y.astype(np.int64)])
    return Counters().add(info, dinfo)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from polarcub_b200.sharding import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(FRAMES, rank, world)
        c = _decode_count(b, e)
        local = c.v.copy()
        c.all_reduce()
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack([local, c.v]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_partition():
    from polarcub_b200.sharding import shard_range
    for total in (0, 1, 7, 64, 203, 10 ** 6):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_inputs_are_rank_count_invariant():
    _, _, _, info_all, y_all = _frames(0, FRAMES)
    _, _, _, info_b, y_b = _frames(101, 150)
    np.testing.assert_array_equal(info_all[101:150], info_b)
    np.testing.assert_array_equal(y_all[101:150], y_b)


@pytest.mark.timeout(300)
def test_two_rank_gloo_counters(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    single = _decode_count(0, FRAMES).v
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0[1], r1[1])          # both ranks hold the reduced counters
    np.testing.assert_array_equal(r0[0] + r1[0], r0[1])  # reduction = sum of the shards
    np.testing.assert_array_equal(r0[1], single)         # = the single-process run over all frames
    assert single[0] == FRAMES


def test_wilson_interval_contains_rate():
    from polarcub_b200.sharding import Counters
    c = Counters()
    c.v[:3] = [1000, 37, 410]
    lo, hi = c.wilson_interval()
    assert lo < c.fer < hi and 0.02 < lo and hi < 0.06
