"""Encode-butterfly throughput (pc_encode_bits), device-resident packed words: python scripts/bench_encode.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from polarcub_b200 import engine
for n, B in ((10, 1 << 22), (12, 1 << 20), (16, 1 << 16), (18, 1 << 14), (20, 1 << 12)):
    N = 1 << n
    rng = np.random.default_rng(n)
    fm = np.zeros(N, dtype=np.uint8); fm[rng.permutation(N)[:N // 2]] = 1
    plan = engine.Plan(2, n, fm, np.zeros(N, dtype=np.uint8))
    info = torch.randint(-2**31, 2**31 - 1, (B, plan.Kw), dtype=torch.int64, device="cuda").to(torch.int32)
    for _ in range(3): cw = engine.encode_bits(plan, info)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): cw = engine.encode_bits(plan, info)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = B * (plan.Kw + plan.Nw) * 4
    print("N=2^%d B=%d: %.3f ms, %.1f M frames/s, %.1f Gbit/s coded, %.0f GB/s of packed traffic" % (n, B, ms, B / ms / 1e3, B * N / ms / 1e6, byts / ms / 1e6))
