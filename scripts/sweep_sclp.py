"""Device-resident throughput of the binary SC-list decoder on the C2 code (N=4096, K=2048, L=8, BI-AWGN 2 dB) for one
setting of the tuning knobs (they are read once per process): python scripts/sweep_sclp.py [--frames F] [--mode probs|sym]"""
import argparse
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polarcub_b200 import engine  # noqa: E402
from polarcub_b200.construction import frozen_set_from_pe, load_pe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=0, help="0 = three resident waves")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--mode", default="probs")
ap.add_argument("--L", type=int, default=8)
ap.add_argument("--ebn0", type=float, default=2.0)
a = ap.parse_args()
n, N, K, L = 12, 4096, 2048, a.L
fs = frozen_set_from_pe(load_pe("biawgn_ebn02.0_n12_L100_pe.npy"), K)
fm = np.zeros(N, dtype=np.uint8)
fm[list(fs)] = 1
dev = torch.device("cuda", 0)
plan = engine.Plan(2, n, fm, None, device=dev)
gen = torch.Generator(device=dev)
gen.manual_seed(99)
B = a.frames or 3 * engine.scl_wave_frames(plan, L)
info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device=dev, generator=gen)
cw = engine.qsc_encode(plan, info)
sigma = math.sqrt(1.0 / (2.0 * 0.5 * 10.0 ** (a.ebn0 / 10.0)))
y = (1.0 - 2.0 * cw.to(torch.float64)) + sigma * torch.randn(cw.shape, dtype=torch.float64, device=dev, generator=gen)
ai = torch.from_numpy(engine.pack_bits(info.cpu().numpy()).view(np.int32)).to(dev)
kw = {}
if a.mode == "probs":
    l0 = -(y - 1.0) ** 2 / (2 * sigma * sigma)
    l1 = -(y + 1.0) ** 2 / (2 * sigma * sigma)
    m = torch.maximum(l0, l1)
    kw["xy"] = torch.stack([torch.exp(l0 - m), torch.exp(l1 - m)], dim=-1).contiguous()
else:
    Y, ymax = 256, 1.0 + 4.0 * sigma
    step = 2 * ymax / Y
    kw["y"] = torch.clamp(torch.floor((y + ymax) / step), 0, Y - 1).to(torch.uint8).contiguous()
    edges = -ymax + step * np.arange(Y + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    from scipy.stats import norm
    kw["table"] = np.stack([0.5 * (norm.cdf((edges[1:] - 1) / sigma) - norm.cdf((edges[:-1] - 1) / sigma)),
                            0.5 * (norm.cdf((edges[1:] + 1) / sigma) - norm.cdf((edges[:-1] + 1) / sigma))], axis=-1)
del y, cw
for _ in range(2):
    o = engine.scl_decode_packed(plan, L, ai, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    o = engine.scl_decode_packed(plan, L, ai, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
ferr = int((o["info_packed"] != ai).any(dim=1).sum())
res = torch.bincount(o["prob_result"], minlength=6).cpu().numpy().tolist()
knobs = {k: v for k, v in os.environ.items() if k.startswith("PC_SCL")}
print("SWEEP mode=%s L=%d frames=%d ms=%.2f kframes/s=%.1f Gbit/s=%.3f frame_errors=%d prob_result=%s knobs=%s" % (
    a.mode, L, B, ms, B / ms, B * K / ms / 1e6, ferr, res, knobs), flush=True)
