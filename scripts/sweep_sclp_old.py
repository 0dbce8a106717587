"""Round 1's frame-per-warp kernel on the same inputs as scripts/sweep_sclp.py (PC_SCL_WARP=1): byte-per-symbol ABI."""
import math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polarcub_b200 import engine
from polarcub_b200.construction import frozen_set_from_pe, load_pe
n, N, K, L, B = 12, 4096, 2048, 8, 16384
fs = frozen_set_from_pe(load_pe("biawgn_ebn02.0_n12_L100_pe.npy"), K)
fm = np.zeros(N, dtype=np.uint8); fm[list(fs)] = 1
dev = torch.device("cuda", 0)
plan = engine.Plan(2, n, fm, None, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(99)
info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device=dev, generator=gen)
cw = engine.qsc_encode(plan, info)
sigma = math.sqrt(1.0 / (2.0 * 0.5 * 10.0 ** 0.2))
y = (1.0 - 2.0 * cw.to(torch.float64)) + sigma * torch.randn(cw.shape, dtype=torch.float64, device=dev, generator=gen)
l0 = -(y - 1.0) ** 2 / (2 * sigma * sigma); l1 = -(y + 1.0) ** 2 / (2 * sigma * sigma); m = torch.maximum(l0, l1)
xy = torch.stack([torch.exp(l0 - m), torch.exp(l1 - m)], dim=-1).contiguous()
fv = torch.zeros((B, N - K), dtype=torch.uint8, device=dev)
for _ in range(2):
    o = engine.scl_decode_probs(plan, L, xy, fv, info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    o = engine.scl_decode_probs(plan, L, xy, fv, info)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("SWEEP old-or-bytes-abi frames=%d ms=%.2f kframes/s=%.1f frame_errors=%d knobs=%s" % (
    B, ms, B / ms, int((o["info"] != info).any(dim=1).sum()), {k: v for k, v in os.environ.items() if k.startswith("PC_SCL")}), flush=True)
