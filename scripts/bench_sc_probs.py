"""pc_sc_decode_probs (float64 pair input) throughput at N=1024, with the time split by kernel (CUDA events via profile hook)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from polarcub_b200 import engine
w = bench.ScBinary1024(); w.code()
B = 1 << 18
plan = engine.Plan(2, w.n, w.fm, w.fv if hasattr(w, "fv") else None)
rng = np.random.default_rng(0)
y = torch.from_numpy(rng.integers(0, 2, size=(B, w.N)).astype(np.int64)).cuda()
tab = torch.from_numpy(w.tab).cuda()
xy = tab[y].contiguous()
for _ in range(2): engine.sc_decode_probs(plan, xy)
torch.cuda.synchronize()
engine.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): engine.sc_decode_probs(plan, xy)
e1.record(); torch.cuda.synchronize()
kms, kn = engine.profile_read()
ms = e0.elapsed_time(e1) / 5
print("probs path: %.2f ms per %d frames = %.1f M frames/s; decode kernel %.2f ms (%.0f %%)" % (ms, B, B / ms / 1e3, kms / 5, 100 * kms / 5 / ms))
