"""Sweep of the host-pipeline chunk size for the SCL workload + raw pinned H2D bandwidth (scratch tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from polarcub_b200 import engine
w = bench.SclBinary4096()
class A: pass
a = A()
w.allow_ga = False
w.setup(torch.device("cuda", 0), 0, 32768, 32768)
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.time()
for _ in range(4): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 4 * (1 << 30) / (time.time() - t) / 1e9)
for chunk in (1776, 2368, 3552, 7104):
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        engine.scl_decode_probs_host(w.plan, w.L, w.xy_host, w.fv_host, w.ai_host, w.info_host, w.res_host, chunk=chunk)
        torch.cuda.synchronize(); dt = time.time() - t
    print("chunk", chunk, "frames/s", round(w.Be / dt))
