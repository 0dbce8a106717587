"""Sweep of the host-pipeline chunk size for the C1 workload (SC N=1024, bit-packed symbols in pinned host memory): scratch tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from polarcub_b200 import engine  # noqa: E402

w = bench.ScBinary1024()
w.allow_ga = False
B = 1 << 20
w.setup(torch.device("cuda", 0), 0, B, B)
wave = engine.sc_wave_frames(w.plan)
print("wave", wave)
for mult in (0.5, 1, 2, 3, 4, 9.3):
    chunk = min(B, int(wave * mult) // 32 * 32)
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.time()
        engine.sc_decode_symbols_host(w.plan, w.y_host, w.tab, w.cw_host, w.info_host, packed_bits=w.host_bits, chunk=chunk)
        torch.cuda.synchronize()
        dt = time.time() - t
    print("chunk", chunk, "M frames/s %.2f" % (B / dt / 1e6), "Gbit/s %.2f" % (B * 512 / dt / 1e9))
