#!/usr/bin/env python
"""bench.py -- decoded info Gbit/s of the batched polar decoders on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sc1024|scl4096|qsc2048] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the decode hot path over one batch of synthetic channel outputs (frames are
independent, so every rank decodes its own slice: weak scaling, no data-path collective; one NCCL
all-reduce of the int64[3] error counters after the timed region).  Rank 0 prints ONE JSON line.

  value     whole-job decoded information Gbit/s, inputs resident in HBM, CUDA events, max over ranks
  e2e       same metric through the public batched API with HOST (pinned) buffers: H2D of the channel outputs
            and D2H of the decoded words inside the timed region
  roofline  dominant kernel (the SC/SCL decode kernel): algorithmic bytes per launch / its CUDA-event duration
  cpu_baseline  the oracle port (oracle/polar_oracle.c, the reference's algorithm restated in C) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_BSC = 0.11


def bsc_table(p=P_BSC):
    return np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])  # makeBSC, BinaryMemorylessDistribution.py:485-490


def c1_code():
    """C1: N=1024, K=512, frozen set = 512 worst indices by Tal-Vardy Pe (reference degrade pass, L=100)."""
    from polarcub_b200.construction import frozen_set_from_pe, load_pe
    pe = load_pe("bsc_p0.11_n10_L100_pe.npy")
    return 1024, 512, frozen_set_from_pe(pe, 512)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, gpu_indices):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                if int(c[0]) not in gpu_indices:
                    continue
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------
def cpu_decode_sample(N, fm, r, xy_tab_idx, table, threads):
    """Time the oracle port on `threads` host threads over a sample of frames (symbols -> table lookup -> decode)."""
    import oracle
    B = xy_tab_idx.shape[0]
    xp = np.full((N, 2), 0.5)
    parts = np.array_split(np.arange(B), threads)

    def work(idx):
        if len(idx) == 0:
            return 0
        xy = table[xy_tab_idx[idx]]  # makeBinaryMemorylessVectorDistribution, BinaryMemorylessDistribution.py:245-258
        oracle.bin_decode_batch(N, fm, r, xp, xy)
        return len(idx)

    oracle.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        done = sum(ex.map(work, parts))
    dt = time.perf_counter() - t0
    return done, dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python original cannot travel to the GPU box)."""
    if rank != 0:
        return
    import oracle
    N, K, fs = c1_code()
    fm = oracle.frozen_mask(N, fs)
    r = oracle.common_randomness(N, 1)
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1234)
    frames = args.ref_frames
    info = rng.integers(0, 2, size=(frames, K))
    cw = oracle.bin_encode_batch(N, fm, r, np.full((N, 2), 0.5), info[:min(frames, 2048)])
    cw = np.tile(cw, ((frames + cw.shape[0] - 1) // cw.shape[0], 1))[:frames]
    y = (cw ^ (rng.random((frames, N)) < P_BSC)).astype(np.uint8)
    tab = bsc_table()
    for _ in range(args.warmup):
        cpu_decode_sample(N, fm, r, y[:max(cores * 64, 256)], tab, cores)
    t = 0.0
    done = 0
    for _ in range(args.steps):
        d, dt = cpu_decode_sample(N, fm, r, y, tab, cores)
        t += dt
        done += d
    val = done * K / t / 1e9
    line = {
        "impl": "reference", "metric": "decoded info Gbit/s", "value": val, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "sc_n1024_k512_bsc0.11", "frames_per_step": frames},
        "frames_per_s": done / t,
        "cpu_baseline": {"value": val, "unit": "Gbit/s", "cores": cores, "kind": "port",
                         "sample": "%d frames/step x %d steps, oracle/polar_oracle.c (C restatement of the reference's "
                                   "float64 SC recursion, both prior and posterior trees), %d threads" % (frames, args.steps, cores)},
        "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from polarcub_b200 import engine
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N, K, fs = c1_code()
    fm = np.zeros(N, dtype=np.uint8)
    fm[list(fs)] = 1
    import random as _random
    rng_cr = _random.Random()
    rng_cr.seed(1)  # commonRandomnessSeed=1, BinaryPolarEncoderDecoder.py:36-41
    r = np.array([rng_cr.random() for _ in range(N)])
    fv = np.where(0.5 >= r, 0, 1).astype(np.uint8)
    plan = engine.Plan(2, 10, fm, fv, device=dev)
    tab = bsc_table()
    B = args.frames
    Be = min(args.e2e_frames, B)

    # ---- synthetic frames, generated on the device, keyed by the global frame index (rank-count invariant) ----
    gen = torch.Generator(device=dev)
    y = torch.empty((B, N), dtype=torch.uint8, device=dev)
    info_tx = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
    shifts = torch.arange(32, device=dev, dtype=torch.int32)
    CH = 1 << 16
    for c0 in range(0, B, CH):
        c1 = min(B, c0 + CH)
        gen.manual_seed(1234 + 7919 * ((rank * B + c0) // CH))
        it = torch.randint(-2 ** 31, 2 ** 31 - 1, (c1 - c0, plan.Kw), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
        info_tx[c0:c1] = it
        cwp = engine.encode_bits(plan, it.contiguous())
        bits = ((cwp.unsqueeze(-1) >> shifts) & 1).reshape(c1 - c0, N).to(torch.uint8)
        flips = (torch.rand((c1 - c0, N), device=dev, generator=gen) < P_BSC).to(torch.uint8)
        y[c0:c1] = bits ^ flips
        del it, cwp, bits, flips
    cw_out = torch.empty((B, plan.Nw), dtype=torch.int32, device=dev)
    info_out = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
    y_host = torch.empty((Be, N), dtype=torch.uint8).pin_memory()
    y_host.copy_(y[:Be])
    cw_host = torch.empty((Be, plan.Nw), dtype=torch.int32).pin_memory()
    info_host = torch.empty((Be, plan.Kw), dtype=torch.int32).pin_memory()
    y_e = torch.empty((Be, N), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        engine.sc_decode_symbols(plan, y, tab, out=(cw_out, info_out))

    def e2e_step():
        y_e.copy_(y_host, non_blocking=True)
        c, i = engine.sc_decode_symbols(plan, y_e, tab, out=(cw_out[:Be], info_out[:Be]))
        cw_host.copy_(c, non_blocking=True)
        info_host.copy_(i, non_blocking=True)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    engine.profile_enable(True)
    l0 = engine.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = engine.kernel_launch_count() - l0
    k_ms, k_launches = engine.profile_read()
    engine.profile_enable(False)
    clocks = sampler.stop(set(range(world)) if world > 1 else {local_rank}) if rank == 0 else None

    # ---- end to end through host buffers --------------------------------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- error counters: one kernel per rank + ONE all-reduce (the only collective of the run) ----------
    counters = engine.count_errors(info_out, info_tx, K)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    cnt = counters.cpu().numpy()
    if rank != 0:
        return

    frames_total = B * world * args.steps
    value = frames_total * K / (ms * 1e-3) / 1e9
    e2e_val = Be * world * args.steps * K / (ms_e2e * 1e-3) / 1e9
    peak, peak_kind = measured_peak_hbm()
    alg_bytes_frame = 4288  # 4 N bytes of soft input + (N + K)/8 bytes out
    achieved = (B * args.steps * alg_bytes_frame) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    line = {
        "metric": "decoded info Gbit/s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "sc_n1024_k512_bsc0.11", "N": N, "K": K, "decoder": "SC", "frames_per_step_per_gpu": B,
                   "input": "uint8 channel symbols [B,N] (1 GiB per step per GPU, larger than L2: no flush needed)",
                   "parity": "bit-identical to the reference (float64 probability-pair arithmetic)"},
        "frames_per_s": frames_total / (ms * 1e-3),
        "fer": float(cnt[1]) / max(1, int(cnt[0])), "ber": float(cnt[2]) / max(1, int(cnt[0]) * K),
        "e2e": {"value": e2e_val, "unit": "Gbit/s", "h2d_bytes_per_step": int(Be * N), "d2h_bytes_per_step": int(Be * (plan.Nw + plan.Kw) * 4),
                "frames_per_step_per_gpu": Be},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None,
                     "kernel": "sc_decode_kernel<symbols>", "kernel_ms_per_launch": k_ms / max(1, k_launches),
                     "kernel_launches": int(k_launches), "kernel_share_of_step": k_ms / ms,
                     "algorithmic_bytes_per_frame": alg_bytes_frame, "peak_source": peak_kind,
                     "note": "SC decoding is FP64-issue bound, not HBM bound (SURVEY.md 8d); see profiles/ for pipe utilisation"},
    }
    if world == 1:
        import oracle
        cores = os.cpu_count() or 1
        sample = min(B, args.cpu_frames)
        ys = y[:sample].cpu().numpy()
        oracle.build()
        cpu_decode_sample(N, fm, r, ys[:256], tab, cores)
        done, dt = cpu_decode_sample(N, fm, r, ys, tab, cores)
        line["cpu_baseline"] = {"value": done * K / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "port",
                                "frames_per_s": done / dt,
                                "sample": "first %d frames of the GPU batch, oracle/polar_oracle.c on %d threads "
                                          "(%.1f s wall)" % (sample, cores, dt)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sc1024")
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per step per GPU")
    ap.add_argument("--e2e-frames", type=int, default=1 << 18)
    ap.add_argument("--cpu-frames", type=int, default=1 << 16)
    ap.add_argument("--ref-frames", type=int, default=1 << 16)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
