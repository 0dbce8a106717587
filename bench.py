#!/usr/bin/env python
"""bench.py -- decoded info Gbit/s of the batched polar decoders on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload scl4096|sc1024|qsc2048|sc2p20|del256] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the decode hot path over one batch of synthetic channel outputs (frames are
independent, so every rank decodes its own slice: weak scaling, no data-path collective; one NCCL
all-reduce of the int64[3] error counters after the timed region).  Rank 0 prints ONE JSON line.

Default workload = BASELINE.json configs[1]: binary SCL L=8, N=4096, R=1/2 over BI-AWGN (Eb/N0 = 2 dB).
`--workload sc1024` is configs[0] (binary SC N=1024 K=512 over BSC(0.11)), `--workload qsc2048` configs[2].
The default line also carries a short `secondary` object with the SC N=1024 device-resident throughput, since
BASELINE.json's metric names both decoders.

  value     whole-job decoded information Gbit/s, inputs resident in HBM, CUDA events, max over ranks
  e2e       same metric through the batched C-ABI call with HOST (pinned) buffers: H2D of the channel
            probabilities / symbols and D2H of the decoded words inside the timed region
  roofline  dominant kernel (the SC/SCL decode kernel): algorithmic bytes per launch / its CUDA-event duration
  cpu_baseline  the oracle port (oracle/polar_oracle*.c, the reference's algorithm restated in C) on the host cores
"""
import argparse
import json
import math
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
EBN0_DB = 2.0
P_QSC = 0.02


def bsc_table(p=P_BSC):
    return np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])  # makeBSC, BinaryMemorylessDistribution.py:485-490


def awgn_sigma(rate=0.5, ebn0_db=EBN0_DB):
    return math.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0)))


def common_randomness(N, seed):
    import random as _random
    rng = _random.Random()
    rng.seed(seed)  # BinaryPolarEncoderDecoder.py:36-41
    return np.array([rng.random() for _ in range(N)])


def mask_of(N, fs):
    fm = np.zeros(N, dtype=np.uint8)
    fm[list(fs)] = 1
    return fm


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
            busy = [s for s in sm if s > 0]
            out["sm_mhz"] = float(np.median(busy or sm))
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


def run_threads(work, parts, threads):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        done = sum(ex.map(work, parts))
    return done, time.perf_counter() - t0


# ---------------------------------------------------------------------------------------------------
# Workloads.  Each one owns: the code (N, K, frozen set), device-side synthetic inputs keyed by the GLOBAL frame
# index (results do not depend on the rank count), the timed step, the e2e step and the CPU (oracle) leg.
# ---------------------------------------------------------------------------------------------------
class ScBinary1024:
    """C1: binary SC, N=1024, K=512, BSC(0.11); frozen set = 512 worst indices by Tal-Vardy Pe (reference degrade pass)."""
    name = "sc_n1024_k512_bsc0.11"
    kernel = "sc_decode_kernel<symbols>"
    dtype = "f64"
    default_frames, default_e2e, default_cpu = 1 << 20, 1 << 20, 1 << 16
    N, K, n = 1024, 512, 10
    alg_bytes_frame = 4288  # 4 N bytes of soft input + (N + K)/8 bytes out
    info_bits = 512
    ncu_key = "sc_decode_kernel_symbols"

    def code(self):
        from polarcub_b200.construction import frozen_set_from_pe, load_pe
        fs = frozen_set_from_pe(load_pe("bsc_p0.11_n10_L100_pe.npy"), self.K)
        self.fm = mask_of(self.N, fs)
        self.r = common_randomness(self.N, 1)
        self.fv = np.where(0.5 >= self.r, 0, 1).astype(np.uint8)
        self.tab = bsc_table()
        self.construction = "Tal-Vardy degrade L=100 via the live reference (tests/golden/constructions/bsc_p0.11_n10_L100_pe.npy)"

    def setup(self, dev, rank, B, Be):
        import torch
        from polarcub_b200 import channels, engine
        self.engine, self.torch = engine, torch
        self.code()
        N = self.N
        self.plan = plan = engine.Plan(2, self.n, self.fm, self.fv, device=dev)
        gen = torch.Generator(device=dev)
        self.y = torch.empty((B, N), dtype=torch.uint8, device=dev)
        self.info_tx = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        shifts = torch.arange(32, device=dev, dtype=torch.int32)
        CH = 1 << 16
        for c0 in range(0, B, CH):
            c1 = min(B, c0 + CH)
            gen.manual_seed(1234 + 7919 * ((rank * B + c0) // CH))
            it = torch.randint(-2 ** 31, 2 ** 31 - 1, (c1 - c0, plan.Kw), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
            self.info_tx[c0:c1] = it
            cwp = engine.encode_bits(plan, it.contiguous())
            # BSC(p) by the product's device simulator (csrc/channel.cu: counter-based noise keyed by the global frame index)
            self.y[c0:c1] = channels.simulate_dmc(cwp, [[1 - P_BSC, P_BSC], [P_BSC, 1 - P_BSC]], seed=1234, frame0=rank * B + c0,
                                                  packed_bits=N)
            del it, cwp
        self.cw_out = torch.empty((B, plan.Nw), dtype=torch.int32, device=dev)
        self.info_out = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        self.Be = Be
        # host side of the e2e leg: a BSC output symbol is one bit -- the host buffer holds the symbols bit-packed
        # (channels.pack_symbols layout), the device unpacks them in front of the decoder (pc_unpack_symbols)
        self.host_bits = channels.symbol_bits(self.tab.shape[0])
        self.y_host = torch.empty((Be, N * self.host_bits // 8), dtype=torch.uint8).pin_memory()
        for c0 in range(0, Be, 1 << 16):
            c1 = min(Be, c0 + (1 << 16))
            self.y_host[c0:c1].copy_(channels.pack_symbols(self.y[c0:c1].contiguous(), self.host_bits))
        self.cw_host = torch.empty((Be, plan.Nw), dtype=torch.int32).pin_memory()
        self.info_host = torch.empty((Be, plan.Kw), dtype=torch.int32).pin_memory()
        self.h2d = int(Be * N * self.host_bits // 8)
        self.d2h = int(Be * (plan.Nw + plan.Kw) * 4)
        self.input_note = ("uint8 channel symbols [B,N] (%.2f GiB per step per GPU, larger than L2: no flush needed); e2e: %d-bit packed "
                           "symbols in pinned host memory" % (B * N / 2 ** 30, self.host_bits))

    def step(self):
        self.engine.sc_decode_symbols(self.plan, self.y, self.tab, out=(self.cw_out, self.info_out))

    def e2e_step(self):
        # the public host-batch call: chunked, H2D / decode / D2H overlapped on two streams (engine.host_pipeline)
        self.engine.sc_decode_symbols_host(self.plan, self.y_host, self.tab, self.cw_host, self.info_host,
                                           packed_bits=self.host_bits)

    def counters(self):
        return self.engine.count_errors(self.info_out, self.info_tx, self.K)

    def gpu_info(self, sample):
        return self.engine.unpack_bits(self.info_out[:sample].cpu().numpy(), self.K).astype(np.int64)

    # ---- CPU leg (oracle port) ----
    def cpu_inputs_from_gpu(self, sample):
        return self.y[:sample].cpu().numpy()

    def cpu_inputs_synth(self, frames):
        import oracle
        self.code()
        rng = np.random.default_rng(1234)
        info = rng.integers(0, 2, size=(min(frames, 2048), self.K))
        cw = oracle.bin_encode_batch(self.N, self.fm, self.r, np.full((self.N, 2), 0.5), info)
        cw = np.tile(cw, ((frames + cw.shape[0] - 1) // cw.shape[0], 1))[:frames]
        return (cw ^ (rng.random((frames, self.N)) < P_BSC)).astype(np.uint8)

    def cpu_decode(self, ys, threads, out=None):
        import oracle
        oracle.lib()
        xp = np.full((self.N, 2), 0.5)

        def work(idx):
            if len(idx) == 0:
                return 0
            xy = self.tab[ys[idx]]  # makeBinaryMemorylessVectorDistribution, BinaryMemorylessDistribution.py:245-258
            _, info = oracle.bin_decode_batch(self.N, self.fm, self.r, xp, xy)
            if out is not None:
                out[idx] = info
            return len(idx)

        return run_threads(work, np.array_split(np.arange(ys.shape[0]), threads), threads)

    cpu_what = "oracle/polar_oracle.c (C restatement of the reference's float64 SC recursion, prior and posterior trees)"


class ScBinaryLarge:
    """C4: large-block binary SC, N=2^20, R=0.8 over BEC(0.1); upper stages streamed through HBM (sc_stream.cu)."""
    name = "sc_n2p%s_r0.8_bec0.1" % os.environ.get("PC_BENCH_LARGE_N", "20")
    kernel = ("hybrid walk (whole batch = one profiled unit): HBM-streamed upper stages hy_level8_kernel / hy_level_sym8_kernel on one-byte "
              "state codes (erasure-type channel) + sc_decode8_kernel (1024-leaf sub-blocks, frame per lane, exact rate-1 shortcut); "
              "other discrete channels: hy_level_kernel + sc_decode_kernel<packed> on float64; batches below 6144 frames: sc_stream_kernel")
    dtype = "f64"
    # one full wave of the sub-block kernel (148 SMs x 256 frames): its time does not depend on the batch size below that
    default_frames, default_e2e, default_cpu = 148 * 256, 32768, 32
    n = int(os.environ.get("PC_BENCH_LARGE_N", "20"))  # 20 is the BASELINE configuration; smaller values are for profiling runs
    N, K = 1 << n, int(0.8 * (1 << n))
    # stages above 2^13 stream 12 N bytes each + channel ingest 4 N + (N + K)/8 out (fp32 soft-input contract)
    alg_bytes_frame = 4 * N * (1 + 3 * max(0, n - 13)) + (N + K) // 8
    info_bits = K
    P_BEC = 0.1
    roofline_note = ("achieved = SURVEY.md 8d's algorithmic bytes (fp32 soft-input contract, stages above 2^13 streamed: 92.5 MB per "
                     "2^20 frame) / time of the whole hybrid walk (~4,000 launches, one profiled unit).  Over this erasure-type "
                     "channel the walk keeps ONE BYTE per tree element where the contract assumes a 4-byte value, so frac can "
                     "exceed 1; `design` is the traffic of the byte-state layout itself (~37 N bytes per frame: ingest 2 N, "
                     "level n-2 from the symbols 5 N, eight streamed levels 3 N each, sub-block input 2 N, decision bits / "
                     "partial sums / codeword and information egress ~4 N).  The float64 walk of the same code (PC_SC_HY8=0) "
                     "reaches 9.8 Gbit/s (profiles/r1_h_c4_hybrid_rate1.md)")
    design_bytes_frame = 37 * N

    def code(self):
        from polarcub_b200.construction import bec_pe
        pe = bec_pe(self.n, self.P_BEC)
        order = np.argsort(pe, kind="stable")
        self.fm = np.zeros(self.N, dtype=np.uint8)
        self.fm[order[self.K:]] = 1
        self.r = common_randomness(self.N, 1)
        self.fv = np.where(0.5 >= self.r, 0, 1).astype(np.uint8)
        p = self.P_BEC
        self.tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])  # makeBEC, BinaryMemorylessDistribution.py:493-499
        self.construction = "closed-form BEC recursion z -> (2z - z^2, z^2), K = 0.8 N best indices (SURVEY.md 8d, C4)"

    def setup(self, dev, rank, B, Be):
        import torch
        from polarcub_b200 import channels, engine
        self.engine, self.torch = engine, torch
        self.code()
        N = self.N
        self.plan = plan = engine.Plan(2, self.n, self.fm, self.fv, device=dev)
        gen = torch.Generator(device=dev)
        self.y = torch.empty((B, N), dtype=torch.uint8, device=dev)
        self.info_tx = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        shifts = torch.arange(32, device=dev, dtype=torch.int32)
        CH = 64
        for c0 in range(0, B, CH):
            c1 = min(B, c0 + CH)
            gen.manual_seed(2020 + 7919 * ((rank * B + c0) // CH))
            it = torch.randint(-2 ** 31, 2 ** 31 - 1, (c1 - c0, plan.Kw), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
            if self.K & 31:
                it[:, -1] &= (1 << (self.K & 31)) - 1
            self.info_tx[c0:c1] = it
            cwp = engine.encode_bits(plan, it.contiguous())
            e = self.P_BEC  # BEC(e): output 2 is the erasure (makeBEC, BinaryMemorylessDistribution.py:493-499)
            self.y[c0:c1] = channels.simulate_dmc(cwp, [[1 - e, 0.0, e], [0.0, 1 - e, e]], seed=2020, frame0=rank * B + c0,
                                                  packed_bits=N)
            del it, cwp
        self.cw_out = torch.empty((B, plan.Nw), dtype=torch.int32, device=dev)
        self.info_out = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        self.Be = Be
        # e2e: a BEC output symbol (0, 1, erasure) travels as 2 bits (channels.pack_symbols layout), unpacked on the device
        self.host_bits = channels.symbol_bits(self.tab.shape[0])
        self.y_host = torch.empty((Be, N * self.host_bits // 8), dtype=torch.uint8).pin_memory()
        for c0 in range(0, Be, 256):
            c1 = min(Be, c0 + 256)
            self.y_host[c0:c1].copy_(channels.pack_symbols(self.y[c0:c1].contiguous(), self.host_bits))
        self.cw_host = torch.empty((Be, plan.Nw), dtype=torch.int32).pin_memory()
        self.info_host = torch.empty((Be, plan.Kw), dtype=torch.int32).pin_memory()
        self.h2d = int(Be * N * self.host_bits // 8)
        self.d2h = int(Be * (plan.Nw + plan.Kw) * 4)
        self.input_note = ("uint8 channel symbols [B,N] (%.2f GiB per step per GPU, larger than L2); e2e: %d-bit packed symbols in "
                           "pinned host memory" % (B * N / 2 ** 30, self.host_bits))

    step = ScBinary1024.step
    e2e_step = ScBinary1024.e2e_step
    counters = ScBinary1024.counters
    gpu_info = ScBinary1024.gpu_info
    cpu_inputs_from_gpu = ScBinary1024.cpu_inputs_from_gpu
    cpu_decode = ScBinary1024.cpu_decode
    cpu_what = ScBinary1024.cpu_what

    def cpu_inputs_synth(self, frames):
        import oracle
        self.code()
        rng = np.random.default_rng(2020)
        info = rng.integers(0, 2, size=(min(frames, 4), self.K))
        cw = oracle.bin_encode_batch(self.N, self.fm, self.r, np.full((self.N, 2), 0.5), info)
        cw = np.tile(cw, ((frames + cw.shape[0] - 1) // cw.shape[0], 1))[:frames]
        return np.where(rng.random((frames, self.N)) < self.P_BEC, 2, cw).astype(np.uint8)


def awgn_frozen_set(n, K, allow_ga):
    """C2 code: Pe vector of the reference's degrade pass over a 400-bin quantised BI-AWGN (oracle/gen_constructions.py)."""
    from polarcub_b200.construction import frozen_set_from_pe, load_pe, ga_awgn_pe
    name = "biawgn_ebn0%s_n%d_L100_pe.npy" % (EBN0_DB, n)
    try:
        return frozen_set_from_pe(load_pe(name), K), "Tal-Vardy degrade L=100 via the live reference on a 400-bin BI-AWGN (tests/golden/constructions/%s)" % name
    except FileNotFoundError:
        if not allow_ga:
            raise
        return frozen_set_from_pe(ga_awgn_pe(n, awgn_sigma()), K), "Gaussian approximation (Chung et al.), --construction ga"


def awgn_quantiser(sigma, Y=256):
    """Uniform Y-level quantiser of the BI-AWGN output over [-1 - 4 sigma, 1 + 4 sigma] (outer bins open) and its channel
    table [Y, 2]: row s = (P(bin s, x=0), P(bin s, x=1)) for equiprobable inputs -- the rows a
    QaryMemorylessDistribution.makeQaryMemorylessVectorDistribution would copy into probs."""
    ymax = 1.0 + 4.0 * sigma
    step = 2.0 * ymax / Y
    edges = -ymax + step * np.arange(Y + 1)
    edges[0], edges[-1] = -np.inf, np.inf

    def cdf(x):
        return 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))

    tab = np.stack([0.5 * (cdf((edges[1:] - 1.0) / sigma) - cdf((edges[:-1] - 1.0) / sigma)),
                    0.5 * (cdf((edges[1:] + 1.0) / sigma) - cdf((edges[:-1] + 1.0) / sigma))], axis=-1)
    return ymax, step, np.ascontiguousarray(tab, dtype=np.float64)


class SclBinary4096:
    """C2: binary SCL L=8, N=4096, K=2048 over BI-AWGN at Eb/N0 = 2 dB, linear-domain float64 (listDecode with q=2).
    The channel output is quantised to 256 levels (one byte per symbol) and enters the decoder as symbols + channel table
    (pc_scl_decode_symbols = listDecode on makeQaryMemorylessVectorDistribution(length, y)); `records.float64_pairs` runs the
    same frames' unquantised outputs through the float64-pair entry point (round 1's headline input)."""
    name = "scl_l8_n4096_k2048_biawgn2dB"
    kernel = "sclp_kernel (binary SCL, one path per lane, 32/L frames per warp)"
    dtype = "f64"
    default_frames, default_e2e, default_cpu = 0, 0, 1 << 14  # frames: five resident waves of the decode kernel
    N, K, n, L = 4096, 2048, 12, 8
    alg_bytes_frame = 16640  # 4 N bytes of soft input + K/8 bytes out
    info_bits = 2048
    ncu_key = "sclp_kernel"
    allow_ga = False
    ebn0_db = EBN0_DB
    pairs_frames = 1 << 14

    def code(self):
        fs, self.construction = awgn_frozen_set(self.n, self.K, self.allow_ga)
        self.fm = mask_of(self.N, fs)
        self.sigma = awgn_sigma(ebn0_db=self.ebn0_db)
        self.ymax, self.qstep, self.tab = awgn_quantiser(self.sigma)

    def pairs(self, torch, y):
        s = self.sigma
        l0 = -(y - 1.0) ** 2 / (2 * s * s)
        l1 = -(y + 1.0) ** 2 / (2 * s * s)
        m = torch.maximum(l0, l1)
        return torch.stack([torch.exp(l0 - m), torch.exp(l1 - m)], dim=-1).contiguous()

    def setup(self, dev, rank, B, Be):
        import torch
        from polarcub_b200 import channels, engine
        self.engine, self.torch = engine, torch
        self.code()
        N, K = self.N, self.K
        self.plan = plan = engine.Plan(2, self.n, self.fm, None, device=dev)
        wave = engine.scl_wave_frames(plan, self.L)
        if not B:
            B = 5 * wave
        if not Be:
            Be = B
        Be = min(Be, B)
        self.B, self.Be = B, Be
        gen = torch.Generator(device=dev)
        self.ys = torch.empty((B, N), dtype=torch.uint8, device=dev)
        self.info_tx = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        npairs = min(self.pairs_frames, B)
        self.xy = torch.empty((npairs, N, 2), dtype=torch.float64, device=dev)
        shifts = torch.arange(32, device=dev, dtype=torch.int32)
        CH = 1 << 12
        for c0 in range(0, B, CH):
            c1 = min(B, c0 + CH)
            gen.manual_seed(4321 + 7919 * ((rank * B + c0) // CH))
            it = torch.randint(-2 ** 31, 2 ** 31 - 1, (c1 - c0, plan.Kw), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
            self.info_tx[c0:c1] = it
            bits = ((it.unsqueeze(-1) >> shifts) & 1).reshape(c1 - c0, K).to(torch.uint8)
            cw = engine.qsc_encode(plan, bits.contiguous())
            yq, y = channels.simulate_biawgn(cw, self.sigma, seed=4321, frame0=rank * B + c0, levels=self.tab.shape[0],
                                             ymax=self.ymax, want_real=c0 < npairs)
            self.ys[c0:c1] = yq
            if c0 < npairs:
                m = min(c1, npairs) - c0
                self.xy[c0:c0 + m] = self.pairs(torch, y[:m])
            del it, bits, cw, y, yq
        self.info_out = torch.empty((B, plan.Kw), dtype=torch.int32, device=dev)
        self.res_out = torch.empty((B,), dtype=torch.int32, device=dev)
        self.y_host = torch.empty((Be, N), dtype=torch.uint8).pin_memory()
        self.y_host.copy_(self.ys[:Be])
        self.ai_host = torch.empty((Be, plan.Kw), dtype=torch.int32).pin_memory()
        self.ai_host.copy_(self.info_tx[:Be])
        self.info_host = torch.empty((Be, plan.Kw), dtype=torch.int32).pin_memory()
        self.res_host = torch.empty((Be,), dtype=torch.int32).pin_memory()
        self.h2d = int(Be * (N + plan.Kw * 4))
        self.d2h = int(Be * (plan.Kw * 4 + 4))
        self.input_note = ("uint8 channel symbols [B,N] (BI-AWGN output quantised to 256 levels) + the [256,2] channel table; "
                           "actual information bit-packed; %.2f GiB of symbols per step per GPU (device-resident run: inputs "
                           "larger than L2, no flush needed)" % (B * N / 2 ** 30))

    def step(self):
        self.engine.scl_decode_packed(self.plan, self.L, self.info_tx, y=self.ys, table=self.tab, out=(self.info_out, self.res_out))

    def e2e_step(self):
        # the public host-batch call: chunks of one resident wave, H2D / decode / D2H overlapped on three streams
        self.engine.scl_decode_symbols_host(self.plan, self.L, self.y_host, self.tab, self.ai_host, self.info_host, self.res_host)

    def counters(self):
        return self.engine.count_errors(self.info_out, self.info_tx, self.K)

    def prob_result_hist(self):
        return self.torch.bincount(self.res_out.clamp(min=0), minlength=6).cpu().numpy().tolist()

    # ---- second record: the unquantised outputs of the first frames as float64 pairs ----
    def pairs_setup(self):
        torch, plan = self.torch, self.plan
        m = self.xy.shape[0]
        self.p_info = torch.empty((m, plan.Kw), dtype=torch.int32, device=self.xy.device)
        self.p_res = torch.empty((m,), dtype=torch.int32, device=self.xy.device)
        self.xy_host = torch.empty((m, self.N, 2), dtype=torch.float64).pin_memory()
        self.xy_host.copy_(self.xy)
        self.p_ai_host = torch.empty((m, plan.Kw), dtype=torch.int32).pin_memory()
        self.p_ai_host.copy_(self.info_tx[:m])
        self.p_info_host = torch.empty((m, plan.Kw), dtype=torch.int32).pin_memory()
        self.p_res_host = torch.empty((m,), dtype=torch.int32).pin_memory()

    def pairs_step(self):
        m = self.xy.shape[0]
        self.engine.scl_decode_packed(self.plan, self.L, self.info_tx[:m], xy=self.xy, out=(self.p_info, self.p_res))

    def pairs_e2e_step(self):
        self.engine.scl_decode_packed_host(self.plan, self.L, self.xy_host, self.p_ai_host, self.p_info_host, self.p_res_host)

    def cpu_inputs_from_gpu(self, sample):
        ys = self.ys[:sample].cpu().numpy()
        info = self.engine.unpack_bits(self.info_tx[:sample].cpu().numpy(), self.K).astype(np.int64)
        return (self.tab[ys], info)  # makeQaryMemorylessVectorDistribution, QaryMemorylessDistribution.py:757-776

    def gpu_info(self, sample):
        return self.engine.unpack_bits(self.info_out[:sample].cpu().numpy(), self.K).astype(np.int64)

    def cpu_inputs_synth(self, frames):
        import oracle
        self.code()
        rng = np.random.default_rng(4321)
        info = rng.integers(0, 2, size=(frames, self.K))
        u = np.zeros((frames, self.N), dtype=np.int64)
        u[:, self.fm == 0] = info
        cw = np.stack([oracle.polar_transform_qudits(2, u[b]) for b in range(frames)])
        y = (1.0 - 2.0 * cw) + self.sigma * rng.standard_normal(cw.shape)
        ys = np.clip(np.floor((y + self.ymax) / self.qstep), 0, self.tab.shape[0] - 1).astype(np.int64)
        return (self.tab[ys], info)

    def cpu_decode(self, inputs, threads, out=None):
        import oracle
        oracle.lib()
        xy, info = inputs
        fv = np.zeros((xy.shape[0], self.N - self.K), dtype=np.int64)

        def work(idx):
            if len(idx) == 0:
                return 0
            dinfo, _ = oracle.list_decode_batch(2, self.N, self.L, self.fm, xy[idx], fv[idx], info[idx])
            if out is not None:
                out[idx] = dinfo
            return len(idx)

        return run_threads(work, np.array_split(np.arange(xy.shape[0]), threads), threads)

    cpu_what = "oracle/polar_oracle_list.c (C restatement of the reference's float64 listDecode with fast nodes)"


class ScQary2048:
    """C3: q=3 SC, N=2048, K=1024 over the ternary symmetric channel QSC(p=0.02)."""
    name = "qsc_q3_n2048_k1024_qsc0.02"
    kernel = "qsc_decode_kernel<3>"
    dtype = "f64"
    default_frames, default_e2e, default_cpu = 151552, 151552, 1 << 11
    N, K, n, q = 2048, 1024, 11, 3
    alg_bytes_frame = 25344
    ncu_key = "qsc_decode_kernel_q3"
    info_bits = 1024 * math.log2(3)
    allow_ga = False

    def code(self):
        from polarcub_b200.construction import frozen_set_from_pe, load_pe, bec_pe
        name = "qsc_q3_p%s_n%d_L100_pe.npy" % (P_QSC, self.n)
        try:
            fs = frozen_set_from_pe(load_pe(name), self.K)
            self.construction = "reference calcTVAndPe_degradingUpgrading L=100 (tests/golden/constructions/%s)" % name
        except FileNotFoundError:
            if not self.allow_ga:
                raise
            fs = frozen_set_from_pe(bec_pe(self.n, 0.15), self.K)
            self.construction = ("BEC(0.15) Bhattacharyya heuristic (the reference-derived Pe vector %s is not under "
                                 "tests/golden/constructions/: its degrade/upgrade pass takes hours; --construction reference insists on it)" % name)
        self.fm = mask_of(self.N, fs)

    def setup(self, dev, rank, B, Be):
        import torch
        from polarcub_b200 import channels, engine
        self.engine, self.torch = engine, torch
        self.code()
        N, K, q = self.N, self.K, self.q
        self.plan = plan = engine.Plan(q, self.n, self.fm, None, device=dev)
        gen = torch.Generator(device=dev)
        self.y = torch.empty((B, N), dtype=torch.uint8, device=dev)
        self.info_tx = torch.empty((B, K), dtype=torch.uint8, device=dev)
        self.tab = np.full((q, q), P_QSC / (q - 1), dtype=np.float64)  # makeQSC, QaryMemorylessDistribution.py:780-784
        np.fill_diagonal(self.tab, 1.0 - P_QSC)
        CH = 1 << 13
        for c0 in range(0, B, CH):
            c1 = min(B, c0 + CH)
            gen.manual_seed(777 + 7919 * ((rank * B + c0) // CH))
            it = torch.randint(0, q, (c1 - c0, K), dtype=torch.uint8, device=dev, generator=gen)
            self.info_tx[c0:c1] = it
            cw = engine.qsc_encode(plan, it.contiguous())
            self.y[c0:c1] = channels.simulate_dmc(cw, channels.conditional_table(self.tab), seed=777, frame0=rank * B + c0)
            del it, cw
        self.out = None
        self.Be = Be
        self.y_host = torch.empty((Be, N), dtype=torch.uint8).pin_memory()
        self.y_host.copy_(self.y[:Be])
        self.info_host = torch.empty((Be, K), dtype=torch.uint8).pin_memory()
        self.h2d = int(Be * N)
        self.d2h = int(Be * K)
        self.input_note = ("uint8 channel symbols [B,N] + the [3,3] channel table (makeQaryMemorylessVectorDistribution fused into the "
                           "ingest; %.2f GiB of float64 triples per step per GPU after expansion, larger than L2)" % (B * N * q * 8 / 2 ** 30))

    def step(self):
        self.out = self.engine.qsc_decode_symbols(self.plan, self.y, self.tab)

    def e2e_step(self):
        self.engine.qsc_decode_symbols_host(self.plan, self.y_host, self.tab, self.info_host)

    def counters(self):
        torch = self.torch
        diff = (self.out[1] != self.info_tx)
        c = torch.zeros(3, dtype=torch.int64, device=diff.device)
        c[0] = diff.shape[0]
        c[1] = diff.any(dim=1).sum()
        c[2] = diff.sum()
        return c

    def cpu_inputs_from_gpu(self, sample):
        return self.tab[self.y[:sample].cpu().numpy()]  # makeQaryMemorylessVectorDistribution, QaryMemorylessDistribution.py:757-766

    def gpu_info(self, sample):
        return self.out[1][:sample].cpu().numpy().astype(np.int64)

    def cpu_inputs_synth(self, frames):
        import oracle
        self.code()
        q = self.q
        rng = np.random.default_rng(777)
        info = rng.integers(0, q, size=(frames, self.K))
        u = np.zeros((frames, self.N), dtype=np.int64)
        u[:, self.fm == 0] = info
        cw = np.stack([oracle.polar_transform_qudits(q, u[b]) for b in range(frames)])
        err = rng.random(cw.shape) < P_QSC
        y = np.where(err, (cw + rng.integers(1, q, size=cw.shape)) % q, cw)
        tab = np.full((q, q), P_QSC / (q - 1))
        np.fill_diagonal(tab, 1 - P_QSC)
        return tab[y]

    def cpu_decode(self, xy, threads, out=None):
        import oracle
        oracle.lib()
        xp = np.full((self.N, self.q), 1.0 / self.q)

        def work(idx):
            if len(idx) == 0:
                return 0
            _, dinfo = oracle.q_decode_batch(self.q, self.N, self.fm, xp, xy[idx])
            if out is not None:
                out[idx] = dinfo
            return len(idx)

        return run_threads(work, np.array_split(np.arange(xy.shape[0]), threads), threads)

    cpu_what = "oracle/polar_oracle.c (C restatement of the reference's float64 q-ary SC recursion)"


class DeletionTrellis256:
    """C5: deletion-channel SC decoding over trellis collections with guard bands, N = 256, n0 = 2 (main_deletion.py
    defaults: deletion probability 0.1, xi = 0.1, all-zero guard bands).  The frozen set comes from the GENIE pass
    (genieEncodeDecodeSimulation, 2000 trials, error bound 0.1) run on this stack, whose output is identical to the
    reference's (tests/test_gpu_genie.py)."""
    name = "deletion_trellis_n256_n02_d0.1"
    kernel = "trellis_step_kernel + sc_decode_kernel<probs> (batch-wide top-tree walk)"
    dtype = "f64"
    default_frames, default_e2e, default_cpu = 1 << 15, 1 << 15, 1 << 10
    n, n0, N = 8, 2, 256
    delta, xi, ones = 0.1, 0.1, 0
    GENIE_TRIALS = 2000

    def code(self):
        import contextlib
        import io
        import random
        import polarcub_b200 as pcb
        from polarcub_b200 import Guardbands, BinaryTrellis
        from polarcub_b200.CollectionOfBinaryTrellises import buildCollectionOfBinaryTrellises_uniformInput_deletion as build
        n, n0, N = self.n, self.n0, self.N
        chan = random.Random()
        chan.seed(100)
        with contextlib.redirect_stdout(io.StringIO()):
            fs = pcb.genieEncodeDecodeSimulation(
                N, lambda: np.full((N, 2), 0.5), lambda enc: Guardbands.addDeletionGuardBands([int(b) for b in enc], n, n0, self.xi, self.ones),
                lambda cw: BinaryTrellis.deletionChannelSimulation(cw, self.delta, seed=None, randomNumberGenerator=chan),
                lambda rw: build(rw, self.delta, self.xi, n, n0, self.ones), self.GENIE_TRIALS, 0.1, 300, trustXYProbs=False)
        self.fs = fs
        self.fm = mask_of(N, fs)
        self.K = int(N - self.fm.sum())
        self.info_bits = self.K
        self.r = common_randomness(N, 200)
        self.construction = "genie pass, %d trials, error bound 0.1, seeds 100/300 (main_deletion.py defaults; identical to the reference's genie)" % self.GENIE_TRIALS

    def setup(self, dev, rank, B, Be):
        import torch
        import polarcub_b200 as pcb
        from polarcub_b200 import engine, Guardbands
        self.engine, self.torch = engine, torch
        self.code()
        N, K, n, n0 = self.N, self.K, self.n, self.n0
        self.ed = pcb.BinaryPolarEncoderDecoder(N, self.fs, 200)
        self.plan = self.ed.plan
        rng = np.random.default_rng(5150 + rank)
        info = rng.integers(0, 2, size=(B, K))
        enc = self.ed.encode_batch(info)
        starts, total = Guardbands.guard_band_layout(n, n0, self.xi, self.ones)
        sub = (1 << n0)
        tx = np.zeros((B, total), dtype=np.uint8)
        for t, s0 in enumerate(starts):
            tx[:, s0:s0 + sub] = enc[:, t * sub:(t + 1) * sub]
        keep = rng.random((B, total)) >= self.delta
        rxs = [tx[b][keep[b]] for b in range(B)]
        bits, lens = Guardbands.split_batch(rxs, n, n0)
        self.alg_bytes_frame = int(bits.shape[1] * bits.shape[2] + 4 * lens.shape[1] + (N + K) // 8)
        self.info_np = info
        self.bits_np, self.lens_np = bits, lens
        self.bits = torch.from_numpy(bits).to(dev)
        self.lens = torch.from_numpy(lens).to(dev)
        self.info_tx = torch.from_numpy(engine.pack_bits(info).view(np.int32)).to(dev)
        self.Be = Be
        self.bits_host = torch.from_numpy(bits[:Be]).pin_memory()
        self.lens_host = torch.from_numpy(lens[:Be]).pin_memory()
        self.cw_host = torch.empty((Be, self.plan.Nw), dtype=torch.int32).pin_memory()
        self.info_host = torch.empty((Be, self.plan.Kw), dtype=torch.int32).pin_memory()
        self.h2d = int(Be * (bits.shape[1] * bits.shape[2] + 4 * lens.shape[1]))
        self.d2h = int(Be * (self.plan.Nw + self.plan.Kw) * 4)
        self.input_note = ("trimmed sub-words uint8 [B,%d,%d] + lengths (received words split at the guard bands on the host; "
                           "%.3f GiB per step per GPU)" % (bits.shape[1], bits.shape[2], B * self.alg_bytes_frame / 2 ** 30))

    def step(self):
        self.cw_out, self.info_out = self.engine.trellis_decode(self.plan, self.n0, self.delta, self.ones, self.bits, self.lens)

    def e2e_step(self):
        eng, plan = self.engine, self.plan
        sl = eng._Slots(plan, "del")
        chunk = 8192

        def body(lo, hi, slot):
            m = hi - lo
            b = sl.get(slot, "bits", (chunk,) + tuple(self.bits_host.shape[1:]), self.torch.uint8)[:m]
            ln = sl.get(slot, "lens", (chunk, self.lens_host.shape[1]), self.torch.int32)[:m]
            b.copy_(self.bits_host[lo:hi], non_blocking=True)
            ln.copy_(self.lens_host[lo:hi], non_blocking=True)
            cw, info = eng.trellis_decode(plan, self.n0, self.delta, self.ones, b, ln)
            self.cw_host[lo:hi].copy_(cw, non_blocking=True)
            self.info_host[lo:hi].copy_(info, non_blocking=True)

        eng.host_pipeline(plan, self.Be, chunk, body)

    def counters(self):
        return self.engine.count_errors(self.info_out.contiguous(), self.info_tx, self.K)

    def gpu_info(self, sample):
        return self.engine.unpack_bits(self.info_out[:sample].cpu().numpy(), self.K).astype(np.int64)

    def cpu_inputs_from_gpu(self, sample):
        return (self.bits_np[:sample], self.lens_np[:sample])

    def cpu_inputs_synth(self, frames):
        from polarcub_b200 import Guardbands
        self.code()
        rng = np.random.default_rng(5150)
        enc = rng.integers(0, 2, size=(frames, self.N)).astype(np.uint8)  # decoding work does not depend on the codeword
        starts, total = Guardbands.guard_band_layout(self.n, self.n0, self.xi, self.ones)
        sub = 1 << self.n0
        tx = np.zeros((frames, total), dtype=np.uint8)
        for t, s0 in enumerate(starts):
            tx[:, s0:s0 + sub] = enc[:, t * sub:(t + 1) * sub]
        keep = rng.random((frames, total)) >= self.delta
        return Guardbands.split_batch([tx[b][keep[b]] for b in range(frames)], self.n, self.n0)

    def cpu_decode(self, inputs, threads, out=None):
        import oracle
        oracle.lib()
        bits, lens = inputs

        def work(idx):
            for f in idx:
                _, info = oracle.trellis_decode(self.n, self.n0, self.fm, self.r, bits[f], lens[f], self.delta, self.ones)
                if out is not None:
                    out[f] = info
            return len(idx)

        return run_threads(work, np.array_split(np.arange(bits.shape[0]), threads), threads)

    cpu_what = "oracle/polar_oracle_trellis.c (C restatement of BinaryTrellis / CollectionOfBinaryTrellises under decode)"


WORKLOADS = {"scl4096": SclBinary4096, "sc1024": ScBinary1024, "qsc2048": ScQary2048, "sc2p20": ScBinaryLarge,
             "del256": DeletionTrellis256}


def nframes(x):
    return x[0].shape[0] if isinstance(x, tuple) else x.shape[0]


def head(x, m):
    return tuple(a[:m] for a in x) if isinstance(x, tuple) else x[:m]


# ---------------------------------------------------------------------------------------------------
def wilson(errors, frames, z=1.959963984540054):
    """95 % Wilson score interval of an error rate (the reference prints a bare ratio, BinaryPolarEncoderDecoder.py:387)."""
    if frames <= 0:
        return [0.0, 1.0]
    ph = errors / frames
    d = 1.0 + z * z / frames
    c = ph + z * z / (2 * frames)
    h = z * math.sqrt(ph * (1 - ph) / frames + z * z / (4.0 * frames * frames))
    return [max(0.0, (c - h) / d), min(1.0, (c + h) / d)]


def source_sha(files):
    import hashlib
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, "polarcub_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_figures(key):
    """Figures of the committed `ncu --set full` capture of a kernel (profiles/r2_ncu.json, written by scripts/ncu_to_json.py
    from the .ncu-rep of the same bench command).  Each entry records a hash of the kernel's source files; when the sources
    have changed since the capture the figures are refused (None) instead of being quoted stale."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu.json")) as f:
            e = json.load(f).get(key)
    except Exception:
        return None, "profiles/r2_ncu.json missing"
    if not e:
        return None, "no capture of %s in profiles/r2_ncu.json" % key
    try:
        now = source_sha(e["source_files"])
    except Exception:
        return None, "source files of the capture not found"
    if now != e["source_sha"]:
        return None, "capture %s is of other sources (%s, now %s): refused as stale" % (e.get("capture"), e["source_sha"], now)
    return e, None


def config_of(w):
    """Identical in both arms (ours / reference)."""
    return {"workload": w.name, "N": w.N, "K": int(w.K), "construction": w.construction,
            "parity": "bit-identical to the reference (float64 probability arithmetic, same rounding order)"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (the oracle PORT of it in C; the Python original cannot travel to the
    GPU box), all host threads, a bounded sample of the workload per step."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    w = WORKLOADS[args.workload]()
    w.allow_ga = args.construction != "reference"
    cores = os.cpu_count() or 1
    frames = args.ref_frames or w.default_cpu
    inputs = w.cpu_inputs_synth(frames)
    for _ in range(args.warmup):
        w.cpu_decode(head(inputs, max(cores, frames // 8)), cores)
    t, done = 0.0, 0
    for _ in range(args.steps):
        d, dt = w.cpu_decode(inputs, cores)
        t += dt
        done += d
    val = done * w.info_bits / t / 1e9
    line = {
        "impl": "reference", "metric": "decoded info Gbit/s", "value": val, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": w.dtype, "data": "synthetic",
        "config": config_of(w), "frames_per_step": frames,
        "frames_per_s": done / t,
        "cpu_baseline": {"value": val, "unit": "Gbit/s", "cores": cores, "kind": "port",
                         "sample": "%d frames/step x %d steps, %s, %d threads" % (frames, args.steps, w.cpu_what, cores)},
        "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def time_steps(torch, fn, steps, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def measure(w, args, torch, engine, dist, world, rank, local_rank, dev, B, Be, with_cpu):
    """One full record of a workload: device-resident value, e2e through host buffers, roofline of the dominant kernel, error
    counters (+ the CPU leg and the frame-by-frame parity check at N = 1).  Every rank runs it; rank 0 returns the record."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        w.step()
    barrier()
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    engine.profile_enable(True)
    l0 = engine.kernel_launch_count()
    ms = time_steps(torch, w.step, args.steps, barrier)
    launches = engine.kernel_launch_count() - l0
    k_ms, k_launches = engine.profile_read()
    engine.profile_enable(False)
    clocks = sampler.stop(set(range(world)) if world > 1 else {local_rank}) if rank == 0 else None

    # ---- end to end through host buffers --------------------------------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        w.e2e_step()
    ms_e2e = time_steps(torch, w.e2e_step, args.steps, barrier)

    # ---- error counters: one reduction per rank + ONE all-reduce (the only collective of the run) --------
    w.step()
    counters = w.counters()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    cnt = counters.cpu().numpy()
    if rank != 0:
        return None

    frames_total = B * world * args.steps
    bits = w.info_bits
    value = frames_total * bits / (ms * 1e-3) / 1e9
    e2e_val = Be * world * args.steps * bits / (ms_e2e * 1e-3) / 1e9
    peak, peak_kind = measured_peak_hbm()
    achieved = (B * args.steps * w.alg_bytes_frame) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    frames_counted, ferr = int(cnt[0]), int(cnt[1])
    rec = {
        "metric": "decoded info Gbit/s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": w.dtype, "data": "synthetic",
        "config": config_of(w), "frames_per_step_per_gpu": B, "input": w.input_note,
        "frames_per_s": frames_total / (ms * 1e-3),
        "fer": ferr / max(1, frames_counted), "fer_ci95": wilson(ferr, frames_counted), "frames_counted": frames_counted,
        "ber": float(cnt[2]) / max(1, frames_counted * w.K),
        "e2e": {"value": e2e_val, "unit": "Gbit/s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h,
                "frames_per_step_per_gpu": Be},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None,
                     "kernel": w.kernel, "kernel_ms_per_launch": k_ms / max(1, k_launches),
                     "kernel_launches": int(k_launches), "kernel_share_of_step": k_ms / ms,
                     "algorithmic_bytes_per_frame": w.alg_bytes_frame, "peak_source": peak_kind,
                     "note": getattr(w, "roofline_note", None) or
                             "float64 SC/SCL decoding is bound by instruction issue and the latency of the per-frame scratch, "
                             "not by the algorithmic HBM bytes (SURVEY.md 8d); see `issue` / `traffic` and profiles/"},
    }
    if hasattr(w, "prob_result_hist"):
        rec["prob_result_hist"] = w.prob_result_hist()
    if getattr(w, "design_bytes_frame", None) and k_ms > 0:
        d_ach = (B * args.steps * w.design_bytes_frame) / (k_ms * 1e-3) / 1e9
        rec["roofline"]["design"] = {"bytes_per_frame": w.design_bytes_frame, "achieved": d_ach, "unit": "GB/s", "frac": d_ach / peak}
    key = getattr(w, "ncu_key", None)
    if key and k_ms > 0:
        # figures of the committed `ncu --set full` capture of this kernel, scaled to this run's launches:
        # traffic = measured DRAM bytes per launch; issue = warp instructions issued per second vs 4 schedulers x SMs x clock
        ncu, why = ncu_figures(key)
        r = rec["roofline"]
        if ncu is None:
            r["traffic_note"] = why
        else:
            fpl = B * args.steps / max(1, k_launches)
            r["traffic"] = ncu["dram_bytes_per_frame"] * fpl
            r["traffic_bytes_per_frame"] = ncu["dram_bytes_per_frame"]
            r["traffic_frac_of_peak"] = ncu["dram_bytes_per_frame"] * B * args.steps / (k_ms * 1e-3) / 1e9 / peak
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            ipeak = 4 * 148 * sm_mhz * 1e6
            iach = ncu["warp_inst_per_frame"] * B * args.steps / (k_ms * 1e-3)
            r["issue"] = {"achieved": iach, "peak": ipeak, "unit": "warp inst/s", "frac": iach / ipeak,
                          "ncu_issue_active_pct": ncu["issue_active_pct"], "capture": ncu["capture"], "source_sha": ncu["source_sha"]}
    if with_cpu:
        import oracle
        cores = os.cpu_count() or 1
        sample = min(B, args.cpu_frames or w.default_cpu)
        oracle.build()
        inputs = w.cpu_inputs_from_gpu(sample)
        w.cpu_decode(head(inputs, max(cores, sample // 8)), cores)
        cpu_info = np.full((sample, w.K), -1, dtype=np.int64)
        done, dt = w.cpu_decode(inputs, cores, out=cpu_info)
        # FER parity on identical channel outputs: the CPU port's decisions against the GPU's, frame by frame
        same = (w.gpu_info(sample) == cpu_info).all(axis=1)
        rec["parity_check"] = {"frames_compared": int(sample), "identical": int(same.sum()), "unexplained": int((~same).sum()),
                               "against": "oracle port on the same channel outputs (bit-exact bar, no tie tolerance needed)"}
        rec["cpu_baseline"] = {"value": done * bits / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "port",
                               "frames_per_s": done / dt,
                               "sample": "first %d frames of the GPU batch, %s on %d threads (%.1f s wall)" % (
                                   sample, w.cpu_what, cores, dt)}
    return rec


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from polarcub_b200 import engine
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = WORKLOADS[args.workload]()
    w.allow_ga = args.construction != "reference"
    B = args.frames or w.default_frames
    Be = min(args.e2e_frames or w.default_e2e, B) if B else args.e2e_frames
    if world > 1 and args.workload == "sc2p20" and not args.e2e_frames:
        Be = min(Be, 8192)  # pinned host staging is per rank: 8 GiB each instead of 32
    w.setup(dev, rank, B, Be)
    B, Be = getattr(w, "B", B), getattr(w, "Be", Be)
    torch.cuda.synchronize()
    line = measure(w, args, torch, engine, dist, world, rank, local_rank, dev, B, Be, with_cpu=(world == 1))

    records = {}
    if args.workload == "scl4096" and not args.no_secondary:
        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # (1) the same code on float64 probability pairs (xyVectorDistribution.probs, round 1's input): device-resident and
        #     end to end (64 KiB per frame over PCIe)
        w.pairs_setup()
        for _ in range(2):
            w.pairs_step()
        st = max(3, args.steps // 4)
        ms_p = time_steps(torch, w.pairs_step, st, barrier)
        w.pairs_e2e_step()
        ms_pe = time_steps(torch, w.pairs_e2e_step, st, barrier)
        tt = torch.tensor([ms_p, ms_pe], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        m = w.xy.shape[0]
        if rank == 0:
            records["float64_pairs"] = {
                "input": "float64 probability pairs [B,N,2] of the unquantised channel outputs of the first %d frames" % m,
                "value": st * m * world * w.info_bits / (float(tt[0]) * 1e-3) / 1e9, "unit": "Gbit/s",
                "frames_per_s": st * m * world / (float(tt[0]) * 1e-3), "steps": st, "frames_per_step_per_gpu": m,
                "e2e": {"value": st * m * world * w.info_bits / (float(tt[1]) * 1e-3) / 1e9, "unit": "Gbit/s",
                        "h2d_bytes_per_step": int(m * (w.N * 16 + w.plan.Kw * 4)), "d2h_bytes_per_step": int(m * (w.plan.Kw * 4 + 4))}}
        del w.xy, w.xy_host
        # (2) a noisier operating point: more frames end outside the list, so the genie replay of the final kernel runs
        w2 = SclBinary4096()
        w2.ebn0_db, w2.pairs_frames, w2.allow_ga = 1.0, 0, w.allow_ga
        w2.setup(dev, rank, 2 * engine.scl_wave_frames(w.plan, w.L), 1024)
        for _ in range(2):
            w2.step()
        ms_n = time_steps(torch, w2.step, st, barrier)
        c2 = w2.counters()
        tt = torch.tensor([ms_n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(c2, op=dist.ReduceOp.SUM)
        if rank == 0:
            c2 = c2.cpu().numpy()
            records["ebn0_1dB"] = {"value": st * w2.B * world * w2.info_bits / (float(tt[0]) * 1e-3) / 1e9, "unit": "Gbit/s",
                                   "frames_per_s": st * w2.B * world / (float(tt[0]) * 1e-3), "steps": st,
                                   "frames_per_step_per_gpu": w2.B, "fer": int(c2[1]) / max(1, int(c2[0])),
                                   "fer_ci95": wilson(int(c2[1]), int(c2[0])), "prob_result_hist": w2.prob_result_hist(),
                                   "note": "Eb/N0 = 1 dB, same code: frames whose actual word is not in the final list take the "
                                           "genie replay of sclp_final_kernel (ProbResult >= 2)"}
        del w2
        # (3) BASELINE.json's metric also names SC N=1024: a FULL second record (value, e2e, roofline, cpu_baseline, parity)
        del w
        torch.cuda.empty_cache()
        s = ScBinary1024()
        s.allow_ga = args.construction != "reference"
        Bs = s.default_frames
        s.setup(dev, rank, Bs, Bs)
        rec = measure(s, args, torch, engine, dist, world, rank, local_rank, dev, Bs, Bs, with_cpu=(world == 1))
        if rank == 0:
            records["sc_n1024"] = rec
    if rank == 0:
        if records:
            line["records"] = records
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scl4096", choices=sorted(WORKLOADS))
    ap.add_argument("--construction", default="auto", choices=["auto", "reference", "ga"],
                    help="ga: allow a heuristic frozen set when the reference-derived Pe vector is not in tests/golden/constructions")
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (0 = workload default)")
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--ref-frames", type=int, default=0)
    ap.add_argument("--no-secondary", action="store_true")
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
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=300))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
