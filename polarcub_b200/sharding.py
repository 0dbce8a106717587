"""Frame sharding across ranks and the one collective of a Monte-Carlo run.

Frames are independent (the reference's loop carries no state between iterations, BinaryPolarEncoderDecoder.py:354,
QaryPolarEncoderDecoder.py:896,961), so a batch is partitioned contiguously, one slice per rank, with NO data-path
collective.  The only exchange is one all-reduce(sum) of the int64 counters {frames, frame errors, symbol errors}
(+ the 6-bin ProbResult histogram for list decoding, QaryPolarEncoderDecoder.py:18-24) at the end of a run -- NCCL over
NVLink on the GPU box, gloo in the CPU tests.  Synthetic inputs are keyed by the GLOBAL frame index, so results do not
depend on the number of ranks.
"""
import numpy as np


def shard_range(total_frames, rank, world_size):
    """Contiguous slice [begin, end) of `total_frames` owned by `rank`; slices differ in length by at most one."""
    assert 0 <= rank < world_size and total_frames >= 0
    base, rem = divmod(int(total_frames), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def chunk_seed(base_seed, global_frame_index, chunk):
    """Seed of the generator that produces frames [k*chunk, (k+1)*chunk): a function of the global index only."""
    return int(base_seed) + 7919 * (int(global_frame_index) // int(chunk))


class Counters:
    """{frames, frame_errors, symbol_errors} + ProbResult histogram; the serial comparison loop of
    BinaryPolarEncoderDecoder.py:374-387 / QaryPolarEncoderDecoder.py:907-930 as integer accumulators."""
    SIZE = 9

    def __init__(self):
        self.v = np.zeros(self.SIZE, dtype=np.int64)

    def add(self, sent, decoded, prob_result=None):
        sent, decoded = np.asarray(sent), np.asarray(decoded)
        assert sent.shape == decoded.shape and sent.ndim == 2
        diff = sent != decoded
        self.v[0] += sent.shape[0]
        self.v[1] += int(diff.any(axis=1).sum())
        self.v[2] += int(diff.sum())
        if prob_result is not None:
            self.v[3:9] += np.bincount(np.asarray(prob_result, dtype=np.int64), minlength=6)[:6]
        return self

    def add_device(self, counters3):
        """Accumulate a device int64[3] tensor produced by engine.count_errors."""
        self.v[:3] += counters3.detach().cpu().numpy().astype(np.int64)
        return self

    def all_reduce(self, device=None):
        """Sum over all ranks of the default process group (no-op when torch.distributed is not initialised)."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self
        t = torch.from_numpy(self.v.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        self.v = t.cpu().numpy()
        return self

    @property
    def frames(self):
        return int(self.v[0])

    @property
    def fer(self):
        return self.v[1] / max(1, self.v[0])

    def ser(self, symbols_per_frame):
        return self.v[2] / max(1, self.v[0] * symbols_per_frame)

    def wilson_interval(self, z=1.96):
        """95 % Wilson interval of the frame-error rate (the reference prints a bare ratio; SURVEY.md 8d defines this CI)."""
        n, k = max(1, int(self.v[0])), int(self.v[1])
        p = k / n
        den = 1 + z * z / n
        mid = (p + z * z / (2 * n)) / den
        half = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
        return max(0.0, mid - half), min(1.0, mid + half)
