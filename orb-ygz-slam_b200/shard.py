"""Multi-GPU plumbing of the front end: the path shards over independent frames / streams, one
process per GPU, no data-path collective.  The only exchange is ONE all-gather of per-rank throughput counters
(NCCL on the GPU box; gloo in the CPU tests); the device-timed durations ride in the same row (as nanoseconds) and
the max over ranks is taken from the gathered table, so there is exactly one collective per run."""
import numpy as np

COUNTER_COLUMNS = ["frames", "keypoints", "matches", "sia_iters", "bytes_in", "bytes_out", "kernel_launches", "elapsed_ns"]


def shard_range(n_units, rank, world):
    """Block partition of n_units independent units (frames, stereo pairs, streams): [begin, end) of `rank`."""
    assert 0 <= rank < world
    return (n_units * rank) // world, (n_units * (rank + 1)) // world


def stream_of_rank(n_streams, rank, world):
    """C4: stream i -> GPU i (round-robin when there are more streams than ranks)."""
    return [s for s in range(n_streams) if s % world == rank]


def gather_counters(dist, counters, device=None, timings_s=()):
    """The ONE collective: all-gather of every rank's row = the 8 u64 counters (+ the rank's timings in ns) -> int64 array
    [world, 8 + len(timings_s)].  dist=None: single rank."""
    row = np.array([int(counters.get(k, 0)) for k in COUNTER_COLUMNS] + [int(round(float(t) * 1e9)) for t in timings_s], np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return row[None]
    import torch
    mine = torch.from_numpy(row.copy())
    if device is not None:
        mine = mine.to(device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return torch.stack(out).cpu().numpy()


def timings_max(gathered, n_timings):
    """Slowest rank's value of each timing column appended by gather_counters (seconds)."""
    return [float(gathered[:, len(COUNTER_COLUMNS) + i].max()) / 1e9 for i in range(n_timings)]


def max_over_ranks(dist, values, device=None):
    """Element-wise max of a list of floats over ranks as its own all-reduce (kept for callers outside bench.py, which
    folds its timings into gather_counters)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def job_throughput(gathered, seconds):
    """Whole-job frames/s: all ranks' frames over the slowest rank's time."""
    return float(gathered[:, 0].sum()) / seconds
