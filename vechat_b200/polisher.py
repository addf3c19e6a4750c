"""Host-side mirror of racon::Polisher::polish (reference src/polisher.cpp:491-562) over the C-ABI engine.

The reference submits one task per window to a thread pool, waits for them in window order and stitches the
window consensuses of each target read into one Sequence whose name carries the LN/RC/XC tags.  Here:

  * windows -> one vgc_polish call per GPU (vechat_b200.engine.Engine): all windows of a shard in one batch;
  * multi-GPU: one process per GPU; whole targets are assigned to ranks in contiguous ranges balanced by
    estimated work, so every read is stitched where it was corrected; the only collective is the
    gather of the corrected reads to rank 0 (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests);
  * stitching and header tags byte-for-byte as src/polisher.cpp:525-546 (std::to_string(double) == "%f").

Nothing here computes a consensus: without the CUDA engine polish() raises.
"""
import numpy as np

from .engine import Engine


def window_work(batch):
    """Per-window work estimate: (sum of layer lengths) x backbone length ~ DP cells of the build phase."""
    lens = np.diff(batch.seq_off.astype(np.int64))
    csum = np.concatenate([[0], np.cumsum(lens)])
    wf = batch.win_first.astype(np.int64)
    tot = csum[wf[1:]] - csum[wf[:-1]]
    blen = lens[wf[:-1]] if len(lens) else np.zeros(0, dtype=np.int64)
    return tot * blen


def shard_targets(win_target, work, world):
    """Contiguous window ranges [w0, w1) per rank, cut only at target boundaries (Window::rank() == 0,
    src/polisher.cpp:530), balanced by `work`.  Returns a list of `world` (w0, w1) pairs covering all windows."""
    nw = len(win_target)
    if nw == 0:
        return [(0, 0)] * world
    starts = np.flatnonzero(np.concatenate([[True], win_target[1:] != win_target[:-1]]))
    cum = np.concatenate([[0], np.cumsum(np.asarray(work, dtype=np.float64))])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        goal = total * r / world
        # target boundary whose cumulative work is closest to the goal, not before the previous cut
        i = int(np.argmin(np.abs(cum[starts] - goal)))
        cuts.append(max(int(starts[i]), cuts[-1]))
    cuts.append(nw)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def stitch(result, win_target, win_rank, names, coverages, fragment_correction=True, drop_unpolished=False,
           w0=0):
    """src/polisher.cpp:520-546 for windows [w0, w0 + n) of a shard whose results are in `result` (local index).
    Returns [(header, sequence bytes)].  The windows of a target are consecutive and so are their corrected bytes
    in `result.cons`, so a target's sequence is one slice (no per-window concatenation)."""
    n = len(result.polished)
    if n == 0:
        return []
    rank = np.asarray(win_rank[w0:w0 + n])
    # a target ends at the last window or where the next window has rank 0 (polisher.cpp:530)
    last = np.flatnonzero(np.concatenate([rank[1:] == 0, [True]]))
    first = np.concatenate([[0], last[:-1] + 1])
    csum = np.concatenate([[0], np.cumsum(np.asarray(result.polished, dtype=np.int64))])
    off = np.asarray(result.cons_off, dtype=np.int64)
    cons = result.cons
    tags0 = "r" if fragment_correction else ""
    get_cov = coverages.get if hasattr(coverages, "get") else None
    out = []
    for f, l in zip(first.tolist(), last.tolist()):
        num_polished = int(csum[l + 1] - csum[f])
        ratio = num_polished / float(int(rank[l]) + 1)
        if drop_unpolished and not ratio > 0:
            continue
        seq = cons[off[f]:off[l + 1]].tobytes()
        t = int(win_target[w0 + l])
        cov = get_cov(t, 0) if get_cov else coverages[t]
        tags = "%s LN:i:%d RC:i:%d XC:f:%f" % (tags0, len(seq), int(cov), ratio)
        out.append(((names(t) if callable(names) else names[t]) + tags, seq))
    return out


def _gather_records(records, rank, world, group, device):
    """Variable-length gather of [(header, bytes)] to rank 0: all_gather of byte counts, then one gather of the
    packed payloads (padded to the maximum).  NCCL on GPUs, gloo on CPU."""
    import torch
    import torch.distributed as dist
    heads = "\n".join(h for h, _ in records).encode()
    lens = np.array([len(s) for _, s in records], dtype=np.int64)
    payload = b"".join([np.int64(len(records)).tobytes(), np.int64(len(heads)).tobytes(), lens.tobytes(), heads]
                       + [s for _, s in records])
    n = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    mx = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    # gather as all_gather into rank-0-sized buffers only where needed: dist.gather is supported by both backends
    outs = [torch.zeros(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None, int(len(payload))
    merged = []
    for r in range(world):
        raw = outs[r][:int(sizes[r].item())].cpu().numpy().tobytes()
        k = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
        hl = int(np.frombuffer(raw[8:16], dtype=np.int64)[0])
        ls = np.frombuffer(raw[16:16 + 8 * k], dtype=np.int64)
        o = 16 + 8 * k
        hs = raw[o:o + hl].decode().split("\n") if k else []
        o += hl
        for i in range(k):
            merged.append((hs[i], raw[o:o + int(ls[i])]))
            o += int(ls[i])
    return merged, int(len(payload))


class Polisher:
    """One per process/GPU.  polish() == Polisher::polish for a window set already tiled by initialize()."""

    def __init__(self, device=0, rank=0, world=1, group=None, engine=None, **params):
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self.engine = engine if engine is not None else Engine(device, **params)
        self.last_stats = None

    def shard(self, batch):
        return shard_targets(batch.win_target, window_work(batch), self.world)[self.rank]

    def polish_shard(self, shard_batch):
        """The engine call for this rank's windows (host buffers in, host buffers out)."""
        result, stats = self.engine.polish(shard_batch)
        self.last_stats = stats
        return result

    def submit_shard(self, shard_batch):
        """First half of polish_shard for callers with several batches (vgc_submit): host prepare, packing of the
        bases and H2D of `shard_batch` start on the engine's worker thread.  Submit batch i + 1 before collecting
        batch i and its staging overlaps the kernels of batch i."""
        self.engine.submit(shard_batch)

    def collect_shard(self):
        """Second half (vgc_collect): kernels + D2H of the oldest submitted batch."""
        result, stats = self.engine.collect()
        self.last_stats = stats
        return result

    def polish(self, batch, names, coverages=None, fragment_correction=True, drop_unpolished=False,
               shard_batch=None, gather_device=None):
        """batch: WindowBatch of ALL windows with .win_target/.win_rank (every rank holds the tiling, as every
        reference process would after initialize()); returns [(header, sequence)] in target order on rank 0, None
        on other ranks.  `shard_batch` may carry this rank's pre-sliced windows (bench: avoids re-slicing)."""
        w0, w1 = self.shard(batch)
        if shard_batch is None:
            shard_batch = batch if (w0, w1) == (0, batch.n_windows) else batch.slice(w0, w1)
        result = self.polish_shard(shard_batch)
        cov = coverages if coverages is not None else getattr(batch, "target_coverage", {})
        records = stitch(result, batch.win_target, batch.win_rank, names, cov, fragment_correction, drop_unpolished,
                         w0=w0)
        if self.world == 1:
            return records
        import torch
        dev = gather_device if gather_device is not None else (
            torch.device("cuda", self.device) if torch.cuda.is_available() else torch.device("cpu"))
        merged, _ = _gather_records(records, self.rank, self.world, self.group, dev)
        return merged
