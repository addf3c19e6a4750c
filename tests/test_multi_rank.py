"""The N > 1 host logic on CPU with gloo, world_size 2: sharding by whole target reads, per-rank
stitching (src/polisher.cpp:520-546) and the variable-length gather of corrected reads to rank 0.  The engine call
itself needs a GPU, so each rank takes its shard's window results from the checker instead (this is a test of the
plumbing around the engine, not of the engine)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from vechat_b200._ffi import make_params
from vechat_b200.polisher import _gather_records, shard_targets, stitch, window_work
from vechat_b200.sim import Simulator


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _names(t):
    return "read%d" % t


def _worker(rank, world, port, out_path):
    import torch
    from oracle import checker
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        sim = Simulator("pb_clr_10k_x_10kb", n_reads=200, genome_len=60_000)
        batch = sim.windows(0, 6)
        p = make_params(haplotype=0)  # linear mode: the cheaper program, same plumbing
        shards = shard_targets(batch.win_target, window_work(batch), world)
        w0, w1 = shards[rank]
        sb = batch.slice(w0, w1)
        res = checker.oracle_polish(sb, p, threads=2)
        recs = stitch(res, batch.win_target, batch.win_rank, _names, batch.target_coverage, w0=w0)
        merged, nbytes = _gather_records(recs, rank, world, None, torch.device("cpu"))
        if rank == 0:
            whole = checker.oracle_polish(batch, p, threads=4)
            want = stitch(whole, batch.win_target, batch.win_rank, _names, batch.target_coverage)
            ok = merged == want and shards[0][0] == 0 and shards[-1][1] == batch.n_windows and all(
                shards[i][1] == shards[i + 1][0] for i in range(world - 1)) and all(
                int(batch.win_rank[s[0]]) == 0 for s in shards if s[0] < batch.n_windows)
            with open(out_path, "w") as f:
                f.write("OK %d" % len(merged) if ok else "MISMATCH")
        assert merged is None or rank == 0
    finally:
        dist.destroy_process_group()


def test_shard_stitch_gather_world2(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    txt = open(out).read()
    assert txt.startswith("OK") and int(txt.split()[1]) == 6, txt


def test_shard_targets_properties():
    rng = np.random.default_rng(5)
    # 40 targets with 1..30 windows each
    counts = rng.integers(1, 30, size=40)
    win_target = np.repeat(np.arange(40), counts)
    work = rng.integers(1, 1000, size=len(win_target)).astype(np.float64)
    for world in (1, 2, 3, 8, 64):
        sh = shard_targets(win_target, work, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(win_target)
        for a, b in zip(sh[:-1], sh[1:]):
            assert a[1] == b[0]
        for w0, w1 in sh:
            assert w0 <= w1
            if w0 < len(win_target) and w0 > 0:
                assert win_target[w0] != win_target[w0 - 1]  # cut only between targets
    assert shard_targets(np.zeros(0, dtype=np.int64), np.zeros(0), 4) == [(0, 0)] * 4


def _stitch_literal(result, win_target, win_rank, names, coverages, fragment_correction, drop_unpolished):
    """src/polisher.cpp:520-546 transcribed statement by statement (the check for polisher.stitch)."""
    out, data, num_polished = [], b"", 0
    n = len(result.polished)
    for i in range(n):
        num_polished += 1 if result.polished[i] else 0
        data += result.window(i)
        if i == n - 1 or win_rank[i + 1] == 0:
            ratio = num_polished / float(win_rank[i] + 1)
            if (not drop_unpolished) or ratio > 0:
                tags = "r" if fragment_correction else ""
                tags += " LN:i:" + str(len(data))
                tags += " RC:i:" + str(coverages[int(win_target[i])])
                tags += " XC:f:" + "%f" % ratio
                out.append((names[int(win_target[i])] + tags, data))
            num_polished, data = 0, b""
    return out


def test_stitch_matches_the_reference_loop():
    from vechat_b200._ffi import PolishResult
    rng = np.random.default_rng(9)
    for trial in range(20):
        counts = rng.integers(1, 9, size=int(rng.integers(1, 30)))
        win_target = np.repeat(np.arange(len(counts)), counts)
        win_rank = np.concatenate([np.arange(c) for c in counts])
        lens = rng.integers(0, 40, size=len(win_target))
        cons = rng.integers(65, 70, size=int(lens.sum())).astype(np.uint8)
        cons_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        polished = (rng.random(len(win_target)) < (0.0 if trial % 5 == 0 else 0.6)).astype(np.uint8)
        res = PolishResult(cons, cons_off, polished)
        names = ["t%d" % t for t in range(len(counts))]
        cov = [int(x) for x in rng.integers(0, 50, size=len(counts))]
        for frag in (True, False):
            for drop in (True, False):
                got = stitch(res, win_target, win_rank, names, cov, fragment_correction=frag, drop_unpolished=drop)
                assert got == _stitch_literal(res, win_target, win_rank, names, cov, frag, drop)
