"""-m gpu: the CUDA engine, called through the C-ABI (include/vgc.h via ctypes), against
  * the committed outputs of the unmodified reference (tests/golden/windows_*.npz),
  * the oracle (and the compiled reference when its .so travelled) on seeded windows incl. the edge cases the
    reference's window code distinguishes (empty batch, < 3 sequences, dropped layers, FASTA dummy quality, N
    bases, missing qualities, partial spans, long layers, deep windows with > 16 equal sort keys),
  * size-independent properties at the bench workload's size (determinism, resident == host-buffer path,
    polished flags, per-window independence: a window's result does not depend on its batch neighbours).
Bit-exact: this is byte/integer work."""
import os

import numpy as np
import pytest
import torch

from conftest import assert_same, golden_names, load_golden
from oracle import checker
from vechat_b200._ffi import WindowBatch, make_params, VGC_WIN_TGS
from vechat_b200.engine import Engine, VgcError
from vechat_b200.sim import Simulator, fuzz_batch, fuzz_window

pytestmark = pytest.mark.gpu

_engines = {}


def eng(**pkw):
    key = tuple(sorted(pkw.items()))
    if key not in _engines:
        if len(_engines) >= 6:  # an engine keeps its scratch (GBs after a wide-path test): do not hoard them
            for e in _engines.values():
                e.close()
            _engines.clear()
        _engines[key] = Engine(0, **pkw)
    return _engines[key]


def want_of(batch, pkw):
    """The authority: the compiled unmodified reference when its .so travelled, else the oracle port."""
    p = make_params(**pkw)
    return checker.ref_polish(batch, p, threads=8) if checker.have_ref() else checker.oracle_polish(batch, p, threads=8)


def check(batch, pkw, label):
    got, st = eng(**pkw).polish(batch)
    assert_same(got, want_of(batch, pkw), label)
    return st


def test_gpu_present_and_native_library_loaded():
    assert torch.cuda.is_available()
    e = eng()
    assert e.lib.vgc_version().startswith(b"vechat_b200")


@pytest.mark.parametrize("name", golden_names())
def test_golden(name):
    batch, pkw, want = load_golden(name)
    got, st = eng(**pkw).polish(batch)
    assert_same(got, want, name)
    assert st["kernel_launches"] >= 1 and st["cells"] > 0


@pytest.mark.parametrize("seed,kw,pkw", [
    (401, dict(n_windows=32), dict()),
    (402, dict(n_windows=32, partial=0.8), dict()),
    (403, dict(n_windows=32, fastq=False), dict()),
    (404, dict(n_windows=32, n_frac=0.05, null_qual=0.5), dict()),
    (405, dict(n_windows=32), dict(haplotype=0)),
    (406, dict(n_windows=32, partial=0.8), dict(haplotype=0, trim=0)),
    (407, dict(n_windows=8, depth=70, length=100), dict()),
    (408, dict(n_windows=16, err=0.4), dict(num_prune=4)),
    (409, dict(n_windows=16), dict(num_prune=1)),
    (410, dict(n_windows=16), dict(min_confidence=1.0, min_support=1.0)),
    (411, dict(n_windows=16), dict(min_confidence=0.0, min_support=0.0)),
    (412, dict(n_windows=16), dict(match=5, mismatch=-4, gap=-8)),
    (413, dict(n_windows=8, length=560, depth=12), dict()),           # rows up to ~640 columns
    (414, dict(n_windows=4, length=800, depth=8), dict()),            # wide-row template (K = 16)
    (415, dict(n_windows=4, length=800, depth=8, partial=0.6), dict(haplotype=0)),
    (416, dict(n_windows=3, length=120, depth=200), dict()),          # fuzz to depth 200 ...
    (417, dict(n_windows=3, length=700, depth=40, partial=0.3), dict()),   # ... and length 700
    (418, dict(n_windows=2, length=700, depth=200), dict(haplotype=0)),
])
def test_fuzz(seed, kw, pkw):
    check(fuzz_batch(seed, **kw), pkw, "seed %d" % seed)


# ---- no capacity cliffs: the reference takes any layer length, any byte, any int8 scores ----------
@pytest.mark.parametrize("length,depth,pkw", [
    (1000, 10, dict()),                 # racon -w 1000: layers of ~1000-1150 bases, beyond the int16 kernels' widest row
    (2000, 8, dict()),                  # racon -w 2000
    (2000, 6, dict(haplotype=0)),
    (5000, 4, dict()),
])
def test_long_windows_run_on_the_wide_path(length, depth, pkw):
    b = fuzz_batch(600 + length, n_windows=3, length=length, depth=depth, partial=0.3)
    assert max(np.diff(b.seq_off)) > 1024
    st = check(b, pkw, "-w %d" % length)
    assert st["relaunched_windows"] >= 0


@pytest.mark.parametrize("pkw", [dict(), dict(haplotype=0)])
def test_iupac_reads_more_than_eight_codes(pkw):
    b = fuzz_batch(610, n_windows=24, length=180, depth=14, iupac_frac=0.08)
    assert len(set(b.bases.tobytes())) >= 12
    check(b, pkw, "IUPAC")


def test_bytes_above_127_are_invalid():
    """The reference indexes spoa's coder table with a signed char (it crashes on such input): rejected up front."""
    b = fuzz_batch(611, n_windows=6, length=100, depth=8)
    bases = b.bases.copy()
    bases[bases == ord("T")] = 200
    b2 = WindowBatch(bases, b.quals, b.seq_off, b.has_qual, b.begin, b.end, b.win_first, b.win_flags)
    with pytest.raises(VgcError) as e:
        eng().polish(b2)
    assert e.value.code == 1 and ">= 128" in str(e.value)


def test_more_than_sixteen_codes_is_a_capacity_error():
    b = fuzz_batch(612, n_windows=2, length=100, depth=4)
    bases = b.bases.copy()
    bases[:40] = np.arange(40, dtype=np.uint8) + 40
    b2 = WindowBatch(bases, b.quals, b.seq_off, b.has_qual, b.begin, b.end, b.win_first, b.win_flags)
    with pytest.raises(VgcError) as e:
        eng().polish(b2)
    assert "16 distinct" in str(e.value)


@pytest.mark.parametrize("pkw,depth,length", [
    (dict(match=5, mismatch=-10, gap=-16), 60, 300),    # VERDICT: -m 5 -x -10 -g -16 leaves int16 on ~1900 rows
    (dict(match=5, mismatch=-4, gap=-8), 200, 400),     # scripts/vechat defaults on a deep window
    (dict(match=127, mismatch=-128, gap=-128), 10, 150),  # int8 extremes: wide from the first alignment
    (dict(match=5, mismatch=-10, gap=-16, haplotype=0), 60, 300),
])
def test_scores_beyond_int16_switch_to_int32(pkw, depth, length):
    b = fuzz_batch(620 + depth, n_windows=2, length=length, depth=depth)
    check(b, pkw, "scores %r depth %d" % (pkw, depth))


def test_positive_gap_is_rejected_like_the_reference():
    """spoa::AlignmentEngine::Create throws on g > 0 (vendor/spoa/src/alignment_engine.cpp:44-48)."""
    with pytest.raises(VgcError) as e:
        Engine(0, gap=1)
    assert "gap opening penalty must be non-positive" in str(e.value)


def test_limits_are_reported():
    lim = eng().limits()
    assert lim["max_layer_len"] == 16383 and lim["max_codes"] == 16 and lim["fast_layer_len"] == 1024


@pytest.mark.parametrize("pkw", [dict(haplotype=0, trim=1), dict()])
def test_ngs_windows(pkw):
    """WindowType::kNGS windows (mean read length <= 1000, src/polisher.cpp:289-290) are never trimmed
    (src/window.cpp:141)."""
    b = fuzz_batch(77, n_windows=24, partial=0.7)
    b.win_flags[1::2] &= np.uint8(~VGC_WIN_TGS & 0xFF)
    check(b, pkw, "NGS windows")


def test_empty_batch():
    b = WindowBatch.from_windows([])
    r, st = eng().polish(b)
    assert r.total_bases() == 0 and len(r.polished) == 0


def test_windows_below_three_sequences_return_backbone():
    rng = np.random.default_rng(5)
    wins = [fuzz_window(rng, depth=d) for d in (0, 1, 0, 1)]
    b = WindowBatch.from_windows(wins)
    r, _ = eng().polish(b)
    for w in range(4):
        assert r.window(w) == wins[w][0][0][0] and r.polished[w] == 0


def test_dropped_layers_do_not_count():
    """add_layer drops empty layers and begin == end silently (src/window.cpp:51-54): a window whose third
    sequence is such a layer has < 3 sequences."""
    rng = np.random.default_rng(6)
    layers, flags = fuzz_window(rng, depth=2, partial=0.0)
    layers[2] = (layers[2][0], layers[2][1], 7, 7)
    b = WindowBatch.from_windows([(layers, flags)])
    check(b, dict(), "dropped layer")
    r, _ = eng().polish(b)
    assert r.polished[0] == 0


def test_invalid_layer_positions_are_rejected():
    rng = np.random.default_rng(7)
    layers, flags = fuzz_window(rng, depth=3, partial=0.0)
    layers[1] = (layers[1][0], layers[1][1], 10, 5)    # begin > end: the reference exits (src/window.cpp:62-67)
    with pytest.raises(VgcError) as ei:
        eng().polish(WindowBatch.from_windows([(layers, flags)]))
    assert ei.value.code == 1


def test_mixed_batch_with_tiny_and_large_windows():
    rng = np.random.default_rng(8)
    wins = []
    for i in range(40):
        wins.append(fuzz_window(rng, length=int(rng.integers(4, 500)), depth=int(rng.integers(0, 30)),
                                partial=float(rng.random())))
    check(WindowBatch.from_windows(wins), dict(), "mixed")
    check(WindowBatch.from_windows(wins), dict(haplotype=0), "mixed linear")


def test_scratch_regrow_path(monkeypatch):
    """A tiny memory budget forces small slots -> node overflow -> relaunch with larger slots."""
    monkeypatch.setenv("VGC_MEM_BUDGET_MB", "2048")
    e = Engine(0)
    b = fuzz_batch(420, n_windows=6, depth=45, length=400)
    got, st = e.polish(b)
    p = make_params()
    assert_same(got, want_of(b, {}), "regrow")
    e.close()


def test_retry_pass_with_exact_capacities(monkeypatch):
    """First-pass slots far too small for the graphs (node estimate = backbone + 1/200 of the layer bases): the
    windows overflow, report it, and are re-run in a second pass sized by the exact upper bound."""
    monkeypatch.setenv("VGC_NODE_SHARE_DIV", "200")
    e = Engine(0)
    b = fuzz_batch(421, n_windows=12, depth=30, length=300, err=0.3)
    got, st = e.polish(b)
    assert st["relaunched_windows"] > 0
    assert_same(got, want_of(b, {}), "retry pass")
    e.close()


@pytest.mark.parametrize("env", [dict(VGC_SORT_SMEM="3072"), dict(VGC_SORT_GROWTH="0.0005")])
def test_sort_out_of_hbm_when_shared_memory_is_short(monkeypatch, env):
    """The staged (shared-memory) TopologicalSort / LargestSubgraph fall back to their HBM twins when the graph does
    not fit the kernel's shared memory: same results, slower."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for pkw, seed, kw in ((dict(), 422, dict(n_windows=10, depth=20, length=250, partial=0.4)),
                          (dict(haplotype=0), 423, dict(n_windows=10, depth=20, length=250, partial=0.4))):
        e = Engine(0, **pkw)
        b = fuzz_batch(seed, **kw)
        got, _ = e.polish(b)
        assert_same(got, want_of(b, pkw), "hbm sort %s" % (env,))
        e.close()


@pytest.fixture(scope="module")
def pb_batch():
    sim = Simulator("pb_clr_10k_x_10kb", n_reads=1500, genome_len=500_000)
    return sim.windows(0, 60)     # 60 targets x 20 windows, depth ~30


def test_sim_pb_sample_vs_checker(pb_batch):
    sub = pb_batch.select(range(0, pb_batch.n_windows, 10))
    check(sub, dict(), "sim pb haplotype")
    check(sub, dict(haplotype=0), "sim pb linear")


def test_properties_at_scale(pb_batch):
    e = eng()
    r1, st1 = e.polish(pb_batch)
    r2, st2 = e.polish(pb_batch)
    assert r1.cons.tobytes() == r2.cons.tobytes() and st1["cells"] == st2["cells"]        # deterministic
    e.upload(pb_batch)
    r3, st3 = e.polish_resident()
    assert r3.cons.tobytes() == r1.cons.tobytes()                                        # resident == host path
    # independence: a permuted batch gives the permuted result
    perm = np.random.default_rng(1).permutation(pb_batch.n_windows)[:200]
    rp, _ = e.polish(pb_batch.select(perm))
    for i, w in enumerate(perm):
        assert rp.window(i) == r1.window(int(w))
    # every window with >= 3 sequences is polished; corrected length stays near the backbone's
    nseq = np.diff(pb_batch.win_first)
    assert all(int(r1.polished[w]) == int(nseq[w] >= 3) for w in range(pb_batch.n_windows))
    blen = np.array([int(pb_batch.seq_off[f + 1] - pb_batch.seq_off[f]) for f in pb_batch.win_first[:-1]])
    clen = np.diff(r1.cons_off.astype(np.int64))[:pb_batch.n_windows]
    assert np.all(clen[nseq >= 3] > 0.5 * blen[nseq >= 3]) and np.all(clen < 2.0 * blen + 64)
    assert st1["alignments"] == sum(3 * (int(n) - 1) + 3 for n in nseq if n >= 3)    # 3*depth+3


@pytest.mark.parametrize("config,kw,n_targets,step", [
    ("ont_10k_x_20kb", dict(n_reads=400, genome_len=260_000), 2, 4),      # BASELINE configs[2]: ONT, 20 kb, 10 % error
    ("hap2_50k_x_12kb", dict(n_reads=900, genome_len=180_000), 3, 5),     # configs[3]: 2 haplotypes 50:50, depth ~60
])
def test_other_baseline_configs_sample(config, kw, n_targets, step):
    """Scaled-down instances of BASELINE.json's other workloads (same error model, read length, windowing; smaller
    genome so the depth matches) against the reference, every `step`-th window."""
    sim = Simulator(config, **kw)
    b = sim.windows(0, n_targets)
    sub = b.select(range(0, b.n_windows, step))
    st = check(sub, dict(), config)
    assert st["relaunched_windows"] == 0 or st["relaunched_windows"] < sub.n_windows


def test_large_batch_many_groups():
    """A batch wide enough to be dealt into many stream groups (one per number of alignments) and to exercise the
    lockstep schedule's prefix logic with windows of very different depth."""
    rng = np.random.default_rng(77)
    wins = [fuzz_window(rng, length=int(rng.integers(30, 120)), depth=int(rng.integers(0, 40))) for _ in range(1500)]
    check(WindowBatch.from_windows(wins), dict(), "many groups")
    check(WindowBatch.from_windows(wins[:700]), dict(haplotype=0, trim=0), "many groups linear")
