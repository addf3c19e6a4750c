"""The engine's algorithm template (vechat_b200/csrc/poa_core.h + host_prep.h) instantiated on the host with a
one-lane executor (tests/host_model/host_model.cpp) and compared with the oracle / committed reference outputs.
This checks, without a GPU, every serial piece that runs inside the kernel: AddAlignment, TopologicalSort,
Subgraph, PruneGraph, LargestSubgraph, AddWeights, traceback, heaviest bundle, the host-side rank sort and
average-weight arithmetic.  (The CUDA fill itself is covered by the -m gpu tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_same, golden_names, load_golden
from oracle import checker
from vechat_b200._ffi import VgcBatch, VgcParams, VgcResult, alloc_result, finish_result, make_params
from vechat_b200.sim import fuzz_batch

SRC = os.path.join(ROOT, "tests", "host_model", "host_model.cpp")
OUT = os.path.join(ROOT, "tests", "host_model", "_build", "libhostmodel.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(ROOT, "vechat_b200", "csrc")
        deps = [SRC, os.path.join(csrc, "poa_core.h"), os.path.join(csrc, "host_prep.h")]
        if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                            "-o", OUT, SRC, "-lpthread"], check=True)
        _lib = C.CDLL(OUT)
        _lib.hm_polish.restype = C.c_int
        _lib.hm_polish.argtypes = [C.POINTER(VgcBatch), C.POINTER(VgcParams), C.POINTER(VgcResult), C.c_int, C.c_int,
                                   C.POINTER(C.c_uint32)]
        _lib.hm_order_checks.argtypes = [C.POINTER(C.c_ulonglong)]
    return _lib


def hm_polish(batch, params, flags=0, k=10):
    b = batch.c_struct()
    r, arrays = alloc_result(batch)
    status = (C.c_uint32 * max(batch.n_windows, 1))()
    rc = lib().hm_polish(C.byref(b), C.byref(params), C.byref(r), flags, k, status)
    assert rc == 0, "hm_polish rc=%d" % rc
    assert all(s == 0 for s in status), "window status %s" % list(status)
    return finish_result(batch, arrays)


@pytest.mark.parametrize("name", golden_names())
def test_host_model_vs_committed_reference_outputs(name):
    batch, pkw, want = load_golden(name)
    got = hm_polish(batch, make_params(**pkw))
    assert_same(got, want, name)


@pytest.mark.parametrize("flags,k", [(1, 10), (8 << 8, 10), (0, 16), (2, 10), (4, 10)])
def test_host_model_slow_paths(flags, k):
    """flags bit0: sort without the staged 16-bit CSR; flags>>8: tiny DFS stack -> overflow redo; k=16: wide rows;
    bit1: the incremental order's dirty blocks do not fit their storage -> full sort; bit2: every alignment on the
    wide (int32, any width) path."""
    for seed, kw, pkw in [(301, dict(n_windows=8), dict()), (302, dict(n_windows=8, partial=0.7), dict(haplotype=0))]:
        batch = fuzz_batch(seed, **kw)
        p = make_params(**pkw)
        assert_same(hm_polish(batch, p, flags, k), checker.oracle_polish(batch, p, threads=4), "seed %d" % seed)


@pytest.mark.parametrize("seed", range(310, 316))
def test_host_model_fuzz_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    kw = dict(n_windows=10, partial=float(rng.random()), err=float(rng.uniform(0.02, 0.35)),
              fastq=bool(rng.integers(0, 2)), null_qual=float(rng.choice([0.0, 0.4])),
              n_frac=float(rng.choice([0.0, 0.03])))
    pkw = dict(haplotype=int(rng.integers(0, 2)), num_prune=int(rng.integers(1, 4)),
               min_confidence=float(rng.choice([0.0, 0.2, 0.5])), min_support=float(rng.choice([0.0, 0.2, 1.0])))
    batch = fuzz_batch(seed, **kw)
    p = make_params(**pkw)
    assert_same(hm_polish(batch, p), checker.oracle_polish(batch, p, threads=4), "seed %d %r %r" % (seed, kw, pkw))


@pytest.mark.parametrize("pkw", [dict(), dict(haplotype=0)])
def test_host_model_wide_path_and_alphabet(pkw):
    """Layers longer than the fast rows (K = 10: 640 columns) take the int32 wide path by themselves; IUPAC reads bring
    more than 8 distinct bytes (aligned lists of stride 16); big penalties leave the int16 range on a small graph."""
    p = make_params(**pkw)
    b = fuzz_batch(530, n_windows=3, length=760, depth=5)
    assert_same(hm_polish(b, p), checker.oracle_polish(b, p, threads=4), "long layers %r" % pkw)
    b = fuzz_batch(531, n_windows=6, length=150, depth=10, iupac_frac=0.08)
    assert len(set(b.bases.tobytes())) >= 12
    assert_same(hm_polish(b, p), checker.oracle_polish(b, p, threads=4), "iupac %r" % pkw)
    p2 = make_params(match=127, mismatch=-128, gap=-128, **pkw)
    b = fuzz_batch(532, n_windows=4, length=200, depth=8)
    assert_same(hm_polish(b, p2), checker.oracle_polish(b, p2, threads=4), "int8-extreme scores %r" % pkw)


def test_incremental_order_equals_full_sort():
    """The engine maintains the topological order incrementally after AddAlignment (poa_core.h order_update: only the
    DFS blocks an alignment touched are re-sorted).  The host model re-runs the reference's full DFS after every
    update and compares order, ranks, owners and block tables node for node."""
    out = (C.c_ulonglong * 4)()
    lib().hm_order_checks(out)
    before = list(out)
    for seed, kw, pkw in [(520, dict(n_windows=6, partial=0.5, depth=25, length=200), dict()),
                          (521, dict(n_windows=6, partial=0.9, err=0.3, depth=12, length=120), dict(haplotype=0)),
                          (522, dict(n_windows=4, depth=50, length=80), dict())]:
        batch = fuzz_batch(seed, **kw)
        p = make_params(**pkw)
        assert_same(hm_polish(batch, p), checker.oracle_polish(batch, p, threads=4), "seed %d" % seed)
    lib().hm_order_checks(out)
    same, fallback, different = (out[i] - before[i] for i in range(3))
    assert different == 0, "incremental order differs from the full sort in %d updates" % different
    assert same > 300 and same > 20 * fallback, (same, fallback)


def ngs_batch(seed, **kw):
    """fuzz batch whose odd windows are WindowType::kNGS (flag cleared): never trimmed (src/window.cpp:141)."""
    from vechat_b200._ffi import VGC_WIN_TGS
    b = fuzz_batch(seed, **kw)
    b.win_flags[1::2] &= np.uint8(~VGC_WIN_TGS & 0xFF)
    return b


@pytest.mark.parametrize("pkw", [dict(haplotype=0, trim=1), dict(haplotype=0, trim=0), dict()])
def test_ngs_windows(pkw):
    batch = ngs_batch(77, n_windows=24, partial=0.7)
    p = make_params(**pkw)
    want = checker.ref_polish(batch, p, threads=4) if checker.have_ref() else checker.oracle_polish(batch, p, threads=4)
    assert_same(checker.oracle_polish(batch, p, threads=4), want, "oracle, NGS windows %r" % pkw)
    assert_same(hm_polish(batch, p), want, "host model, NGS windows %r" % pkw)
    if pkw.get("haplotype", 1) == 0 and pkw.get("trim", 1) == 1:
        # the flag matters: with every window TGS some of these consensuses are trimmed
        tgs = checker.oracle_polish(fuzz_batch(77, n_windows=24, partial=0.7), p, threads=4)
        assert any(tgs.window(w) != want.window(w) for w in range(1, 24, 2))


def test_threaded_host_prep_matches_oracle():
    """>= 1024 windows: prepare_batch (rank sort, average weights, alphabet) runs on several host threads; the
    result must not depend on the split."""
    batch = fuzz_batch(311, n_windows=1100, length=24, depth=4, n_frac=0.02)
    p = make_params()
    got = hm_polish(batch, p)
    want = checker.oracle_polish(batch, p, threads=8)
    assert_same(got, want, "threaded prep")
