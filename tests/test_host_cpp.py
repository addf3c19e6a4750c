"""The C++ host side above the C-ABI (vechat_b200/csrc/host/vgc_host.hpp): createWindow / Window::add_layer /
B200Polisher::polish with the reference's names and error behaviour (src/window.hpp:27-77, src/polisher.cpp:491-562),
driven through the flat C test entry of libvgchost.so."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from vechat_b200 import build
from vechat_b200._ffi import VgcBatch, VgcParams, WindowBatch, make_params
from vechat_b200.sim import Simulator, fuzz_batch

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build.build_host())
        _lib.vgch_pack_roundtrip.restype = C.c_long
        _lib.vgch_pack_roundtrip.argtypes = [C.POINTER(VgcBatch), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        _lib.vgch_polish_fasta.restype = C.c_long
        _lib.vgch_polish_fasta.argtypes = [C.POINTER(VgcBatch), C.POINTER(VgcParams), C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                           C.c_char_p, C.c_uint64]
        _lib.vgch_last_error.restype = C.c_char_p
    return _lib


def fnv(h, arr):
    for x in np.asarray(arr).tolist():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def expected_pack_hash(batch):
    """What BatchPacker must produce: the layers add_layer keeps (window.cpp:51-54), in order."""
    keep = []
    for w in range(batch.n_windows):
        f, l = int(batch.win_first[w]), int(batch.win_first[w + 1])
        keep.append(f)
        for i in range(f + 1, l):
            if int(batch.seq_off[i + 1]) == int(batch.seq_off[i]) or batch.begin[i] == batch.end[i]:
                continue
            keep.append(i)
    bases = np.concatenate([batch.bases[int(batch.seq_off[i]):int(batch.seq_off[i + 1])] for i in keep])
    h = 1469598103934665603
    for arr in (bases, batch.begin[keep], batch.end[keep], batch.win_flags, batch.has_qual[keep]):
        h = fnv(h, arr)
    return len(keep), h


def ids_of(batch):
    wid = np.arange(batch.n_windows, dtype=np.uint64)
    wrank = np.zeros(batch.n_windows, dtype=np.uint32)
    return wid, wrank


@pytest.mark.parametrize("seed,kw", [(401, dict(n_windows=10)), (402, dict(n_windows=10, fastq=False)),
                                     (403, dict(n_windows=10, n_frac=0.05, null_qual=0.5, partial=0.6)),
                                     (405, dict(n_windows=300, length=20, depth=3, null_qual=0.3))])  # threaded copy
def test_pack_matches_window_contents(seed, kw):
    batch = fuzz_batch(seed, **kw)
    wid, wrank = ids_of(batch)
    b = batch.c_struct()
    h = C.c_uint64()
    n = lib().vgch_pack_roundtrip(C.byref(b), wid.ctypes.data, wrank.ctypes.data, C.byref(h))
    want_n, want_h = expected_pack_hash(batch)
    assert n == want_n, lib().vgch_last_error()
    assert h.value == want_h


def test_add_layer_error_is_the_references():
    layers = [(b"ACGTACGTAC", b"5555555555", 0, 0), (b"ACGT", b"5555", 8, 4)]  # begin >= end
    batch = WindowBatch.from_windows([(layers, 1)])
    wid, wrank = ids_of(batch)
    b = batch.c_struct()
    h = C.c_uint64()
    assert lib().vgch_pack_roundtrip(C.byref(b), wid.ctypes.data, wrank.ctypes.data, C.byref(h)) == -1
    assert lib().vgch_last_error().decode() == "[racon::Window::add_layer] error: layer begin and end positions are invalid!"


def test_polish_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    batch = fuzz_batch(404, n_windows=2)
    wid, wrank = ids_of(batch)
    b = batch.c_struct()
    p = make_params()
    names = (C.c_char_p * batch.n_windows)(*[b"t%d" % i for i in range(batch.n_windows)])
    cov = np.zeros(batch.n_windows, dtype=np.uint32)
    out = C.create_string_buffer(1 << 16)
    rc = lib().vgch_polish_fasta(C.byref(b), C.byref(p), wid.ctypes.data, wrank.ctypes.data, names, cov.ctypes.data,
                                 batch.n_windows, 1, 0, 0, out, len(out))
    assert rc == -1 and b"no usable CUDA device" in lib().vgch_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("pkw", [dict(), dict(haplotype=0)])
def test_cpp_polisher_fasta_equals_reference_stitch(pkw):
    """createWindow/add_layer/B200Polisher::polish (C++) -> FASTA == the reference's consensus per window stitched
    by Polisher::polish's rules (src/polisher.cpp:520-546)."""
    from oracle import checker
    from vechat_b200.polisher import stitch
    sim = Simulator("pb_clr_10k_x_10kb", n_reads=300, genome_len=100_000)
    batch = sim.windows(0, 4)
    p = make_params(**pkw)
    want = checker.ref_polish(batch, p, threads=8) if checker.have_ref() else checker.oracle_polish(batch, p, threads=8)
    nt = int(batch.win_target.max()) + 1
    recs = stitch(want, batch.win_target, batch.win_rank, lambda t: "read%d" % t, batch.target_coverage)
    text = b"".join(b">" + h.encode() + b"\n" + s + b"\n" for h, s in recs)
    names = (C.c_char_p * nt)(*[b"read%d" % t for t in range(nt)])
    cov = np.array([int(batch.target_coverage.get(t, 0)) for t in range(nt)], dtype=np.uint32)
    wid = np.asarray(batch.win_target, dtype=np.uint64)
    wrank = np.asarray(batch.win_rank, dtype=np.uint32)
    b = batch.c_struct()
    out = C.create_string_buffer(len(text) * 2 + 4096)
    rc = lib().vgch_polish_fasta(C.byref(b), C.byref(p), wid.ctypes.data, wrank.ctypes.data, names, cov.ctypes.data,
                                 nt, 1, 0, 0, out, len(out))
    assert rc >= 0, lib().vgch_last_error()
    assert out.raw[:rc] == text
