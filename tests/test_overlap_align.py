"""Overlap alignment -> CIGAR (include/vga.h, vechat_b200/csrc/ovl_align.cu + ovl_core.h).

Oracle: oracle/shims/edlib_standin.cpp — an independent implementation (own containers, clipped wavefront layout)
of the exact unit-cost global aligner both reference-program builds use in place of the absent edlib
(tests/test_example_binary.py checks it against the textbook DP).  Bar: identical CIGAR strings (byte for byte) and
edit distances.  Equality with the REAL edlib's tie-breaking is unpinned (edlib is not in this image).
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

_model = None


def model():
    global _model
    if _model is None:
        src = os.path.join(ROOT, "tests", "host_model", "ovl_model.cpp")
        out = os.path.join(ROOT, "tests", "host_model", "_build", "libovlmodel.so")
        core = os.path.join(ROOT, "vechat_b200", "csrc", "ovl_core.h")
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in (src, core)):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.dirname(core), "-o", out, src],
                           check=True)
        _model = C.CDLL(out)
        _model.ovl_model_align.restype = C.c_int
        _model.ovl_model_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_uint64,
                                           C.POINTER(C.c_uint64)]
    return _model


def model_cigar(q, t):
    buf = C.create_string_buffer(2 * (len(q) + len(t)) + 16)
    cells = C.c_uint64()
    d = model().ovl_model_align(q, len(q), t, len(t), buf, len(buf), C.byref(cells))
    assert d >= 0 and cells.value == (d + 1) ** 2
    return buf.value.decode(), d


def standin_cigar(q, t):
    from oracle import checker
    return checker.standin_cigar(q, t)


def noisy_pairs(seed, count, max_len, sub=0.04, ins=0.10, dele=0.08):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(rng.integers(1, max_len))
        t = rng.integers(0, 4, n)
        u = rng.random(n)
        q = []
        for x, r in zip(t.tolist(), u.tolist()):
            if r < dele:
                continue
            q.append(int(rng.integers(0, 4)) if r < dele + sub else x)
            while rng.random() < ins:
                q.append(int(rng.integers(0, 4)))
        out.append((bytes(b"ACGT"[x] for x in q), bytes(b"ACGT"[x] for x in t)))
    return out


EDGE = [(b"", b""), (b"", b"ACGT"), (b"ACGT", b""), (b"A", b"A"), (b"A", b"C"), (b"AAAA", b"TTTTTTT"),
        (b"ACGTACGT", b"ACGTACGT"), (b"ACGT" * 50, b"TGCA" * 40), (b"A" * 300, b"A" * 280), (b"AC" * 100, b"CA" * 100)]


def test_host_model_matches_standin_cigars():
    cases = EDGE + noisy_pairs(1, 150, 600) + noisy_pairs(2, 6, 6000) + noisy_pairs(3, 40, 300, 0.3, 0.3, 0.3)
    for q, t in cases:
        assert model_cigar(q, t) == standin_cigar(q, t), (q[:40], t[:40])


def test_host_model_matrix_edges_fuzz():
    """Short strings over two letters: most cells touch the matrix edge, where wf_cell leaves its fast path."""
    rng = np.random.default_rng(5)
    for _ in range(3000):
        q = bytes(rng.choice([65, 67], int(rng.integers(0, 11))).tolist())
        t = bytes(rng.choice([65, 67], int(rng.integers(0, 11))).tolist())
        assert model_cigar(q, t) == standin_cigar(q, t), (q, t)
    for q, t in [(b"ACGT" * 10, b"ACGT" * 3), (b"A" * 5, b"A" * 50), (b"ACGTTGCA", b"TT"), (b"T" * 40, b"ACGT" * 10 + b"T"),
                 (b"ACGT" * 30 + b"A", b"A")]:
        assert model_cigar(q, t) == standin_cigar(q, t)
        assert model_cigar(t, q) == standin_cigar(t, q)


def breaking_points_from_cigar(cigar, t_begin, t_end, q_start, window_length):
    """Checker: Overlap::find_breaking_points_from_cigar (src/overlap.cpp:226-292) restated base by base."""
    window_ends = [i - 1 for i in range(0, t_end, window_length) if i > t_begin] + [t_end - 1]
    out, w, found, first, last = [], 0, False, (0, 0), (0, 0)
    q_ptr, t_ptr = q_start - 1, t_begin - 1
    k = 0
    while k < len(cigar):
        j = k
        while cigar[j].isdigit():
            j += 1
        num, op = int(cigar[k:j]), cigar[j]
        k = j + 1
        if op == "M":
            for _ in range(num):
                q_ptr += 1
                t_ptr += 1
                if not found:
                    found, first = True, (t_ptr, q_ptr)
                last = (t_ptr + 1, q_ptr + 1)
                if t_ptr == window_ends[w]:
                    if found:
                        out += [first, last]
                    found = False
                    w += 1
        elif op == "I":
            q_ptr += num
        else:
            for _ in range(num):
                t_ptr += 1
                if t_ptr == window_ends[w]:
                    if found:
                        out += [first, last]
                    found = False
                    w += 1
    return out


def model_break(q, t, t_begin, q_start, wl):
    lib = model()
    cap = len(t) // wl + 4
    out = (C.c_uint32 * (4 * cap))()
    lib.ovl_model_break.restype = C.c_int
    lib.ovl_model_break.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.c_uint32]
    n = lib.ovl_model_break(q, len(q), t, len(t), t_begin, q_start, wl, out, cap)
    assert 0 <= n <= cap
    return [(out[4 * i + 2 * h], out[4 * i + 2 * h + 1]) for i in range(n) for h in range(2)]


def cut_cases(seed, count, max_len):
    rng = np.random.default_rng(seed)
    out = []
    for q, t in noisy_pairs(seed, count, max_len, 0.05, 0.12, 0.12):
        if not t:
            continue
        wl = int(rng.choice([50, 100, 500, 640]))
        t_begin = int(rng.choice([0, 1, wl - 1, wl, wl + 1, int(rng.integers(0, 3000))]))
        out.append((q, t, t_begin, int(rng.integers(0, 2000)), wl))
    return out


def test_host_model_breaking_points_match_reference_loop():
    cases = cut_cases(31, 250, 1500) + cut_cases(32, 4, 6000)
    cases += [(b"ACGT" * 30, b"ACGT" * 30, 0, 0, 40), (b"A" * 100, b"A" * 100, 100, 7, 100), (b"C" * 20, b"A" * 20 + b"C" * 20, 90, 3, 10),
              (b"AC" * 60, b"AC" * 10, 495, 0, 500), (b"G", b"ACGTACGT", 0, 0, 4)]
    for q, t, t_begin, q_start, wl in cases:
        cigar, _ = standin_cigar(q, t)
        want = breaking_points_from_cigar(cigar, t_begin, t_begin + len(t), q_start, wl)
        assert model_break(q, t, t_begin, q_start, wl) == want, (len(q), len(t), t_begin, q_start, wl)


def test_abi_exports_and_fails_loudly_without_gpu():
    import re
    import torch
    from vechat_b200 import aligner, engine
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vga.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(vga_[a-z_]+)\s*\(", src)))
    assert names == ["vga_align", "vga_break", "vga_create", "vga_destroy", "vga_last_error"]
    lib = engine.load_library()
    for n in names:
        assert getattr(lib, n) is not None
    assert C.sizeof(aligner.VgaBatch) == 56 and C.sizeof(aligner.VgaResult) == 24 and C.sizeof(aligner.VgaStats) == 40
    assert C.sizeof(aligner.VgaCut) == 24 and C.sizeof(aligner.VgaBreaks) == 24
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError) as ei:
            aligner.Aligner(0)
        assert "no usable CUDA device" in str(ei.value)


# ------------------------------------------------------------------------------------------------ GPU

def pack(pairs):
    blob, q_off, q_len, t_off, t_len = bytearray(), [], [], [], []
    for q, t in pairs:
        q_off.append(len(blob)); q_len.append(len(q)); blob += q
        t_off.append(len(blob)); t_len.append(len(t)); blob += t
    blob += b"\0"
    return np.frombuffer(bytes(blob), np.uint8), q_off, q_len, t_off, t_len


@pytest.mark.gpu
def test_gpu_cigars_match_standin():
    from vechat_b200.aligner import Aligner
    rng = np.random.default_rng(6)
    tiny = [(bytes(rng.choice([65, 67], int(rng.integers(0, 11))).tolist()),
             bytes(rng.choice([65, 67], int(rng.integers(0, 11))).tolist())) for _ in range(600)]
    pairs = EDGE + tiny + noisy_pairs(11, 400, 800) + noisy_pairs(12, 12, 9000) + noisy_pairs(13, 60, 400, 0.3, 0.3, 0.3)
    a = Aligner(0)
    cigars, edits, st = a.align(*pack(pairs))
    for (q, t), c, d in zip(pairs, cigars, edits):
        assert (c, d) == standin_cigar(q, t), (len(q), len(t))
    assert st["kernel_launches"] >= 1 and st["cells"] == sum((d + 1) ** 2 for d in edits)
    # same handle, second call, empty batch, and a batch that shares one sequence buffer between overlaps
    assert a.align(np.zeros(1, np.uint8), [], [], [], [])[0] == []
    pick = len(EDGE) + len(tiny) + 5
    while min(len(pairs[pick][0]), len(pairs[pick][1])) < 20:
        pick += 1
    seq = np.frombuffer(pairs[pick][0] + pairs[pick][1], np.uint8)
    m, n = len(pairs[pick][0]), len(pairs[pick][1])
    c2, _, _ = a.align(seq, [0, 0, 5], [m, m, m - 5], [m, m + 3, m], [n, n - 3, n])
    assert c2[0] == cigars[pick]
    assert c2[1] == standin_cigar(pairs[pick][0], pairs[pick][1][3:])[0]
    assert c2[2] == standin_cigar(pairs[pick][0][5:], pairs[pick][1])[0]
    a.close()


@pytest.mark.gpu
def test_gpu_retry_rounds(monkeypatch):
    """Overlaps that outgrow their wavefront arena or find the round's output buffer full are re-run by the host
    loop with more room; results do not change."""
    from vechat_b200.aligner import Aligner
    pairs = noisy_pairs(21, 300, 700) + noisy_pairs(22, 4, 5000)
    a = Aligner(0)
    want, want_d, st0 = a.align(*pack(pairs))
    assert st0["retried"] == 0 and st0["kernel_launches"] == 1
    monkeypatch.setenv("VGA_ARENA_CELLS", "2500")   # edit distance <= 49 fits; the rest overflows, then x4 per round
    got, got_d, st = a.align(*pack(pairs))
    assert (got, got_d) == (want, want_d) and st["retried"] > 0 and st["kernel_launches"] >= 3
    assert st["cells"] == st0["cells"]
    monkeypatch.delenv("VGA_ARENA_CELLS")
    monkeypatch.setenv("VGA_OUT_CAP_RUNS", "1")      # clamped to one worst-case alignment: many rounds
    got, got_d, st = a.align(*pack(pairs))
    assert (got, got_d) == (want, want_d) and st["retried"] > 0 and st["kernel_launches"] >= 2
    a.close()


@pytest.mark.gpu
def test_gpu_breaking_points_match_reference_loop(monkeypatch):
    """vga_break: alignment + the cut of src/overlap.cpp:226-292 on the device, against the base-by-base restatement
    of the reference loop applied to the host aligner's CIGAR."""
    from vechat_b200.aligner import Aligner
    a = Aligner(0)
    for wl in (100, 500):
        cases = [c for c in cut_cases(41, 500, 1500) + cut_cases(42, 6, 7000)]
        pairs = [(q, t) for q, t, _, _, _ in cases]
        got, edits, st = a.breaks(*pack(pairs), [c[2] for c in cases], [c[3] for c in cases], wl)
        for (q, t, t_begin, q_start, _), g, d in zip(cases, got, edits):
            cigar, dist = standin_cigar(q, t)
            assert d == dist
            assert g == breaking_points_from_cigar(cigar, t_begin, t_begin + len(t), q_start, wl), (len(q), len(t), t_begin, wl)
        assert st["retried"] == 0
    # retry rounds in cut mode
    monkeypatch.setenv("VGA_ARENA_CELLS", "2500")
    monkeypatch.setenv("VGA_OUT_CAP_RUNS", "64")
    got2, _, st = a.breaks(*pack(pairs), [c[2] for c in cases], [c[3] for c in cases], 500)
    assert got2 == got and st["retried"] > 0
    a.close()


@pytest.mark.gpu
def test_gpu_example_overlaps_match_standin():
    """The overlaps of the committed example fixture (tests/golden/example): the substrings Overlap::
    find_breaking_points hands to edlib (src/overlap.cpp:195-199)."""
    from vechat_b200.aligner import Aligner
    ex = os.path.join(GOLDEN, "example")
    reads = {}
    with gzip.open(os.path.join(ex, "reads.fq.gz"), "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        reads[lines[i][1:].decode()] = lines[i + 1]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    pairs = []
    for line in open(os.path.join(ex, "overlaps.paf")):
        f = line.split("\t")
        q, t = reads[f[0]], reads[f[5]]
        qb, qe, tb, te = int(f[2]), int(f[3]), int(f[7]), int(f[8])
        qs = q[qb:qe] if f[4] == "+" else q.translate(comp)[::-1][len(q) - qe:len(q) - qb]
        pairs.append((qs, t[tb:te]))
    pairs = pairs[::3]
    a = Aligner(0)
    cigars, edits, st = a.align(*pack(pairs))
    for (q, t), c, d in zip(pairs, cigars, edits):
        assert (c, d) == standin_cigar(q, t)
    a.close()
