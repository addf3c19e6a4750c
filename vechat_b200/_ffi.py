"""ctypes mirror of include/vgc.h (the C-ABI) plus the numpy-backed WindowBatch container.

WindowBatch is the Python spelling of what `friend class CUDABatchProcessor` reads out of
racon::Window (reference src/window.hpp:61-76): sequences_, qualities_, positions_, type_.
"""
import ctypes as C

import numpy as np


class VgcParams(C.Structure):
    _fields_ = [
        ("match", C.c_int8),
        ("mismatch", C.c_int8),
        ("gap", C.c_int8),
        ("haplotype", C.c_uint8),
        ("trim", C.c_uint8),
        ("reserved", C.c_uint8 * 3),
        ("num_prune", C.c_uint32),
        ("min_confidence", C.c_double),
        ("min_support", C.c_double),
    ]


class VgcBatch(C.Structure):
    _fields_ = [
        ("n_windows", C.c_uint32),
        ("n_layers", C.c_uint32),
        ("bases", C.c_void_p),
        ("quals", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("has_qual", C.c_void_p),
        ("begin", C.c_void_p),
        ("end", C.c_void_p),
        ("win_first", C.c_void_p),
        ("win_flags", C.c_void_p),
    ]


class VgcResult(C.Structure):
    _fields_ = [
        ("cons", C.c_void_p),
        ("cons_capacity", C.c_uint64),
        ("cons_off", C.c_void_p),
        ("polished", C.c_void_p),
    ]


class VgcStats(C.Structure):
    _fields_ = [
        ("cells", C.c_uint64),
        ("alignments", C.c_uint64),
        ("input_bytes", C.c_uint64),
        ("output_bytes", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("device_ms", C.c_double),
        ("host_prep_ms", C.c_double),
        ("kernel_launches", C.c_uint32),
        ("relaunched_windows", C.c_uint32),
        ("host_pack_ms", C.c_double),
    ]


VGC_WIN_TGS = 1
VGC_WIN_DUMMY_QUAL = 2

# reference defaults: src/main.cpp:46-66 with the driver's pass-1 flags (scripts/vechat:70-72)
DEFAULT_PARAMS = dict(match=3, mismatch=-5, gap=-4, haplotype=1, trim=1, num_prune=3,
                      min_confidence=0.2, min_support=0.2)


def make_params(**kw):
    d = dict(DEFAULT_PARAMS)
    d.update(kw)
    p = VgcParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class WindowBatch:
    """Structure-of-arrays batch of windows (include/vgc.h: vgc_batch) backed by numpy arrays."""

    def __init__(self, bases, quals, seq_off, has_qual, begin, end, win_first, win_flags):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.quals = np.ascontiguousarray(quals, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        self.has_qual = np.ascontiguousarray(has_qual, dtype=np.uint8)
        self.begin = np.ascontiguousarray(begin, dtype=np.uint32)
        self.end = np.ascontiguousarray(end, dtype=np.uint32)
        self.win_first = np.ascontiguousarray(win_first, dtype=np.uint32)
        self.win_flags = np.ascontiguousarray(win_flags, dtype=np.uint8)
        assert len(self.seq_off) == len(self.begin) + 1
        assert len(self.win_first) == len(self.win_flags) + 1
        assert len(self.quals) == len(self.bases)

    @property
    def n_windows(self):
        return len(self.win_flags)

    @property
    def n_layers(self):
        return len(self.begin)

    def c_struct(self):
        b = VgcBatch()
        b.n_windows = self.n_windows
        b.n_layers = self.n_layers
        b.bases = _ptr(self.bases)
        b.quals = _ptr(self.quals)
        b.seq_off = _ptr(self.seq_off)
        b.has_qual = _ptr(self.has_qual)
        b.begin = _ptr(self.begin)
        b.end = _ptr(self.end)
        b.win_first = _ptr(self.win_first)
        b.win_flags = _ptr(self.win_flags)
        return b

    def result_bound(self):
        """Upper bound of corrected bytes: every layer base could become a graph node."""
        return int(len(self.bases)) + 16

    def input_bytes(self):
        return int(sum(a.nbytes for a in (self.bases, self.quals, self.seq_off, self.has_qual, self.begin,
                                           self.end, self.win_first, self.win_flags)))

    def window(self, w):
        """(backbone, backbone_qual, [(seq, qual|None, begin, end), ...]) of window w as bytes."""
        f, l = int(self.win_first[w]), int(self.win_first[w + 1])
        out = []
        for i in range(f, l):
            o0, o1 = int(self.seq_off[i]), int(self.seq_off[i + 1])
            s = self.bases[o0:o1].tobytes()
            q = self.quals[o0:o1].tobytes() if self.has_qual[i] else None
            out.append((s, q, int(self.begin[i]), int(self.end[i])))
        return out

    def select(self, windows):
        """New batch holding only the given window indices (in that order)."""
        windows = [int(w) for w in windows]
        b = WindowBatch.from_windows([(self.window(w), int(self.win_flags[w])) for w in windows])
        for attr in ("win_target", "win_rank"):
            if hasattr(self, attr):
                setattr(b, attr, np.asarray(getattr(self, attr))[windows].copy())
        if hasattr(self, "target_coverage"):
            b.target_coverage = self.target_coverage
        return b

    def slice(self, w0, w1):
        """Contiguous window range [w0, w1) as a new batch (array slicing, no per-window Python work)."""
        l0, l1 = int(self.win_first[w0]), int(self.win_first[w1])
        o0, o1 = int(self.seq_off[l0]), int(self.seq_off[l1])
        b = WindowBatch(self.bases[o0:o1], self.quals[o0:o1], self.seq_off[l0:l1 + 1] - np.uint64(o0),
                        self.has_qual[l0:l1], self.begin[l0:l1], self.end[l0:l1],
                        self.win_first[w0:w1 + 1] - np.uint32(l0), self.win_flags[w0:w1])
        for attr in ("win_target", "win_rank"):
            if hasattr(self, attr):
                setattr(b, attr, np.asarray(getattr(self, attr))[w0:w1].copy())
        if hasattr(self, "target_coverage"):
            b.target_coverage = self.target_coverage
        return b

    @staticmethod
    def from_windows(windows):
        """windows: list of (layers, flags); layers = [(seq bytes, qual bytes|None, begin, end), ...] with
        the backbone first."""
        bases, quals, off, hq, bg, en, wf, fl = [], [], [0], [], [], [], [0], []
        total = 0
        for layers, flags in windows:
            for (s, q, b, e) in layers:
                bases.append(np.frombuffer(s, dtype=np.uint8))
                if q is None:
                    quals.append(np.full(len(s), 33, dtype=np.uint8))
                    hq.append(0)
                else:
                    assert len(q) == len(s)
                    quals.append(np.frombuffer(q, dtype=np.uint8))
                    hq.append(1)
                total += len(s)
                off.append(total)
                bg.append(b)
                en.append(e)
            wf.append(len(bg))
            fl.append(flags)
        cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.uint8))
        return WindowBatch(cat(bases), cat(quals), off, hq, bg, en, wf, fl)

    @staticmethod
    def from_c(bptr):
        """Deep copy of a C vgc_batch (e.g. one produced by the simulator)."""
        b = bptr.contents if hasattr(bptr, "contents") else bptr
        nl, nw = b.n_layers, b.n_windows

        def arr(p, n, ct):
            if n == 0:
                return np.zeros(0, dtype=ct)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy()

        seq_off = arr(b.seq_off, nl + 1, C.c_uint64)
        nb = int(seq_off[-1]) if nl else 0
        return WindowBatch(arr(b.bases, nb, C.c_uint8), arr(b.quals, nb, C.c_uint8), seq_off,
                           arr(b.has_qual, nl, C.c_uint8), arr(b.begin, nl, C.c_uint32),
                           arr(b.end, nl, C.c_uint32), arr(b.win_first, nw + 1, C.c_uint32),
                           arr(b.win_flags, nw, C.c_uint8))


class PolishResult:
    def __init__(self, cons, cons_off, polished):
        self.cons = cons
        self.cons_off = cons_off
        self.polished = polished

    def window(self, w):
        return self.cons[int(self.cons_off[w]):int(self.cons_off[w + 1])].tobytes()

    def strings(self):
        return [self.window(w) for w in range(len(self.polished))]

    def total_bases(self):
        return int(self.cons_off[-1])


def alloc_result(batch):
    cap = batch.result_bound()
    cons = np.empty(cap, dtype=np.uint8)   # written by the engine up to cons_off[n_windows]
    cons_off = np.zeros(batch.n_windows + 1, dtype=np.uint64)
    polished = np.zeros(max(batch.n_windows, 1), dtype=np.uint8)
    r = VgcResult()
    r.cons = _ptr(cons)
    r.cons_capacity = cap
    r.cons_off = _ptr(cons_off)
    r.polished = _ptr(polished)
    return r, (cons, cons_off, polished)


def finish_result(batch, arrays):
    cons, cons_off, polished = arrays
    return PolishResult(cons[:int(cons_off[batch.n_windows])], cons_off, polished[:batch.n_windows])
