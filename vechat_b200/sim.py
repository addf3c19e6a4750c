"""Synthetic workloads of BASELINE.json — ctypes front end of csrc/sim.cpp — and a small
numpy window fuzzer for the parity tests."""
import ctypes as C
import os

import numpy as np

from ._ffi import VgcBatch, WindowBatch, VGC_WIN_TGS, VGC_WIN_DUMMY_QUAL

_LIB = None


class SimConfig(C.Structure):
    _fields_ = [
        ("genome_len", C.c_uint64),
        ("n_reads", C.c_uint32),
        ("read_len", C.c_uint32),
        ("p_ins", C.c_double),
        ("p_del", C.c_double),
        ("p_sub", C.c_double),
        ("q_mean", C.c_double),
        ("q_sd", C.c_double),
        ("q_lo", C.c_int32),
        ("q_hi", C.c_int32),
        ("window_len", C.c_uint32),
        ("min_overlap", C.c_uint32),
        ("n_haplotypes", C.c_uint32),
        ("snp_rate", C.c_double),
        ("fasta", C.c_uint32),
        ("random_strand", C.c_uint32),
        ("seed", C.c_uint64),
    ]


# BASELINE.json configs[1..3] as SURVEY.md §8(d) spells them out.
CONFIGS = {
    # synthetic PacBio CLR, 10k reads x 10 kb, 15 % error, 500 bp windows
    "pb_clr_10k_x_10kb": dict(genome_len=3_330_000, n_reads=10_000, read_len=10_000, p_ins=0.09, p_del=0.045,
                              p_sub=0.015, q_mean=12.0, q_sd=2.0, q_lo=2, q_hi=20, window_len=500,
                              min_overlap=500, n_haplotypes=1, snp_rate=0.0, fasta=0, random_strand=1,
                              seed=20260001),
    # synthetic ONT, 10k reads x 20 kb, 10 % error
    "ont_10k_x_20kb": dict(genome_len=6_670_000, n_reads=10_000, read_len=20_000, p_ins=0.03, p_del=0.04,
                           p_sub=0.03, q_mean=14.0, q_sd=3.0, q_lo=2, q_hi=30, window_len=500,
                           min_overlap=500, n_haplotypes=1, snp_rate=0.0, fasta=0, random_strand=1,
                           seed=20260002),
    # 2-haplotype 50:50 mix, 50k reads x 12 kb
    "hap2_50k_x_12kb": dict(genome_len=10_000_000, n_reads=50_000, read_len=12_000, p_ins=0.09, p_del=0.045,
                            p_sub=0.015, q_mean=12.0, q_sd=2.0, q_lo=2, q_hi=20, window_len=500,
                            min_overlap=500, n_haplotypes=2, snp_rate=0.001, fasta=0, random_strand=1,
                            seed=20260003),
}


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libvgcsim.so")
        if not os.path.exists(path):
            from . import build
            build.build_sim()
        lib = C.CDLL(path)
        lib.sim_create.restype = C.c_void_p
        lib.sim_create.argtypes = [C.POINTER(SimConfig)]
        lib.sim_destroy.argtypes = [C.c_void_p]
        lib.sim_windows.restype = C.c_void_p
        lib.sim_windows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_double]
        lib.sim_batch_view.restype = C.POINTER(VgcBatch)
        lib.sim_batch_view.argtypes = [C.c_void_p]
        for fn in ("sim_batch_targets", "sim_batch_ranks", "sim_batch_target_coverages"):
            getattr(lib, fn).restype = C.POINTER(C.c_uint32)
            getattr(lib, fn).argtypes = [C.c_void_p]
        lib.sim_batch_num_targets.restype = C.c_uint32
        lib.sim_batch_num_targets.argtypes = [C.c_void_p]
        lib.sim_batch_overlaps.restype = C.c_uint64
        lib.sim_batch_overlaps.argtypes = [C.c_void_p]
        lib.sim_free_batch.argtypes = [C.c_void_p]
        lib.sim_export.restype = C.c_longlong
        lib.sim_export.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.sim_get_read.restype = C.c_uint32
        lib.sim_get_read.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_uint32]
        _LIB = lib
    return _LIB


class Simulator:
    """Genome + reads of one synthetic config; windows(t0, t1) tiles targets [t0, t1) into 500 bp windows and
    attaches every overlapping read as layers."""

    def __init__(self, config="pb_clr_10k_x_10kb", **override):
        d = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
        d.update(override)
        self.cfg = SimConfig(**d)
        self.params = d
        self._h = _lib().sim_create(C.byref(self.cfg))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().sim_destroy(self._h)
            self._h = None

    @property
    def n_reads(self):
        return self.cfg.n_reads

    def read(self, r):
        cap = int(self.cfg.read_len) + 8
        s = C.create_string_buffer(cap)
        q = C.create_string_buffer(cap)
        n = _lib().sim_get_read(self._h, r, s, q, cap)
        return s.raw[:n], q.raw[:n]

    def export(self, reads_path, paf_path):
        """Write the reads (FASTQ / FASTA) and their ground-truth all-vs-all overlaps (PAF) — the input files of the
        whole program (vechat_racon <reads> <overlaps> <targets>).  Returns the number of overlaps."""
        n = _lib().sim_export(self._h, os.fsencode(reads_path), os.fsencode(paf_path))
        if n < 0:
            raise OSError("sim_export could not write %s / %s" % (reads_path, paf_path))
        return int(n)

    def windows(self, t0, t1, quality_threshold=10.0):
        """Windows of target reads [t0, t1).  The returned batch also carries win_target / win_rank (read id and
        window index, Window::id_/rank_) and target_coverage {read id: #overlaps} for the stitcher."""
        L = _lib()
        sb = L.sim_windows(self._h, t0, t1, quality_threshold)
        try:
            b = WindowBatch.from_c(L.sim_batch_view(sb))
            nw = b.n_windows
            b.win_target = np.ctypeslib.as_array(L.sim_batch_targets(sb), shape=(max(nw, 1),))[:nw].copy()
            b.win_rank = np.ctypeslib.as_array(L.sim_batch_ranks(sb), shape=(max(nw, 1),))[:nw].copy()
            nt = L.sim_batch_num_targets(sb)
            cov = np.ctypeslib.as_array(L.sim_batch_target_coverages(sb), shape=(max(nt, 1),))[:nt].copy()
            b.target_coverage = {int(t0 + i): int(c) for i, c in enumerate(cov)}
            return b
        finally:
            _lib().sim_free_batch(sb)


# ---------------------------------------------------------------------------------------------------
# window-level fuzzer (numpy; small cases for the parity tests)

def _mutate(rng, truth, e_ins, e_del, e_sub, alphabet):
    out = []
    for c in truth:
        u = rng.random()
        if u < e_del:
            pass
        elif u < e_del + e_sub:
            out.append(int(rng.choice([a for a in alphabet if a != c] or [c])))
        else:
            out.append(int(c))
        while rng.random() < e_ins:
            out.append(int(rng.choice(alphabet)))
    if not out:
        out.append(int(truth[0]))
    return out


def fuzz_window(rng, length=120, depth=8, err=0.15, partial=0.3, fastq=True, null_qual=0.0, n_frac=0.0,
                n_hap=2, dummy_backbone=False, window_length=None, iupac_frac=0.0):
    """One random window: two haplotypes of a random truth, backbone + `depth` noisy layers, a fraction of
    which are partial spans.  Returns (layers, flags) in WindowBatch.from_windows form."""
    alphabet = [65, 67, 71, 84]
    truth = rng.choice(alphabet, size=length).astype(np.uint8)
    haps = [truth.copy() for _ in range(max(1, n_hap))]
    for h in haps[1:]:
        for _ in range(max(1, length // 60)):
            p = int(rng.integers(0, length))
            h[p] = rng.choice([a for a in alphabet if a != h[p]])
    e_ins, e_del, e_sub = 0.6 * err, 0.3 * err, 0.1 * err

    def noisy(seg):
        s = _mutate(rng, seg, e_ins, e_del, e_sub, alphabet)
        s = np.array(s, dtype=np.uint8)
        if n_frac > 0:
            m = rng.random(len(s)) < n_frac
            s[m] = ord("N")
        if iupac_frac > 0:  # ambiguity codes: up to 15 distinct bytes with A C G T
            m = rng.random(len(s)) < iupac_frac
            s[m] = rng.choice(np.frombuffer(b"NRYKMSWBDHV", dtype=np.uint8), size=int(m.sum()))
        return s

    def qual(n):
        return np.clip(np.rint(rng.normal(12, 4, size=n)), 1, 40).astype(np.uint8) + 33

    backbone = noisy(haps[0])
    blen = len(backbone)
    flags = VGC_WIN_TGS
    if dummy_backbone or not fastq:
        bq = np.full(blen, 33, dtype=np.uint8)
        if window_length is None or blen == window_length:
            flags |= VGC_WIN_DUMMY_QUAL
    else:
        bq = qual(blen)
    layers = [(backbone.tobytes(), bq.tobytes(), 0, 0)]
    for _ in range(depth):
        h = haps[int(rng.integers(0, len(haps)))]
        if rng.random() < partial and blen > 8:
            b = int(rng.integers(0, blen - 2))
            e = int(rng.integers(b + 1, blen))
        else:
            b, e = 0, blen - 1
        # the layer covers backbone positions [b, e]; take the matching stretch of the truth (approximate)
        tb = min(length - 1, int(b * length / blen))
        te = max(tb + 1, min(length, int((e + 1) * length / blen)))
        s = noisy(h[tb:te])
        if fastq and rng.random() >= null_qual:
            q = qual(len(s)).tobytes()
        else:
            q = None
        layers.append((s.tobytes(), q, b, e))
    return layers, flags


def fuzz_batch(seed, n_windows=16, **kw):
    rng = np.random.default_rng(seed)
    wins = []
    for _ in range(n_windows):
        k = dict(kw)
        if "length" not in k:
            k["length"] = int(rng.integers(20, 200))
        if "depth" not in k:
            k["depth"] = int(rng.integers(0, 14))
        wins.append(fuzz_window(rng, **k))
    return WindowBatch.from_windows(wins)
