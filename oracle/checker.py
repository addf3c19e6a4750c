"""TEST INFRASTRUCTURE — Python loaders for the two checkers under oracle/.

* `ref_*`    : oracle/_ref/libvechat_ref.so — the UNMODIFIED reference hot path compiled from /root/reference
               (oracle/Makefile `ref`); kind = "reference".
* `oracle_*` : oracle/_build/liboracle.so — our CPU restatement (oracle/poa_oracle.cpp); kind = "port".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (vechat_b200.engine) never does.
"""
import ctypes as C
import os
import subprocess

from vechat_b200._ffi import VgcBatch, VgcParams, VgcResult, alloc_result, finish_result

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libvechat_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")

_ref = None
_oracle = None


def have_ref():
    return os.path.exists(REF_SO)


def _load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
            subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        lib = C.CDLL(REF_SO)
        lib.ref_polish.restype = C.c_int
        lib.ref_polish.argtypes = [C.POINTER(VgcBatch), C.POINTER(VgcParams), C.POINTER(VgcResult), C.c_int]
        lib.ref_spoa_consensus.restype = C.c_int
        lib.ref_align_probe.restype = C.c_int
        _ref = lib
    return _ref


def _load_oracle():
    global _oracle
    if _oracle is None:
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_polish.restype = C.c_int
        lib.oracle_polish.argtypes = [C.POINTER(VgcBatch), C.POINTER(VgcParams), C.POINTER(VgcResult), C.c_int]
        lib.oracle_spoa_consensus.restype = C.c_int
        lib.oracle_align_probe.restype = C.c_int
        _oracle = lib
    return _oracle


def _polish(fn, batch, params, threads):
    b = batch.c_struct()
    r, arrays = alloc_result(batch)
    rc = fn(C.byref(b), C.byref(params), C.byref(r), int(threads))
    if rc != 0:
        raise RuntimeError("checker polish failed rc=%d" % rc)
    return finish_result(batch, arrays)


def ref_polish(batch, params, threads=1):
    return _polish(_load_ref().ref_polish, batch, params, threads)


REF_AVX2_SO = os.path.join(HERE, "_ref", "libvechat_ref_avx2.so")
_ref_avx2 = None


def have_ref_avx2():
    """The -mavx2 build of the reference (second CPU baseline) is there and this CPU can run it."""
    if not os.path.exists(REF_AVX2_SO):
        return False
    try:
        return " avx2 " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


def ref_avx2_polish(batch, params, threads=1):
    global _ref_avx2
    if _ref_avx2 is None:
        lib = C.CDLL(REF_AVX2_SO)
        lib.ref_polish.restype = C.c_int
        lib.ref_polish.argtypes = [C.POINTER(VgcBatch), C.POINTER(VgcParams), C.POINTER(VgcResult), C.c_int]
        _ref_avx2 = lib
    return _polish(_ref_avx2.ref_polish, batch, params, threads)


def oracle_polish(batch, params, threads=1):
    return _polish(_load_oracle().oracle_polish, batch, params, threads)


def _spoa(fn, type_, m, n, g, seqs, quals):
    k = len(seqs)
    sa = (C.c_char_p * k)(*seqs)
    qa = (C.c_char_p * k)(*quals) if quals is not None else None
    cap = sum(len(s) for s in seqs) + 16
    out = C.create_string_buffer(cap)
    rc = fn(C.c_int(type_), C.c_int(m), C.c_int(n), C.c_int(g), C.c_uint32(k), sa, qa, out, C.c_uint32(cap))
    if rc < 0:
        raise RuntimeError("spoa consensus failed")
    return out.value


def ref_spoa_consensus(type_, m, n, g, seqs, quals=None):
    return _spoa(_load_ref().ref_spoa_consensus, type_, m, n, g, seqs, quals)


def oracle_spoa_consensus(type_, m, n, g, seqs, quals=None):
    return _spoa(_load_oracle().oracle_spoa_consensus, type_, m, n, g, seqs, quals)


def _probe(fn, type_, m, x, g, seqs, query):
    k = len(seqs)
    sa = (C.c_char_p * k)(*seqs)
    cap = 4 * (sum(len(s) for s in seqs) + len(query)) + 16
    nodes = (C.c_int32 * cap)()
    pos = (C.c_int32 * cap)()
    n = fn(C.c_int(type_), C.c_int(m), C.c_int(x), C.c_int(g), C.c_uint32(k), sa, C.c_char_p(query), nodes, pos,
           C.c_uint32(cap))
    if n < 0:
        raise RuntimeError("align probe failed")
    return [(nodes[i], pos[i]) for i in range(n)]


def ref_align_probe(type_, m, x, g, seqs, query):
    return _probe(_load_ref().ref_align_probe, type_, m, x, g, seqs, query)


def oracle_align_probe(type_, m, x, g, seqs, query):
    return _probe(_load_oracle().oracle_align_probe, type_, m, x, g, seqs, query)


# ---- the host aligner both reference-program builds use in place of edlib (oracle/shims/edlib_standin.cpp): the
# checker of the overlap aligner (include/vga.h).

EDLIB_SO = os.path.join(HERE, "_build", "libedlib_standin.so")
_edlib = None


class _EdlibConfig(C.Structure):
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("eq", C.c_void_p), ("neq", C.c_int)]


class _EdlibResult(C.Structure):
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)),
                ("startLocations", C.POINTER(C.c_int)), ("numLocations", C.c_int),
                ("alignment", C.POINTER(C.c_ubyte)), ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


def edlib_standin():
    global _edlib
    if _edlib is None:
        if not os.path.exists(EDLIB_SO):
            subprocess.run(["make", "-s", "-C", HERE, "edlib"], check=True)
        lib = C.CDLL(EDLIB_SO)
        lib.edlibNewAlignConfig.restype = _EdlibConfig
        lib.edlibNewAlignConfig.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.edlibAlign.restype = _EdlibResult
        lib.edlibAlign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, _EdlibConfig]
        lib.edlibAlignmentToCigar.restype = C.c_void_p
        lib.edlibAlignmentToCigar.argtypes = [C.POINTER(C.c_ubyte), C.c_int, C.c_int]
        lib.edlibFreeAlignResult.argtypes = [_EdlibResult]
        _edlib = lib
    return _edlib


def standin_cigar(q, t):
    """(standard CIGAR, edit distance) of the global alignment of bytes q against bytes t."""
    lib = edlib_standin()
    r = lib.edlibAlign(q, len(q), t, len(t), lib.edlibNewAlignConfig(-1, 0, 2, None, 0))  # EDLIB_MODE_NW, TASK_PATH
    if r.status != 0:
        raise RuntimeError("edlib stand-in failed")
    p = lib.edlibAlignmentToCigar(r.alignment, r.alignmentLength, 0)
    s = C.string_at(p).decode()
    C.CDLL(None).free(C.c_void_p(p))
    d = r.editDistance
    lib.edlibFreeAlignResult(r)
    return s, d
