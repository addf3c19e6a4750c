"""ctypes binding of include/vga.h (the overlap aligner of libvgc.so): harness plumbing for tests and evidence tools.
The product call is vga_align; nothing here computes an alignment."""
import ctypes as C

import numpy as np

from .engine import load_library


class VgaBatch(C.Structure):
    _fields_ = [("seqs", C.c_void_p), ("seqs_len", C.c_uint64), ("n", C.c_uint32), ("q_off", C.c_void_p),
                ("q_len", C.c_void_p), ("t_off", C.c_void_p), ("t_len", C.c_void_p)]


class VgaResult(C.Structure):
    _fields_ = [("cigar", C.c_void_p), ("cigar_off", C.POINTER(C.c_uint64)), ("edit_distance", C.POINTER(C.c_int32))]


class VgaCut(C.Structure):
    _fields_ = [("t_begin", C.c_void_p), ("q_start", C.c_void_p), ("window_length", C.c_uint32)]


class VgaBreaks(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_uint32)), ("points_off", C.POINTER(C.c_uint64)),
                ("edit_distance", C.POINTER(C.c_int32))]


class VgaStats(C.Structure):
    _fields_ = [("cells", C.c_uint64), ("wavefront_bytes", C.c_uint64), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double), ("kernel_launches", C.c_uint32), ("retried", C.c_uint32)]


def _lib():
    lib = load_library()
    if not getattr(lib, "_vga_ready", False):
        lib.vga_create.restype = C.c_int
        lib.vga_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        lib.vga_destroy.argtypes = [C.c_void_p]
        lib.vga_align.restype = C.c_int
        lib.vga_align.argtypes = [C.c_void_p, C.POINTER(VgaBatch), C.POINTER(VgaResult), C.POINTER(VgaStats)]
        lib.vga_break.restype = C.c_int
        lib.vga_break.argtypes = [C.c_void_p, C.POINTER(VgaBatch), C.POINTER(VgaCut), C.POINTER(VgaBreaks),
                                  C.POINTER(VgaStats)]
        lib.vga_last_error.restype = C.c_char_p
        lib._vga_ready = True
    return lib


class Aligner:
    """One vga_handle.  align(seqs, q_off, q_len, t_off, t_len) -> (list of CIGAR strings, edit distances, stats)."""

    def __init__(self, device=0):
        self.lib = _lib()
        self.h = C.c_void_p()
        rc = self.lib.vga_create(C.byref(self.h), device)
        if rc != 0:
            raise RuntimeError("vga_create failed (%d): %s" % (rc, self.lib.vga_last_error().decode()))

    def align(self, seqs, q_off, q_len, t_off, t_len):
        seqs = np.ascontiguousarray(seqs, np.uint8)
        q_off, t_off = np.ascontiguousarray(q_off, np.uint64), np.ascontiguousarray(t_off, np.uint64)
        q_len, t_len = np.ascontiguousarray(q_len, np.uint32), np.ascontiguousarray(t_len, np.uint32)
        n = len(q_off)
        b = VgaBatch(seqs.ctypes.data, seqs.size, n, q_off.ctypes.data, q_len.ctypes.data, t_off.ctypes.data,
                     t_len.ctypes.data)
        r, st = VgaResult(), VgaStats()
        rc = self.lib.vga_align(self.h, C.byref(b), C.byref(r), C.byref(st))
        if rc != 0:
            raise RuntimeError("vga_align failed (%d): %s" % (rc, self.lib.vga_last_error().decode()))
        cigars = [C.string_at(r.cigar + r.cigar_off[i]).decode() for i in range(n)]
        edits = [r.edit_distance[i] for i in range(n)]
        return cigars, edits, {f[0]: getattr(st, f[0]) for f in VgaStats._fields_}

    def breaks(self, seqs, q_off, q_len, t_off, t_len, t_begin, q_start, window_length):
        """-> (per overlap: list of (t, q) breaking points as Overlap::breaking_points_ holds them, edits, stats)."""
        seqs = np.ascontiguousarray(seqs, np.uint8)
        q_off, t_off = np.ascontiguousarray(q_off, np.uint64), np.ascontiguousarray(t_off, np.uint64)
        q_len, t_len = np.ascontiguousarray(q_len, np.uint32), np.ascontiguousarray(t_len, np.uint32)
        t_begin, q_start = np.ascontiguousarray(t_begin, np.uint32), np.ascontiguousarray(q_start, np.uint32)
        n = len(q_off)
        b = VgaBatch(seqs.ctypes.data, seqs.size, n, q_off.ctypes.data, q_len.ctypes.data, t_off.ctypes.data,
                     t_len.ctypes.data)
        c = VgaCut(t_begin.ctypes.data, q_start.ctypes.data, window_length)
        r, st = VgaBreaks(), VgaStats()
        rc = self.lib.vga_break(self.h, C.byref(b), C.byref(c), C.byref(r), C.byref(st))
        if rc != 0:
            raise RuntimeError("vga_break failed (%d): %s" % (rc, self.lib.vga_last_error().decode()))
        out = []
        for i in range(n):
            lo, hi = r.points_off[i], r.points_off[i + 1]
            out.append([(r.points[4 * p + 2 * h], r.points[4 * p + 2 * h + 1]) for p in range(lo, hi) for h in range(2)])
        return out, [r.edit_distance[i] for i in range(n)], {f[0]: getattr(st, f[0]) for f in VgaStats._fields_}

    def close(self):
        if self.h:
            self.lib.vga_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
