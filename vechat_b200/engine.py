"""Python face of the C-ABI engine (include/vgc.h, built as vechat_b200/lib/libvgc.so).

Mirrors the reference's per-window contract — Window::generate_consensus (reference src/window.hpp:47-51)
applied to every window of a batch, as Polisher::polish does (src/polisher.cpp:491-517).  The library
is loaded from the in-tree build; if it is missing, or no Blackwell GPU is usable, this module raises —
there is no CPU path behind it.
"""
import ctypes as C
import os

from ._ffi import (VgcBatch, VgcParams, VgcResult, VgcStats, WindowBatch, PolishResult, alloc_result,
                   finish_result, make_params)

_LIB = None
LIB_PATH = os.environ.get("VGC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libvgc.so")  # VGC_LIB: experiments

EXPORTS = ["vgc_create", "vgc_destroy", "vgc_result_bound", "vgc_polish", "vgc_upload", "vgc_polish_resident",
           "vgc_last_error", "vgc_version", "vgc_weight_lut", "vgc_phase_profile", "vgc_limits", "vgc_window_status", "vgc_submit",
           "vgc_collect"]


class VgcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vgc error %d: %s" % (code, msg))
        self.code = code


def load_library():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libvgc.so is not built (%s): run `python -m vechat_b200.build`; "
                              "the engine has no fallback path" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.vgc_create.restype = C.c_int
        lib.vgc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(VgcParams)]
        lib.vgc_destroy.restype = C.c_int
        lib.vgc_destroy.argtypes = [C.c_void_p]
        lib.vgc_result_bound.restype = C.c_uint64
        lib.vgc_result_bound.argtypes = [C.POINTER(VgcBatch)]
        lib.vgc_polish.restype = C.c_int
        lib.vgc_polish.argtypes = [C.c_void_p, C.POINTER(VgcBatch), C.POINTER(VgcResult), C.POINTER(VgcStats)]
        lib.vgc_submit.restype = C.c_int
        lib.vgc_submit.argtypes = [C.c_void_p, C.POINTER(VgcBatch)]
        lib.vgc_collect.restype = C.c_int
        lib.vgc_collect.argtypes = [C.c_void_p, C.POINTER(VgcResult), C.POINTER(VgcStats)]
        lib.vgc_upload.restype = C.c_int
        lib.vgc_upload.argtypes = [C.c_void_p, C.POINTER(VgcBatch)]
        lib.vgc_polish_resident.restype = C.c_int
        lib.vgc_polish_resident.argtypes = [C.c_void_p, C.POINTER(VgcResult), C.POINTER(VgcStats)]
        lib.vgc_phase_profile.restype = C.c_int
        lib.vgc_phase_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.vgc_last_error.restype = C.c_char_p
        lib.vgc_version.restype = C.c_char_p
        lib.vgc_weight_lut.argtypes = [C.POINTER(C.c_uint32)]
        lib.vgc_limits.restype = None
        lib.vgc_limits.argtypes = [C.POINTER(C.c_uint32)]
        lib.vgc_window_status.restype = C.c_int
        lib.vgc_window_status.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32]
        _LIB = lib
    return _LIB


class Engine:
    """One engine per GPU (vgc_create).  polish(batch) == the reference's polish stage for that batch."""

    def __init__(self, device=0, **params):
        self.lib = load_library()
        self.params = make_params(**params)
        self._h = C.c_void_p()
        rc = self.lib.vgc_create(C.byref(self._h), int(device), C.byref(self.params))
        if rc != 0:
            raise VgcError(rc, self.lib.vgc_last_error().decode())
        self._resident = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.vgc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VgcError(rc, self.lib.vgc_last_error().decode())

    def polish(self, batch: WindowBatch):
        """Host buffers in, host buffers out.  Returns (PolishResult, stats dict)."""
        b = batch.c_struct()
        r, arrays = alloc_result(batch)
        st = VgcStats()
        self._check(self.lib.vgc_polish(self._h, C.byref(b), C.byref(r), C.byref(st)))
        return finish_result(batch, arrays), _stats(st)

    def submit(self, batch: WindowBatch):
        """vgc_submit: stage a batch (host prepare, packing, H2D) on the engine's worker thread; returns at once.
        Submit batch i + 1 before collect() of batch i to overlap its staging with the kernels of batch i."""
        b = batch.c_struct()
        if not hasattr(self, "_inflight"):
            self._inflight = []
        self._inflight.append((batch, b))  # the arrays must outlive the collect
        self._check(self.lib.vgc_submit(self._h, C.byref(b)))

    def collect(self):
        """vgc_collect: kernels + D2H of the oldest submitted batch.  Returns (PolishResult, stats dict)."""
        batch, _b = self._inflight[0]
        r, arrays = alloc_result(batch)
        st = VgcStats()
        rc = self.lib.vgc_collect(self._h, C.byref(r), C.byref(st))
        self._inflight.pop(0)
        self._check(rc)
        return finish_result(batch, arrays), _stats(st)

    def limits(self):
        """vgc_limits(): the engine's hard limits (include/vgc.h)."""
        out = (C.c_uint32 * 8)()
        self.lib.vgc_limits(out)
        names = ["max_layer_len", "max_backbone_len", "max_codes", "fast_layer_len", "int16_score_bound"]
        return {n: int(out[i]) for i, n in enumerate(names)}

    def window_status(self, n_windows):
        """Per-window status codes of the last polish call (0 = ok)."""
        out = (C.c_uint32 * max(1, n_windows))()
        self._check(self.lib.vgc_window_status(self._h, out, n_windows))
        return list(out)[:n_windows]

    def upload(self, batch: WindowBatch):
        b = batch.c_struct()
        self._check(self.lib.vgc_upload(self._h, C.byref(b)))
        self._resident = batch

    def polish_resident(self, want_result=True):
        batch = self._resident
        st = VgcStats()
        if want_result:
            r, arrays = alloc_result(batch)
            self._check(self.lib.vgc_polish_resident(self._h, C.byref(r), C.byref(st)))
            return finish_result(batch, arrays), _stats(st)
        self._check(self.lib.vgc_polish_resident(self._h, None, C.byref(st)))
        return None, _stats(st)


PHASES = ["trace_slow_steps", "toposort", "rowprog", "fill", "traceback", "add_alignment", "add_weights", "prune",
          "largest_subgraph", "emit", "trace_refills", "host_launch_ms", "sorts", "sorts_out_of_hbm",
          "max_fill_cycles", "max_trace_cycles"]
# entries of the profile that are counts / maxima / host times, not per-window cycle sums
PHASE_NOT_CYCLES = ("trace_slow_steps", "trace_refills", "host_launch_ms", "sorts", "sorts_out_of_hbm",
                    "max_fill_cycles", "max_trace_cycles")


def _phase_profile(engine):
    out = (C.c_double * 16)()
    engine.lib.vgc_phase_profile(engine._h, out)
    return {n: out[i] for i, n in enumerate(PHASES)}


Engine.phase_profile = _phase_profile


def _stats(st):
    return {k: getattr(st, k) for k, _ in VgcStats._fields_}


def weight_lut():
    lut = (C.c_uint32 * 256)()
    load_library().vgc_weight_lut(lut)
    return list(lut)
