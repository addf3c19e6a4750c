#!/usr/bin/env python
"""Occupancy timeline from a VGC_TIMELINE dump (vgc_engine.cu tl_record): resident warps per SM by kernel kind over time.

usage: tools/timeline.py <file> [bucket_ms]"""
import sys
import numpy as np

KINDS = ["update", "sort", "align NW", "align SW", "align NW round", "align SW round"]


def main():
    path = sys.argv[1]
    bucket = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 20e6
    r = np.fromfile(path, dtype=np.uint32).reshape(-1, 4)
    t0 = r[:, 0].astype(np.uint64) | (r[:, 1].astype(np.uint64) << np.uint64(32))
    dur = r[:, 2].astype(np.float64)
    sm = r[:, 3] & 0xFFFF
    kind = r[:, 3] >> 16
    base = t0.min()
    ts = (t0 - base).astype(np.float64)
    te = ts + dur
    span = te.max()
    nb = int(span / bucket) + 1
    nsm = int(sm.max()) + 1
    print("records %d, span %.1f ms, SMs %d" % (len(r), span / 1e6, nsm))
    occ = np.zeros((len(KINDS), nb))
    for k in range(len(KINDS)):
        m = kind == k
        if not m.any():
            continue
        s, e = ts[m], te[m]
        for b in range(nb):
            lo, hi = b * bucket, (b + 1) * bucket
            occ[k, b] = np.clip(np.minimum(e, hi) - np.maximum(s, lo), 0, None).sum() / bucket / nsm
    print("avg resident warps per SM per %.0f ms bucket" % (bucket / 1e6))
    print("%8s " % "t(ms)" + " ".join("%14s" % k for k in KINDS) + "   total")
    for b in range(nb):
        print("%8.0f " % (b * bucket / 1e6) + " ".join("%14.2f" % occ[k, b] for k in range(len(KINDS))) + "   %5.2f" % occ[:, b].sum())
    print("%8s " % "mean" + " ".join("%14.2f" % occ[k].mean() for k in range(len(KINDS))) + "   %5.2f" % occ.sum(0).mean())
    for k in range(len(KINDS)):
        m = kind == k
        if m.any():
            print("%-16s n %8d  mean %8.1f us  p50 %8.1f  p99 %8.1f  max %8.1f" % (
                KINDS[k], m.sum(), dur[m].mean() / 1e3, np.percentile(dur[m], 50) / 1e3, np.percentile(dur[m], 99) / 1e3, dur[m].max() / 1e3))


if __name__ == "__main__":
    main()
