"""Tiny polish calls for compute-sanitizer (run under gpurun):
   compute-sanitizer --tool memcheck python tools/sanitize.py
   compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import checker
from vechat_b200._ffi import make_params
from vechat_b200.engine import Engine
from vechat_b200.sim import fuzz_batch

bad = 0
for pkw, seed, kw in ((dict(), 21, dict(n_windows=6, length=90, depth=9)),
                      (dict(), 22, dict(n_windows=4, length=70, depth=8, partial=0.7, n_frac=0.03)),
                      (dict(haplotype=0), 23, dict(n_windows=6, length=90, depth=9, partial=0.5))):
    e = Engine(0, **pkw)
    b = fuzz_batch(seed, **kw)
    got, st = e.polish(b)
    want = checker.oracle_polish(b, make_params(**pkw), threads=4)
    bad += sum(got.window(w) != want.window(w) for w in range(b.n_windows))
    print("sanitize batch ok: %d windows, %d alignments, mismatches so far %d" % (b.n_windows, st["alignments"], bad), flush=True)
    e.close()
sys.exit(1 if bad else 0)
