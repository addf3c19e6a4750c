"""Randomised parity stress (run under gpurun): many seeded fuzz batches with random shapes and parameters against
the compiled reference (or the oracle port).  Prints one line per batch and the total number of mismatches."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import checker
from vechat_b200._ffi import make_params
from vechat_b200.engine import Engine
from vechat_b200.sim import fuzz_batch


def main():
    budget = float(os.environ.get("STRESS_SECONDS", "240"))
    rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", "1")))
    fn = checker.ref_polish if checker.have_ref() else checker.oracle_polish
    t_end = time.time() + budget
    bad = n_b = n_w = 0
    engines = {}
    while time.time() < t_end:
        seed = int(rng.integers(1, 1 << 30))
        pkw = dict(haplotype=int(rng.random() < 0.75), trim=int(rng.random() < 0.7),
                   num_prune=int(rng.integers(1, 5)), min_confidence=float(rng.choice([0.0, 0.1, 0.2, 0.35, 1.0])),
                   min_support=float(rng.choice([0.0, 0.1, 0.2, 0.5])))
        if rng.random() < 0.3:
            pkw.update(match=int(rng.integers(1, 6)), mismatch=-int(rng.integers(1, 7)), gap=-int(rng.integers(1, 9)))
        kw = dict(n_windows=int(rng.integers(4, 40)), err=float(rng.choice([0.02, 0.1, 0.15, 0.3])),
                  partial=float(rng.choice([0.0, 0.3, 0.8])), fastq=bool(rng.random() < 0.7),
                  null_qual=float(rng.choice([0.0, 0.0, 0.4])), n_frac=float(rng.choice([0.0, 0.0, 0.03])))
        if rng.random() < 0.5:
            kw["length"] = int(rng.integers(8, 950 if rng.random() < 0.15 else 700))
        if rng.random() < 0.5:
            kw["depth"] = int(rng.integers(0, 60))
        key = tuple(sorted(pkw.items()))
        if key not in engines:
            if len(engines) > 24:
                engines.popitem()[1].close()
            engines[key] = Engine(0, **pkw)
        b = fuzz_batch(seed, **kw)
        got, st = engines[key].polish(b)
        want = fn(b, make_params(**pkw), threads=8)
        m = sum(1 for w in range(b.n_windows) if got.window(w) != want.window(w) or int(got.polished[w]) != int(want.polished[w]))
        bad += m
        n_b += 1
        n_w += b.n_windows
        if m:
            print("MISMATCH seed=%d kw=%s pkw=%s windows=%d bad=%d" % (seed, kw, pkw, b.n_windows, m), flush=True)
    print("stress: %d batches, %d windows, TOTAL MISMATCHES %d" % (n_b, n_w, bad), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
