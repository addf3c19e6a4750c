"""Quick GPU parity + timing probe (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vechat_b200.sim import fuzz_batch, Simulator
from vechat_b200._ffi import make_params
from vechat_b200.engine import Engine
from oracle import checker

def compare(name, batch, eng, p, threads=8):
    t0 = time.time(); r, st = eng.polish(batch); t1 = time.time()
    o = checker.oracle_polish(batch, p, threads=threads) if not checker.have_ref() else checker.ref_polish(batch, p, threads=threads)
    t2 = time.time()
    bad = [w for w in range(batch.n_windows) if r.window(w) != o.window(w) or r.polished[w] != o.polished[w]]
    print("%-28s windows %5d mismatches %4d  gpu %.3fs (kernel %.1f ms, cells %.3g, aln %d, relaunch %d) cpu %.2fs" % (
        name, batch.n_windows, len(bad), t1 - t0, st["kernel_ms"], st["cells"], st["alignments"], st["relaunched_windows"], t2 - t1), flush=True)
    return len(bad)

def main():
    bad = 0
    p = make_params(); eng = Engine(0)
    for seed in range(6):
        bad += compare("fuzz hap seed %d" % seed, fuzz_batch(seed, n_windows=16), eng, p)
    bad += compare("fuzz hap fasta", fuzz_batch(77, n_windows=16, fastq=False), eng, p)
    bad += compare("fuzz hap N/nullq", fuzz_batch(78, n_windows=16, n_frac=0.02, null_qual=0.3), eng, p)
    pl = make_params(haplotype=0); engl = Engine(0, haplotype=0)
    for seed in range(3):
        bad += compare("fuzz lin seed %d" % seed, fuzz_batch(100 + seed, n_windows=16), engl, pl)
    sim = Simulator("pb_clr_10k_x_10kb", n_reads=600, genome_len=200_000)
    b = sim.windows(0, 8)
    bad += compare("sim pb hap 8 targets", b, eng, p)
    bad += compare("sim pb lin 8 targets", b, engl, pl)
    nt = int(os.environ.get("NT", "120"))
    b = sim.windows(0, nt)
    t0 = time.time(); r, st = eng.polish(b); t1 = time.time()
    print("sim pb hap %d targets: windows %d, e2e %.3fs, kernel %.1f ms -> %.0f windows/s (kernel), cells %.4g, %.1f GB/s algorithmic" % (
        nt, b.n_windows, t1 - t0, st["kernel_ms"], b.n_windows / (st["kernel_ms"] / 1e3), st["cells"], 2 * st["cells"] / (st["kernel_ms"] / 1e3) / 1e9), flush=True)
    print("TOTAL MISMATCHES", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
