import sys, time
sys.path.insert(0,'.')
import numpy as np
from vechat_b200.engine import Engine
from vechat_b200.sim import Simulator
sim = Simulator("pb_clr_10k_x_10kb")
for T in (100, 400):
    b = sim.windows(0, T)
    eng = Engine(0)
    for it in range(2):
        t0=time.time(); r, st = eng.polish(b); dt=time.time()-t0
    ph = eng.phase_profile()
    n = st['alignments']
    print("T=%d windows=%d wall %.0f ms kernel %.0f ms launches %d alignments %d" % (T, b.n_windows, dt*1e3, st['kernel_ms'], st['kernel_launches'], n))
    print("  per-aln avg cycles: fill %.0f trace %.0f ; max fill %.0f max trace %.0f ; slow steps/aln %.1f refills/aln %.1f" % (
        ph['fill']/n, ph['traceback']/n, ph['max_fill_cycles'], ph['max_trace_cycles'], ph['trace_slow_steps']/n, ph['trace_refills']/n))
    eng.close()
