"""Generates the committed golden fixtures of tests/golden/.  Runs ONLY in the build container (needs
/root/reference); the tests read the fixtures, never the reference tree.

  spoa_sample.json   : the 55 reads of vendor/spoa/test/data/sample.fastq.gz + the expected consensus strings of
                       the four linear-gap cases of vendor/spoa/test/spoa_test.cpp (Local :150-164,
                       LocalWithQualities :198-212, Global :246-260, GlobalWithQualities :294-308), parsed out of
                       the test source.
  windows_*.npz      : seeded window batches (inputs in the vgc_batch layout) + the corrected windows the
                       UNMODIFIED reference (oracle/_ref/libvechat_ref.so = src/window.cpp + vendored spoa,
                       compiled by oracle/Makefile) produces for them, for several parameter sets.

Usage: python tests/golden/make_golden.py
"""
import gzip
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = "/root/reference"

from oracle import checker  # noqa: E402
from vechat_b200._ffi import make_params  # noqa: E402
from vechat_b200.sim import Simulator, fuzz_batch  # noqa: E402

# name -> (fuzz kwargs | ("sim", ...), engine params)
WINDOW_SETS = {
    "hap_fastq": (dict(seed=1, n_windows=24), dict()),
    "hap_fastq_deep": (dict(seed=2, n_windows=6, depth=40, length=160), dict()),
    "hap_fasta": (dict(seed=3, n_windows=16, fastq=False), dict()),
    "hap_fasta_short_last": (dict(seed=4, n_windows=8, fastq=False, window_length=500), dict()),
    "hap_n_nullq": (dict(seed=5, n_windows=16, n_frac=0.03, null_qual=0.3), dict()),
    "hap_k1": (dict(seed=6, n_windows=12), dict(num_prune=1)),
    "hap_k2_d0": (dict(seed=7, n_windows=12), dict(num_prune=2, min_confidence=0.0, min_support=0.0)),
    "hap_d1": (dict(seed=8, n_windows=12), dict(min_confidence=1.0, min_support=1.0)),
    "hap_scores": (dict(seed=9, n_windows=12), dict(match=5, mismatch=-4, gap=-8)),
    "lin_trim": (dict(seed=10, n_windows=24), dict(haplotype=0)),
    "lin_notrim": (dict(seed=11, n_windows=16), dict(haplotype=0, trim=0)),
    "lin_fasta": (dict(seed=12, n_windows=16, fastq=False), dict(haplotype=0)),
    "sim_pb_hap": (("sim", "pb_clr_10k_x_10kb", dict(n_reads=300, genome_len=100_000), 0, 2), dict()),
    "sim_pb_lin": (("sim", "pb_clr_10k_x_10kb", dict(n_reads=300, genome_len=100_000), 2, 4), dict(haplotype=0)),
    "sim_ont_hap": (("sim", "ont_10k_x_20kb", dict(n_reads=200, genome_len=130_000), 0, 1), dict()),
}


def make_batch(spec):
    if isinstance(spec, tuple) and spec[0] == "sim":
        _, cfg, override, t0, t1 = spec
        return Simulator(cfg, **override).windows(t0, t1)
    return fuzz_batch(**spec)


def spoa_sample():
    reads, quals = [], []
    with gzip.open(os.path.join(REF, "vendor/spoa/test/data/sample.fastq.gz"), "rt") as f:
        lines = [l.rstrip("\n") for l in f]
    for i in range(0, len(lines), 4):
        reads.append(lines[i + 1])
        quals.append(lines[i + 3])
    assert len(reads) == 55
    src = open(os.path.join(REF, "vendor/spoa/test/spoa_test.cpp")).read()
    cases = {}
    for name in ("Local", "LocalWithQualities", "Global", "GlobalWithQualities"):
        m = re.search(r"TEST_F\(SpoaTest, %s\) \{(.*?)Check\(c\);" % name, src, re.S)
        body = m.group(1)
        setup = re.search(r"Setup\(AlignmentType::k(\w+), (-?\d+), (-?\d+), (-?\d+), (-?\d+), (-?\d+), (-?\d+), (\w+)\)",
                          body)
        cons = "".join(re.findall(r'"([ACGT]+)"', body))
        cases[name] = dict(type=setup.group(1), m=int(setup.group(2)), n=int(setup.group(3)), g=int(setup.group(4)),
                           e=int(setup.group(5)), with_qualities=setup.group(8) == "true", consensus=cons)
        assert cases[name]["g"] == cases[name]["e"], "linear-gap cases only"
    return dict(source="vendor/spoa/test/data/sample.fastq.gz + vendor/spoa/test/spoa_test.cpp", reads=reads,
                quals=quals, cases=cases)


def main():
    assert os.path.isdir(REF), "needs the reference tree"
    with open(os.path.join(HERE, "spoa_sample.json"), "w") as f:
        json.dump(spoa_sample(), f, indent=0)
    for name, (spec, pkw) in WINDOW_SETS.items():
        batch = make_batch(spec)
        params = make_params(**pkw)
        r = checker.ref_polish(batch, params, threads=8)
        np.savez_compressed(
            os.path.join(HERE, "windows_%s.npz" % name),
            bases=batch.bases, quals=batch.quals, seq_off=batch.seq_off, has_qual=batch.has_qual, begin=batch.begin,
            end=batch.end, win_first=batch.win_first, win_flags=batch.win_flags,
            params=json.dumps(pkw), cons=r.cons, cons_off=r.cons_off, polished=r.polished)
        print("%-24s windows %4d layers %5d corrected bases %7d" % (name, batch.n_windows, batch.n_layers,
                                                                     r.total_bases()))


if __name__ == "__main__":
    main()
