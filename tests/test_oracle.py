"""The oracle (oracle/poa_oracle.cpp, CPU restatement) pinned against
  (a) the reference's own golden vectors for this path: the linear-gap SW/NW consensus cases of
      vendor/spoa/test/spoa_test.cpp on vendor/spoa/test/data/sample.fastq.gz (tests/golden/spoa_sample.json);
  (b) committed outputs of the unmodified reference compiled here (tests/golden/windows_*.npz);
  (c) when oracle/_ref/libvechat_ref.so is present: the live reference on further seeded windows and on single
      alignments (node/position pairs, i.e. the traceback tie-breaking itself)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same, golden_names, load_golden
from oracle import checker
from vechat_b200._ffi import make_params
from vechat_b200.sim import fuzz_batch

SPOA = json.load(open(os.path.join(GOLDEN, "spoa_sample.json")))


@pytest.mark.parametrize("case", sorted(SPOA["cases"]))
def test_spoa_golden(case):
    c = SPOA["cases"][case]
    seqs = [s.encode() for s in SPOA["reads"]]
    quals = [q.encode() for q in SPOA["quals"]] if c["with_qualities"] else None
    got = checker.oracle_spoa_consensus(0 if c["type"] == "SW" else 1, c["m"], c["n"], c["g"], seqs, quals)
    assert got.decode() == c["consensus"]


@pytest.mark.skipif(not checker.have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("case", sorted(SPOA["cases"]))
def test_spoa_golden_holds_for_compiled_reference(case):
    """VeChat's fork changed the quality->weight formula (graph.cpp:169); the goldens still hold."""
    c = SPOA["cases"][case]
    seqs = [s.encode() for s in SPOA["reads"]]
    quals = [q.encode() for q in SPOA["quals"]] if c["with_qualities"] else None
    got = checker.ref_spoa_consensus(0 if c["type"] == "SW" else 1, c["m"], c["n"], c["g"], seqs, quals)
    assert got.decode() == c["consensus"]


@pytest.mark.parametrize("name", golden_names())
def test_oracle_vs_committed_reference_outputs(name):
    batch, pkw, want = load_golden(name)
    got = checker.oracle_polish(batch, make_params(**pkw), threads=4)
    assert_same(got, want, name)


@pytest.mark.skipif(not checker.have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("seed,kw,pkw", [
    (201, dict(n_windows=12), dict()),
    (202, dict(n_windows=12, partial=0.8), dict()),
    (203, dict(n_windows=12, fastq=False), dict()),
    (204, dict(n_windows=12, n_frac=0.05, null_qual=0.5), dict()),
    (205, dict(n_windows=12), dict(haplotype=0)),
    (206, dict(n_windows=12, partial=0.8), dict(haplotype=0, trim=0)),
    (207, dict(n_windows=4, depth=70, length=100), dict()),     # > 16 equal keys: std::sort scrambles (H4)
    (208, dict(n_windows=6, err=0.4), dict(num_prune=4)),
])
def test_oracle_vs_live_reference(seed, kw, pkw):
    batch = fuzz_batch(seed, **kw)
    p = make_params(**pkw)
    assert_same(checker.oracle_polish(batch, p, threads=4), checker.ref_polish(batch, p, threads=4), "seed %d" % seed)


@pytest.mark.skipif(not checker.have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("type_", [0, 1])
def test_alignment_pairs_vs_live_reference(type_):
    rng = np.random.default_rng(31 + type_)
    for _ in range(12):
        n = int(rng.integers(40, 160))
        truth = rng.choice([65, 67, 71, 84], size=n).astype(np.uint8)

        def noisy():
            s = [int(c) for c in truth if rng.random() > 0.08]
            for _ in range(int(rng.integers(0, 8))):
                s.insert(int(rng.integers(0, len(s))), int(rng.choice([65, 67, 71, 84])))
            return bytes(s)

        seqs = [noisy() for _ in range(int(rng.integers(1, 8)))]
        q = noisy()
        assert checker.oracle_align_probe(type_, 3, -5, -4, seqs, q) == checker.ref_align_probe(type_, 3, -5, -4, seqs, q)
