"""pytest configuration: the `gpu` marker, the in-tree builds the tests need, and golden-fixture helpers."""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what is missing (in-tree, idempotent).  libvgc.so is compiled by nvcc — no GPU needed for that."""
    from vechat_b200 import build
    build.build_sim()
    build.build_engine()
    build.build_oracle()
    yield


def golden_names():
    return sorted(os.path.basename(p)[len("windows_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "windows_*.npz")))


def load_golden(name):
    """-> (WindowBatch, params kwargs, PolishResult of the reference)."""
    from vechat_b200._ffi import PolishResult, WindowBatch
    z = np.load(os.path.join(GOLDEN, "windows_%s.npz" % name))
    batch = WindowBatch(z["bases"], z["quals"], z["seq_off"], z["has_qual"], z["begin"], z["end"], z["win_first"],
                        z["win_flags"])
    want = PolishResult(z["cons"], z["cons_off"], z["polished"])
    return batch, json.loads(str(z["params"])), want


def assert_same(got, want, label=""):
    assert len(got.polished) == len(want.polished), label
    bad = [w for w in range(len(want.polished))
           if got.window(w) != want.window(w) or int(got.polished[w]) != int(want.polished[w])]
    assert not bad, "%s: %d/%d windows differ (first %d: got %r want %r)" % (
        label, len(bad), len(want.polished), bad[0], got.window(bad[0])[:60], want.window(bad[0])[:60])
