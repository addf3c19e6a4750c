"""The C-ABI library: loads, exports every symbol include/vgc.h declares, struct layouts agree with the ctypes
mirror, and — with no GPU — every compute entry point fails loudly (there is no CPU path)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from vechat_b200 import engine
from vechat_b200._ffi import VgcBatch, VgcParams, VgcResult, VgcStats, make_params
from vechat_b200.sim import fuzz_batch


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vgc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vgc_[a-z_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    lib = engine.load_library()
    names = declared_symbols()
    assert set(names) == set(engine.EXPORTS), (names, engine.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_sizes():
    # the sizes a C compiler gives the structs of include/vgc.h on LP64
    assert C.sizeof(VgcParams) == 32
    assert C.sizeof(VgcBatch) == 8 + 8 * 8
    assert C.sizeof(VgcResult) == 32
    assert C.sizeof(VgcStats) == 4 * 8 + 5 * 8 + 8 + 8  # + host_pack_ms


def test_weight_lut_formula():
    lut = engine.weight_lut()
    for q in (33, 34, 40, 43, 53, 73, 126):
        assert lut[q] == int((1 - 10 ** ((33 - q) / 10.0)) * 1000) or abs(lut[q] - (1 - 10 ** ((33 - q) / 10.0)) * 1000) < 1
    assert lut[33] == 0


def test_version_string():
    assert b"sm_100a" in engine.load_library().vgc_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_device_fails_loudly():
    with pytest.raises(engine.VgcError) as ei:
        engine.Engine(0)
    assert ei.value.code == 2  # VGC_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value) or "no usable" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must not import, link or dlopen anything under oracle/ (it would void every parity
    claim).  build.py may *build* the checker; nothing else may touch it."""
    pkg = os.path.join(ROOT, "vechat_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|libvechat_ref|oracle/|oracle_polish|ref_polish|#include\s+\"[^\"]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) and f != "build.py":
                for ln in open(os.path.join(dirpath, f)):
                    assert not pat.search(ln), (os.path.join(dirpath, f), ln)
