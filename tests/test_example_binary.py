"""Config 1 plumbing: the reference PROGRAM with the B200 engine dropped in.

* oracle/_ref/vechat_racon        — the unmodified reference (main.cpp, Polisher::initialize/polish, window.cpp, spoa)
                                    compiled by oracle/Makefile `racon`; the authority here.
* vechat_b200/lib/vechat_racon_b200 — the same objects + csrc/racon_binding/b200polisher.cpp (racon::B200Polisher,
                                    overrides polish() only) + libvgc.so.

Both read the same reads / overlaps / targets, so the window tilings are identical by construction (initialize() is
inherited) and the corrected FASTA must be identical byte for byte, headers (LN/RC/XC tags) included.

Fixture: tests/golden/example/ = a cluster of 10 neighbouring reads of the reference's example/reads.fq.gz as targets,
the 89 reads overlapping them, overlaps from tools/example_overlaps.py (minimap2 is not in this image), and the
reference binary's outputs (corrected.hap.fa: `-f -p -d 0.2 -s 0.2`, the vechat driver's first pass,
scripts/vechat:69-72; corrected.lin.fa: `-f`, its second pass, :90-93).  One of the 10 targets has no polished
window and is dropped by both (polisher.cpp:529).
"""
import ctypes as C
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

EX = os.path.join(GOLDEN, "example")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "vechat_racon")
B200_BIN = os.path.join(ROOT, "vechat_b200", "lib", "vechat_racon_b200")
HAP = ["-f", "-p", "-d", "0.2", "-s", "0.2"]
LIN = ["-f"]

need_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/vechat_racon not built (needs /root/reference)")
need_b200 = pytest.mark.skipif(not os.path.exists(B200_BIN), reason="vechat_racon_b200 not built (needs /root/reference)")


def run(binary, opts, reads="reads.fq.gz", paf="overlaps.paf", targets="targets.fq.gz", cwd=EX, devices=None, threads=8,
        gpu_align=False):
    env = dict(os.environ)
    env.pop("VECHAT_B200_DEVICES", None)
    env.pop("VECHAT_B200_ALIGN", None)
    if devices is not None:
        env["VECHAT_B200_DEVICES"] = devices
    if gpu_align:
        env["VECHAT_B200_ALIGN"] = gpu_align if isinstance(gpu_align, str) else "1"
    return subprocess.run([binary] + opts + ["-t", str(threads), reads, paf, targets], cwd=cwd, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)


def golden(name):
    with open(os.path.join(EX, name), "rb") as f:
        return f.read()


# ---------------------------------------------------------------- CPU: the authority and the seam

@need_ref
@pytest.mark.parametrize("opts,want", [(HAP, "corrected.hap.fa"), (LIN, "corrected.lin.fa")])
def test_reference_binary_reproduces_committed_fasta(opts, want):
    r = run(REF_BIN, opts)
    assert r.returncode == 0, r.stderr[-400:]
    assert r.stdout == golden(want)
    assert r.stdout.count(b">") == 9  # 10 targets, one without a polished window


@need_ref
def test_reference_binary_thread_count_is_not_a_parity_variable():
    assert run(REF_BIN, HAP, threads=1).stdout == golden("corrected.hap.fa")


@need_b200
def test_b200_binary_without_devices_is_the_reference():
    """No VECHAT_B200_DEVICES and no -c: createPolisherB200 forwards to the reference's own factory."""
    r = run(B200_BIN, HAP)
    assert r.returncode == 0
    assert r.stdout == golden("corrected.hap.fa")


@need_b200
def test_b200_binary_fails_loudly_without_gpu():
    """The engine has no CPU path: asking for a device that is not there ends like the reference's error sites
    (stderr message + exit(1)), after initialize() and before any output."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(B200_BIN, HAP, devices="0")
    assert r.returncode == 1
    assert r.stdout == b""
    assert b"[racon::B200Polisher::polish] error:" in r.stderr and b"no usable CUDA device" in r.stderr


@need_b200
def test_b200_binary_rejects_bad_device_list():
    r = run(B200_BIN, HAP, devices="zero")
    assert r.returncode == 1 and b"VECHAT_B200_DEVICES" in r.stderr


@need_b200
def test_b200_binary_gpu_alignment_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(B200_BIN, HAP, devices="0", gpu_align=True)
    assert r.returncode == 1 and r.stdout == b""
    assert b"[racon::B200Polisher::find_overlap_breaking_points] error:" in r.stderr


# ---------------------------------------------------------------- CPU: the binding's host side behind a mock engine

def _mock_env():
    """LD_PRELOAD shim (tests/host_model/mock_vgc.cpp, test-only) that answers vgc_polish with the checker, so the
    REAL vechat_racon_b200 binary runs here and its device ranges / batching / packing / store / stitch are checked
    without a GPU."""
    src = os.path.join(ROOT, "tests", "host_model", "mock_vgc.cpp")
    out = os.path.join(ROOT, "tests", "host_model", "_build", "libmockvgc.so")
    core = os.path.join(ROOT, "vechat_b200", "csrc", "ovl_core.h")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "vechat_b200", "csrc"), "-o", out, src, "-ldl"], check=True)
    return {"LD_PRELOAD": out, "MOCK_VGC_REF_SO": os.path.join(ROOT, "oracle", "_ref", "libvechat_ref.so")}


def _run_mock(opts, devices, batch_windows, cwd=EX, targets="targets.fq.gz", align=None):
    env = dict(os.environ, VECHAT_B200_DEVICES=devices, VECHAT_B200_BATCH_WINDOWS=str(batch_windows), **_mock_env())
    env.pop("VECHAT_B200_ALIGN", None)
    if align:
        env["VECHAT_B200_ALIGN"] = align
    return subprocess.run([B200_BIN] + opts + ["-t", "8", "reads.fq.gz", "overlaps.paf", targets], cwd=cwd, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)


@need_b200
@need_ref
@pytest.mark.parametrize("opts,want,devices,batch", [
    (HAP, "corrected.hap.fa", "0", 1 << 16), (HAP, "corrected.hap.fa", "0,1", 40), (LIN, "corrected.lin.fa", "0,1,2", 7),
    (HAP, "corrected.hap.fa", "0,1,2,3,4,5,6,7,8,9,10,11", 1000),  # more devices than targets: empty ranges
])
def test_binding_host_side_behind_mock_engine(opts, want, devices, batch):
    r = _run_mock(opts, devices, batch)
    assert r.returncode == 0, r.stderr[-400:]
    assert r.stdout == golden(want)
    calls = [l for l in r.stderr.decode().split("\n") if l.startswith("[mock_vgc]")]
    windows = [int(l.split(":")[1].split()[0]) for l in calls]
    assert sum(windows) == 168 and max(windows) <= batch  # every window exactly once, batches respected
    nd, used = len(devices.split(",")), len({l.split()[4] for l in calls})
    assert used == nd if nd <= 3 else 1 <= used <= 10  # contiguous ranges of whole targets (10 targets here)


@need_b200
@need_ref
@pytest.mark.parametrize("align,call,devices", [("1", "break", "0"), ("cigar", "align", "0"), ("1", "break", "0,1,2")])
@pytest.mark.parametrize("opts,want", [(HAP, "corrected.hap.fa"), (LIN + ["-w", "300"], None)])
def test_binding_aligner_glue_behind_mock(align, call, devices, opts, want):
    """VECHAT_B200_ALIGN=1 / =cigar through the binding's CUDABatchAligner (substring offsets, strands, the hand-over
    of breaking_points_ / cigar_) with the aligner core run on the host by the mock: FASTA = the reference program's,
    i.e. the tilings did not move."""
    r = _run_mock(opts, devices, 1 << 16, align=align)
    assert r.returncode == 0, r.stderr[-400:]
    calls = [l.split() for l in r.stderr.decode().split("\n") if l.startswith("[mock_vga] %s:" % call)]
    assert len(calls) == len(devices.split(",")) and sum(int(c[2]) for c in calls) == 422  # one range per device
    assert r.stdout == (golden(want) if want else run(REF_BIN, opts).stdout)


def _write_sam(path):
    """The fixture's overlaps as SAM records whose CIGARs are the host aligner's (soft clips for the unaligned ends;
    for a reverse-strand record the CIGAR runs along the reverse complement, as SAM has it)."""
    from oracle import checker
    reads = {}
    with gzip.open(os.path.join(EX, "reads.fq.gz"), "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        reads[lines[i][1:].decode()] = lines[i + 1]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    with open(path, "w") as out:
        out.write("@HD\tVN:1.6\n")
        for line in open(os.path.join(EX, "overlaps.paf")):
            f = line.split("\t")
            q, t = reads[f[0]], reads[f[5]]
            ql, qb, qe, tb, te = int(f[1]), int(f[2]), int(f[3]), int(f[7]), int(f[8])
            rev = f[4] == "-"
            qs = q.translate(comp)[::-1][ql - qe:ql - qb] if rev else q[qb:qe]
            cigar, _ = checker.standin_cigar(qs, t[tb:te])
            left, right = (ql - qe, qb) if rev else (qb, ql - qe)
            cigar = ("%dS" % left if left else "") + cigar + ("%dS" % right if right else "")
            out.write("\t".join([f[0], "16" if rev else "0", f[5], str(tb + 1), "255", cigar, "*", "0", "0", "*", "*"]) + "\n")


@need_b200
@need_ref
def test_sam_input_carries_its_own_alignments(tmp_path):
    """SAM overlaps come with CIGARs: the reference skips edlib for them (overlap.cpp:191) and the binding's GPU
    aligner must skip them too.  With the host aligner's CIGARs in the file the tilings equal the PAF run's, so the
    FASTA equals the committed one — which also pins the I / D orientation of the aligner against the reference's
    own CIGAR reading (overlap.cpp:71-93, 240-290)."""
    sam = str(tmp_path / "overlaps.sam")
    _write_sam(sam)
    r = run(REF_BIN, HAP, paf=sam)
    assert r.returncode == 0, r.stderr[-400:]
    assert r.stdout == golden("corrected.hap.fa")
    env = dict(os.environ, VECHAT_B200_DEVICES="0", VECHAT_B200_ALIGN="1", **_mock_env())
    m = subprocess.run([B200_BIN] + HAP + ["-t", "8", "reads.fq.gz", sam, "targets.fq.gz"], cwd=EX, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert m.returncode == 0, m.stderr[-400:]
    assert m.stdout == golden("corrected.hap.fa")
    assert b"[mock_vga]" not in m.stderr  # nothing left to align


@need_b200
@need_ref
def test_binding_host_side_300_targets_behind_mock_engine():
    """5 600 windows: the threaded packer, several batches per device, three device ranges."""
    d = os.path.join(ROOT, "oracle", "_ref", "example_300")
    want = os.path.join(d, "corrected.ref.fa")
    if not os.path.exists(want):
        pytest.skip("oracle/_ref/example_300 not generated (tools/example_overlaps.py --targets 300)")
    r = _run_mock(HAP, "0,1,2", 700, cwd=d)
    assert r.returncode == 0, r.stderr[-400:]
    with open(want, "rb") as f:
        assert r.stdout == f.read()


# ---------------------------------------------------------------- the edlib stand-in both binaries are built on

def _edlib():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libedlib_standin.so"))

    class Cfg(C.Structure):
        _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("eq", C.c_void_p), ("neq", C.c_int)]

    class Res(C.Structure):
        _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)),
                    ("startLocations", C.POINTER(C.c_int)), ("numLocations", C.c_int),
                    ("alignment", C.POINTER(C.c_ubyte)), ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]

    lib.edlibNewAlignConfig.restype = Cfg
    lib.edlibNewAlignConfig.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.edlibAlign.restype = Res
    lib.edlibAlign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, Cfg]
    lib.edlibAlignmentToCigar.restype = C.c_void_p
    lib.edlibAlignmentToCigar.argtypes = [C.POINTER(C.c_ubyte), C.c_int, C.c_int]
    lib.edlibFreeAlignResult.argtypes = [Res]
    return lib


def _edit_distance(a, b):
    prev = np.arange(len(b) + 1)
    for i, ca in enumerate(a, 1):
        cur = np.empty_like(prev)
        cur[0] = i
        sub = prev[:-1] + (np.frombuffer(b, np.uint8) != ca)
        best = np.minimum(sub, prev[1:] + 1)
        # horizontal moves: cur[j] = min(best[j], cur[j-1] + 1)  -> prefix scan on (value - index)
        row = np.concatenate(([i], best))
        row = np.minimum.accumulate(row - np.arange(len(row))) + np.arange(len(row))
        cur = row
        prev = cur
    return int(prev[-1])


def test_edlib_standin_is_an_exact_global_aligner():
    """edit distance == textbook DP; the path consumes both strings, its edits == the distance, M columns agree
    with match/mismatch, and the standard CIGAR is the run-length form overlap.cpp:225-292 parses."""
    lib = _edlib()
    rng = np.random.default_rng(7)
    cases = [(b"", b"ACGT"), (b"ACGT", b""), (b"A", b"A"), (b"A", b"C"), (b"ACGT" * 5, b"ACGT" * 5)]
    for _ in range(60):
        n = int(rng.integers(1, 400))
        t = rng.integers(0, 4, n)
        q = []
        for x in t:  # noisy copy: 10 % sub, 8 % ins, 6 % del
            u = rng.random()
            if u < 0.06:
                continue
            q.append(int(rng.integers(0, 4)) if u < 0.16 else int(x))
            if rng.random() < 0.08:
                q.append(int(rng.integers(0, 4)))
        cases.append((bytes(b"ACGT"[x] for x in q), bytes(b"ACGT"[x] for x in t)))
    for q, t in cases:
        cfg = lib.edlibNewAlignConfig(-1, 0, 2, None, 0)  # EDLIB_MODE_NW, EDLIB_TASK_PATH
        r = lib.edlibAlign(q, len(q), t, len(t), cfg)
        assert r.status == 0
        assert r.editDistance == _edit_distance(q, t), (q, t)
        ops = [r.alignment[i] for i in range(r.alignmentLength)]
        i = j = edits = 0
        for op in ops:
            if op in (0, 3):
                assert (q[i] == t[j]) == (op == 0)
                i, j, edits = i + 1, j + 1, edits + (op == 3)
            elif op == 1:
                i, edits = i + 1, edits + 1
            else:
                j, edits = j + 1, edits + 1
        assert (i, j, edits) == (len(q), len(t), r.editDistance)
        p = lib.edlibAlignmentToCigar(r.alignment, r.alignmentLength, 0)
        cigar = C.string_at(p).decode()
        C.CDLL(None).free(C.c_void_p(p))
        runs, k = [], 0
        while k < len(cigar):
            m = k
            while cigar[m].isdigit():
                m += 1
            runs.append((int(cigar[k:m]), cigar[m]))
            k = m + 1
        assert sum(n for n, c in runs if c in "MI") == len(q) and sum(n for n, c in runs if c in "MD") == len(t)
        assert all(a[1] != b[1] for a, b in zip(runs, runs[1:]))
        lib.edlibFreeAlignResult(r)


# ---------------------------------------------------------------- GPU: byte-for-byte FASTA parity

def _b200(opts, **kw):
    assert os.path.exists(B200_BIN), "vechat_racon_b200 is missing: build it in the development container " \
                                     "(python -m vechat_b200.build) so that it travels with the snapshot"
    r = run(B200_BIN, opts, devices="0", **kw)
    assert r.returncode == 0, r.stderr[-600:]
    assert b"[racon::createPolisherB200] consensus on B200 device(s)" in r.stderr  # the GPU polisher ran, not the CPU one
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("opts,want", [(HAP, "corrected.hap.fa"), (LIN, "corrected.lin.fa")])
def test_gpu_binary_matches_committed_reference_fasta(opts, want):
    assert _b200(opts) == golden(want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "cigar"])
@pytest.mark.parametrize("opts,want", [(HAP, "corrected.hap.fa"), (LIN, "corrected.lin.fa")])
def test_gpu_binary_with_gpu_overlap_alignment(opts, want, mode):
    """VECHAT_B200_ALIGN=1: overlaps aligned and cut into breaking points on the GPU (vga_break); =cigar: only the
    CIGARs come from the GPU (vga_align).  The tilings, hence the FASTA, stay identical because the GPU aligner
    reproduces the host aligner's alignments exactly."""
    r = run(B200_BIN, opts, devices="0", gpu_align=mode)
    assert r.returncode == 0, r.stderr[-600:]
    assert b"aligned overlaps on the GPU" in r.stderr and b"[racon::createPolisherB200] consensus on B200 device(s)" in r.stderr
    assert r.stdout == golden(want)


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [
    HAP + ["-u"],                               # keep unpolished targets (drop_unpolished_sequences = false)
    HAP + ["-k", "1"], HAP + ["-k", "2", "-d", "0.3", "-s", "0.1"],
    HAP + ["-w", "300"], HAP + ["-w", "640"],   # other window lengths (tilings change with them)
    HAP + ["-q", "-1"],                         # no quality filter: every layer kept, deepest windows
    HAP + ["-m", "5", "-x", "-4", "-g", "-8"],  # racon's own default scores
    LIN + ["--no-trimming"], LIN + ["-u", "-w", "300"],
], ids=lambda o: " ".join(o))
def test_gpu_binary_matches_reference_binary_live(opts):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/vechat_racon not built")
    want = run(REF_BIN, opts)
    assert want.returncode == 0
    assert _b200(opts) == want.stdout


@pytest.mark.gpu
def test_gpu_binary_two_handles(monkeypatch):
    """The multi-device path of B200Polisher::polish: one host thread + one vgc_handle per listed device, each over
    a contiguous range of whole targets.  Listing device 0 twice runs it on a single-GPU box (two handles on one
    device, concurrent vgc_polish calls: the library keeps no global state); a second GPU is used when present."""
    import torch
    monkeypatch.setenv("VGC_MEM_BUDGET_MB", "12000")  # two engines share one device here
    devices = "0,1" if torch.cuda.device_count() > 1 else "0,0"
    r = run(B200_BIN, HAP, devices=devices)
    assert r.returncode == 0, r.stderr[-600:]
    assert r.stdout == golden("corrected.hap.fa")
    r = run(B200_BIN, LIN + ["-u"], devices="0,0,0")
    want = run(REF_BIN, LIN + ["-u"]) if os.path.exists(REF_BIN) else None
    assert r.returncode == 0, r.stderr[-600:]
    if want is not None:
        assert r.stdout == want.stdout and r.stdout.count(b">") == 10


@pytest.mark.gpu
def test_gpu_binary_fasta_input(tmp_path):
    """Second-pass shape (scripts/vechat:372-397): reads and targets are FASTA, every window takes the dummy-quality
    branch of window.cpp:223 except each target's last, shorter window."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/vechat_racon not built")
    for src, dst in (("reads.fq.gz", "reads.fa"), ("targets.fq.gz", "targets.fa")):
        with gzip.open(os.path.join(EX, src), "rt") as f, open(tmp_path / dst, "w") as g:
            lines = f.read().split("\n")
            for i in range(0, len(lines) - 3, 4):
                g.write(">" + lines[i][1:] + "\n" + lines[i + 1] + "\n")
    paf = os.path.join(EX, "overlaps.paf")
    for opts in (HAP, LIN):
        want = run(REF_BIN, opts, "reads.fa", paf, "targets.fa", cwd=str(tmp_path))
        got = run(B200_BIN, opts, "reads.fa", paf, "targets.fa", cwd=str(tmp_path), devices="0")
        assert want.returncode == 0 and got.returncode == 0, got.stderr[-400:]
        assert got.stdout == want.stdout and got.stdout.count(b">") >= 9


@pytest.mark.gpu
@pytest.mark.skipif(not os.environ.get("VGC_FULL_EXAMPLE"), reason="opt-in (VGC_FULL_EXAMPLE=1): whole example, minutes")
def test_gpu_binary_whole_example():
    """All 2 802 reads of example/reads.fq.gz against themselves (oracle/_ref/example/, generated by
    tools/example_overlaps.py; git-ignored, travels with the snapshot)."""
    d = os.path.join(ROOT, "oracle", "_ref", "example")
    want = os.path.join(d, "corrected.ref.fa")
    if not os.path.exists(want):
        pytest.skip("oracle/_ref/example not generated")
    got = run(B200_BIN, HAP, cwd=d, devices="0", threads=os.cpu_count() or 8)
    assert got.returncode == 0, got.stderr[-400:]
    with open(want, "rb") as f:
        assert got.stdout == f.read()


# ---------------------------------------------------------------- the streaming job tool (config 5's shape, tiny)

@need_b200
@need_ref
def test_stream_job_tool_behind_mock_engine(tmp_path):
    """tools/stream_job.py end to end on the CPU box: simulator export (FASTQ + ground-truth PAF) -> the real
    vechat_racon_b200 binary (two 'devices' pulling batches from one queue, mock engine) -> parity sample against the
    reference program on the same files."""
    import json
    out = tmp_path / "stream.json"
    env = dict(os.environ, VECHAT_B200_BATCH_WINDOWS="60", **_mock_env())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stream_job.py"), "--reads", "48", "--read-len", "3000",
                        "--devices", "0,1", "--check", "5", "--threads", "8", "--dir", str(tmp_path), "--out", str(out)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert r.returncode == 0, (r.stdout[-600:], r.stderr[-600:])
    s = json.loads(out.read_text())
    assert s["corrected_reads"] == 48 and s["windows"] == 48 * 6
    assert s["parity"]["targets_checked"] == 5 and s["parity"]["mismatches"] == 0
    assert "batches from one queue" in s["queue"] and "device 1 took" in s["queue"]


# ---------------------------------------------------------------- f-2: the binding's own tiling (host threads)

@need_b200
@need_ref
@pytest.mark.parametrize("opts", [HAP, LIN, HAP + ["-q", "12"], HAP + ["-w", "300"], HAP + ["-q", "11.5", "-w", "777"],
                                  LIN + ["--no-trimming"]])
def test_binding_tiling_equals_the_reference_tiling(opts):
    """B200Polisher::build_tiles (layers per window built on host threads inside find_overlap_breaking_points, the
    reference's serial loop of polisher.cpp:408-462 left with nothing to add) against the reference's own tiling
    (VECHAT_B200_TILING=0: layers read from Window::sequences_): same FASTA, byte for byte, also where the length and
    mean-quality filters bite (-q, -w); and both equal the reference program."""
    outs = {}
    for tiling in ("1", "0"):
        env = dict(os.environ, VECHAT_B200_DEVICES="0,1", VECHAT_B200_BATCH_WINDOWS="50", VECHAT_B200_TILING=tiling,
                   **_mock_env())
        env.pop("VECHAT_B200_ALIGN", None)
        r = subprocess.run([B200_BIN] + opts + ["-t", "8", "reads.fq.gz", "overlaps.paf", "targets.fq.gz"], cwd=EX, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        assert r.returncode == 0, r.stderr[-400:]
        assert (b"tiled the windows on host threads" in r.stderr) == (tiling == "1")
        outs[tiling] = r.stdout
    assert outs["1"] == outs["0"] and len(outs["1"]) > 1000
    ref = run(REF_BIN, opts)
    assert ref.returncode == 0 and ref.stdout == outs["1"]


@need_b200
@need_ref
def test_binding_tiling_fasta_and_contig_mode(tmp_path):
    """The tiles on inputs without qualities (FASTA reads: the mean-quality filter is skipped, layers carry no quality
    pointer, window.cpp:223's dummy-quality branch) and in contig mode (no -f: PolisherType::kC keeps one overlap per
    read, polisher.cpp:292-316) — binding tiles == reference tiling == reference program."""
    for src, dst in (("reads.fq.gz", "reads.fa"), ("targets.fq.gz", "targets.fa")):
        with gzip.open(os.path.join(EX, src), "rt") as f, open(tmp_path / dst, "w") as g:
            lines = f.read().split("\n")
            for i in range(0, len(lines) - 3, 4):
                g.write(">" + lines[i][1:] + "\n" + lines[i + 1] + "\n")
    paf = os.path.join(EX, "overlaps.paf")
    cases = [(HAP, "reads.fa", "targets.fa", str(tmp_path)), (LIN, "reads.fa", "targets.fa", str(tmp_path)),
             (["-p", "-d", "0.2", "-s", "0.2"], "reads.fq.gz", "targets.fq.gz", EX), ([], "reads.fq.gz", "targets.fq.gz", EX)]
    for opts, reads, targets, cwd in cases:
        want = run(REF_BIN, opts, reads, paf, targets, cwd=cwd)
        assert want.returncode == 0, want.stderr[-400:]
        for tiling in ("1", "0"):
            env = dict(os.environ, VECHAT_B200_DEVICES="0", VECHAT_B200_BATCH_WINDOWS="70", VECHAT_B200_TILING=tiling,
                       **_mock_env())
            env.pop("VECHAT_B200_ALIGN", None)
            r = subprocess.run([B200_BIN] + opts + ["-t", "8", reads, paf, targets], cwd=cwd, env=env,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
            assert r.returncode == 0, r.stderr[-400:]
            assert r.stdout == want.stdout, (opts, reads, tiling)
