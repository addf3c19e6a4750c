"""The CPU-runnable half of bench.py's contract: the reference arm (`--impl reference`) prints ONE JSON line with the
keys the driver reads, and the GPU arm fails loudly without a device (no CPU fallback).  The GPU arm itself is measured
on the B200 box (profiles/r02_*.json)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--targets", "20"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "poa_windows_per_sec" and d["unit"] == "windows/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_gpu_arm_fails_loudly_without_a_device():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0", "--targets", "20"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert b"no CUDA device" in r.stderr or b"no CPU path" in r.stderr
    assert not [l for l in r.stdout.decode().splitlines() if l.startswith("{")]
