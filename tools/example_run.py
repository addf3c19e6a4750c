"""Config 1 on the box: run the reference program and the B200 program on the same example inputs, compare the FASTA
byte for byte, and report the stage times both print through the reference's own Logger (stderr).

    python tools/example_run.py [--dir oracle/_ref/example | tests/golden/example] [--skip-ref] [--out gpurun_out/example.json]

`--skip-ref` compares against the committed / pre-generated reference output instead of re-running the CPU program
(the whole example costs minutes of host time).  Evidence tool, not the bench.
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "vechat_racon")
B200_BIN = os.path.join(ROOT, "vechat_b200", "lib", "vechat_racon_b200")


def stages(stderr):
    """last progress report of every Logger message -> seconds"""
    out = {}
    for piece in re.split(r"[\r\n]", stderr):
        m = re.match(r"\[racon::([\w:]*)\] (.*?)(?: \[[=> ]*\])? ([0-9.]+) s$", piece.strip())
        if m:
            out[m.group(1).split("::")[-1] + ": " + m.group(2)] = float(m.group(3))
    return out


def run(binary, args, cwd, devices=None, gpu_align=False):
    env = dict(os.environ)
    env.pop("VECHAT_B200_DEVICES", None)
    env.pop("VECHAT_B200_ALIGN", None)
    if devices:
        env["VECHAT_B200_DEVICES"] = devices
    if gpu_align:
        env["VECHAT_B200_ALIGN"] = gpu_align if isinstance(gpu_align, str) else "1"
    t0 = time.time()
    r = subprocess.run([binary] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    wall = time.time() - t0
    if r.returncode != 0:
        sys.exit("%s failed: %s" % (binary, r.stderr.decode()[-500:]))
    return r.stdout, stages(r.stderr.decode()), wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(ROOT, "oracle", "_ref", "example"))
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--gpu-align", nargs="?", const="1", default=None,
                    help="VECHAT_B200_ALIGN: 1 = align + cut on the GPU (vga_break), cigar = CIGARs only (vga_align)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    targets = "targets.fq.gz" if os.path.exists(os.path.join(a.dir, "targets.fq.gz")) else "reads.fq.gz"
    paf = "overlaps.paf.gz" if os.path.exists(os.path.join(a.dir, "overlaps.paf.gz")) else "overlaps.paf"
    args = ["-f", "-p", "-d", "0.2", "-s", "0.2", "-t", str(a.threads), "reads.fq.gz", paf, targets]
    rep = {"dir": os.path.relpath(a.dir, ROOT), "args": " ".join(args), "host_threads": a.threads}
    got, st, wall = run(B200_BIN, args, a.dir, a.devices, a.gpu_align)
    rep["gpu_align"] = a.gpu_align or False
    rep["b200"] = {"wall_s": round(wall, 2), "stages_s": st, "fasta_sha256": hashlib.sha256(got).hexdigest(),
                   "reads_out": got.count(b">"), "bases_out": sum(len(l) for l in got.split(b"\n") if not l.startswith(b">"))}
    if a.skip_ref:
        for name in ("corrected.ref.fa", "corrected.hap.fa"):
            p = os.path.join(a.dir, name)
            if os.path.exists(p):
                want = open(p, "rb").read()
                rep["reference"] = {"from_file": name, "fasta_sha256": hashlib.sha256(want).hexdigest()}
                break
        else:
            sys.exit("no stored reference output in " + a.dir)
    else:
        want, st, wall = run(REF_BIN, args, a.dir)
        rep["reference"] = {"wall_s": round(wall, 2), "stages_s": st, "fasta_sha256": hashlib.sha256(want).hexdigest()}
    rep["identical"] = bool(got == want)
    s = json.dumps(rep, indent=1)
    print(s)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(s + "\n")
    sys.exit(0 if rep["identical"] else 1)


if __name__ == "__main__":
    main()
