#!/usr/bin/env python
"""Config-5-shaped streaming job through the WHOLE program (VERDICT round 1 "Next" #4):
many batches of windows pulled from one queue by every listed device.

  1. synthesize the reads and their ground-truth all-vs-all overlaps (vechat_b200.sim: FASTQ + PAF files);
  2. run `vechat_racon_b200 -f -p -d 0.2 -s 0.2` on them with VECHAT_B200_DEVICES (POA on the listed GPUs, batches from
     one queue) and VECHAT_B200_ALIGN=1 (overlap alignment + breaking points on the GPUs as well — the host aligner
     would need ~1 h for 3 M overlaps);
  3. report the stage times the reference's own Logger prints (initialize() separately from polish()) and the
     windows/s of the polish stage;
  4. parity sample: the UNMODIFIED reference program (oracle/_ref/vechat_racon, CPU) corrects a handful of the same
     targets from the same files; its FASTA records must equal the B200 program's byte for byte.

usage: tools/stream_job.py [--reads 50000] [--devices 0,1] [--check 24] [--out gpurun_out/r02_stream.json]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "vechat_racon")
B200_BIN = os.path.join(ROOT, "vechat_b200", "lib", "vechat_racon_b200")
OPTS = ["-f", "-p", "-d", "0.2", "-s", "0.2"]


def stages(stderr):
    """Logger lines -> [(stage, seconds)]; a progress bar's last frame carries the stage's time."""
    out = []
    for line in stderr.replace("\r", "\n").split("\n"):
        m = re.match(r"\[racon::([A-Za-z0-9_:]+)\] (.*?) (?:\[[=> ]*\] )?([0-9.]+) s$", line.strip())
        if m:
            key = "%s %s" % (m.group(1), m.group(2).strip())
            if out and out[-1][0] == key:
                out[-1] = (key, float(m.group(3)))
            else:
                out.append((key, float(m.group(3))))
    return out


def fasta_records(path):
    recs, name = {}, None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                name = line[1:].split()[0]
                recs[name] = [line.rstrip(b"\n"), b""]
            elif name is not None:
                recs[name][1] += line.rstrip(b"\n")
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=50000)
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--check", type=int, default=24, help="targets the reference program re-corrects (0: skip)")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--dir", default="/tmp/vgc_stream")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_stream.json"))
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    from vechat_b200.sim import Simulator

    genome = int(args.reads * args.read_len / args.depth)
    t0 = time.perf_counter()
    sim = Simulator("pb_clr_10k_x_10kb", n_reads=args.reads, read_len=args.read_len, genome_len=genome)
    reads, paf = os.path.join(args.dir, "reads.fq"), os.path.join(args.dir, "overlaps.paf")
    n_ovl = sim.export(reads, paf)
    del sim
    t_gen = time.perf_counter() - t0
    n_windows = args.reads * ((args.read_len + 499) // 500)
    print("synthesized %d reads x %d b, %d overlaps in %.1f s (%.2f GB FASTQ)" % (
        args.reads, args.read_len, n_ovl, t_gen, os.path.getsize(reads) / 1e9), flush=True)

    env = dict(os.environ, VECHAT_B200_DEVICES=args.devices, VECHAT_B200_ALIGN="1")
    out_fa = os.path.join(args.dir, "corrected.b200.fa")
    t0 = time.perf_counter()
    with open(out_fa, "wb") as fo:
        r = subprocess.run([B200_BIN] + OPTS + ["-t", str(args.threads), reads, paf, reads], env=env, stdout=fo,
                           stderr=subprocess.PIPE)
    wall = time.perf_counter() - t0
    err = r.stderr.decode(errors="replace")
    if r.returncode != 0:
        print(err[-3000:])
        raise SystemExit("vechat_racon_b200 failed with %d" % r.returncode)
    st = stages(err)
    polish_s = sum(s for k, s in st if k.startswith("Polisher::polish"))
    total_s = sum(s for k, s in st if "total" in k)
    init_s = total_s - polish_s  # everything before polish(): parsing, overlap alignment, tiling
    queue = [l for l in err.split("\n") if "batches from one queue" in l]
    got = fasta_records(out_fa)
    summary = {
        "job": "config-5 shape, streamed: %d reads x %d b at %.0fx, %d overlaps (PAF), %d windows of 500 b, "
               "haplotype mode (-f -p -d 0.2 -s 0.2)" % (args.reads, args.read_len, args.depth, n_ovl, n_windows),
        "program": "vechat_racon_b200, VECHAT_B200_DEVICES=%s, VECHAT_B200_ALIGN=1, -t %d" % (args.devices, args.threads),
        "devices": args.devices, "windows": n_windows, "overlaps": n_ovl, "corrected_reads": len(got),
        "corrected_bases": sum(len(v[1]) for v in got.values()),
        "wall_s": wall, "synthesis_s": t_gen,
        "program_total_s": total_s, "initialize_s": init_s, "polish_s": polish_s,
        "polish_windows_per_s": n_windows / polish_s if polish_s else None,
        "whole_program_windows_per_s": n_windows / wall,
        "stages": [{"stage": k, "seconds": s} for k, s in st],
        "queue": queue[0].strip() if queue else None,
    }
    # VGC_VERBOSE=1 / VECHAT_B200_VERBOSE=1: per-batch engine and binding timing
    trace = [l.strip() for l in err.split("\n") if l.startswith("[vgc]") or l.startswith("[racon::B200Polisher::polish] ")]
    if trace:
        summary["engine_trace"] = trace[:64]
    print(json.dumps(summary, indent=1), flush=True)

    if args.check and os.path.exists(REF_BIN):
        # the reference corrects a handful of the same targets (an overlap whose target is not in the target file is
        # dropped by Overlap::transmute, src/overlap.cpp:160-164, so the same PAF serves)
        ids = sorted(set(int(i * (args.reads - 1) / max(1, args.check - 1)) for i in range(args.check)))
        want_names = set(b"read%d" % i for i in ids)
        sub = os.path.join(args.dir, "targets_sub.fq")
        with open(reads, "rb") as f, open(sub, "wb") as fo:
            while True:
                rec = [f.readline() for _ in range(4)]
                if not rec[0]:
                    break
                if rec[0][1:].strip() in want_names:
                    fo.writelines(rec)
        ref_env = dict(os.environ)
        ref_env.pop("VECHAT_B200_DEVICES", None)
        ref_env.pop("VECHAT_B200_ALIGN", None)
        ref_fa = os.path.join(args.dir, "corrected.ref.fa")
        t0 = time.perf_counter()
        with open(ref_fa, "wb") as fo:
            rr = subprocess.run([REF_BIN] + OPTS + ["-t", str(args.threads), reads, paf, sub], env=ref_env, stdout=fo,
                                stderr=subprocess.PIPE)
        ref_wall = time.perf_counter() - t0
        if rr.returncode != 0:
            print(rr.stderr.decode(errors="replace")[-2000:])
            raise SystemExit("reference program failed with %d" % rr.returncode)
        want = fasta_records(ref_fa)
        bad = [k.decode() for k, v in want.items() if got.get(k) != v]
        ref_windows = sum((len(v[1]) + 499) // 500 for v in want.values())
        ref_st = stages(rr.stderr.decode(errors="replace"))
        summary["parity"] = {"targets_checked": len(want), "mismatches": len(bad), "against": "reference program "
                             "(oracle/_ref/vechat_racon, CPU) on the same files", "reference_wall_s": ref_wall,
                             "reference_polish_s": sum(s for k, s in ref_st if k.startswith("Polisher::polish")),
                             "reference_windows": ref_windows}
        print(json.dumps(summary["parity"]), flush=True)
        if bad:
            print("MISMATCH:", bad[:10])
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    return 1 if summary.get("parity", {}).get("mismatches") else 0


if __name__ == "__main__":
    sys.exit(main())
