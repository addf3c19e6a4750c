#!/usr/bin/env python
"""Correlate an ncu SASS source page with source lines (nvdisasm --print-line-info) and aggregate per line.

usage: tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Prints: share of executed warp instructions and of stall samples per source line, plus per-function totals.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    dis = []
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        out = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True,
                             text=True).stdout
        if kname in out:
            dis = out.splitlines()
            break
    # offset -> (file, line) for the chosen kernel
    loc, cur, inside = {}, ("?", 0), False
    for ln in dis:
        if ln.startswith(".text."):
            inside = kname in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m:
            loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    base = int(rows[2][ia], 16)
    inst = collections.Counter()
    samp = collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        if len(r) <= isamp:
            continue
        off = int(r[ia], 16) - base
        key = loc.get(off, (("?", 0), ""))[0]
        inst[key] += int(r[ii] or 0)
        samp[key] += int(r[isamp] or 0)
        for h, i in stall_cols.items():
            v = int(r[i] or 0)
            if v:
                stalls[key][h] += v
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print("total warp instructions %d, samples %d" % (ti, ts))
    print("%-28s %8s %8s  top stalls" % ("file:line", "inst%", "samp%"))
    for key, _ in samp.most_common(top):
        st = ", ".join("%s %.0f%%" % (h[6:], 100.0 * v / samp[key]) for h, v in stalls[key].most_common(3))
        print("%-28s %8.2f %8.2f  %s" % ("%s:%d" % key, 100.0 * inst[key] / ti, 100.0 * samp[key] / ts, st))
    byfile = collections.Counter()
    for key, v in samp.items():
        byfile[key[0]] += v
    print("by file:", {k: round(100.0 * v / ts, 1) for k, v in byfile.most_common()})


if __name__ == "__main__":
    main()
