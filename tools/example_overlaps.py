"""Config 1 plumbing: overlaps for the reference's example/reads.fq.gz without minimap2.

minimap2 / fpa are not in this image, so the all-vs-all overlap file the `vechat` driver would produce
(scripts/vechat:37, `minimap2 -x ava-pb --dual=yes | awk '$11>=500' | fpa drop --same-name --internalmatch`) is
replaced by overlaps derived from placing every read on example/ref.fa with exact 15-mer matches (numpy only):
two reads overlap when their placements on the same reference sequence share >= MIN_OVL bases; coordinates are
interpolated from each read's linear fit.  The result is a plain 12-column PAF, both directions.  It is an INPUT
fixture: the reference binary and the B200 binary read the same file, which is what the parity contract fixes
("identical overlaps and window tilings").

    python tools/example_overlaps.py --out-dir oracle/_ref/example            # whole example (git-ignored, travels)
    python tools/example_overlaps.py --out-dir tests/golden/example --targets 10   # small committed fixture

Reads /root/reference/example at generation time only; tests and bench read the generated files.
"""
import argparse
import gzip
import os

import numpy as np

K = 15
MIN_HITS = 25
MIN_OVL = 1000
_CODE = np.full(256, 0, np.int64)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i
_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def read_fasta(path):
    out, name, chunks = [], None, []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, b"".join(chunks)))
                name, chunks = line[1:].split()[0].decode(), []
            else:
                chunks.append(line.strip())
    out.append((name, b"".join(chunks)))
    return out


def read_fastq(path):
    out = []
    with gzip.open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\n")
            f.readline()
            q = f.readline().rstrip(b"\n")
            out.append((h[1:].split()[0].decode(), s, q))
    return out


def kmers(seq):
    a = _CODE[np.frombuffer(seq, np.uint8)]
    n = len(a) - K + 1
    if n <= 0:
        return np.zeros(0, np.int64)
    c = np.zeros(n, np.int64)
    for j in range(K):
        c = (c << 2) | a[j:j + n]
    return c


class RefIndex:
    def __init__(self, seq):
        c = kmers(seq)
        order = np.argsort(c, kind="stable")
        cs = c[order]
        uniq = np.ones(len(cs), bool)
        uniq[1:] &= cs[1:] != cs[:-1]
        uniq[:-1] &= cs[:-1] != cs[1:]
        self.codes, self.pos, self.length = cs[uniq], order[uniq], len(seq)

    def place(self, seq):
        """(hits, a, s): reference position ~ a + s * read position, or None."""
        c = kmers(seq)
        i = np.searchsorted(self.codes, c)
        i[i >= len(self.codes)] = 0
        ok = self.codes[i] == c
        qp, rp = np.nonzero(ok)[0], self.pos[i[ok]]
        if len(qp) < MIN_HITS:
            return None
        b = (rp - qp) // 500
        lo = b.min()
        h = np.bincount(b - lo)
        h2 = h.copy()
        h2[:-1] += h[1:]
        best = int(np.argmax(h2)) + lo
        sel = (rp - qp >= best * 500 - 250) & (rp - qp < (best + 2) * 500 + 250)
        if sel.sum() < MIN_HITS:
            return None
        s, a = np.polyfit(qp[sel].astype(float), rp[sel].astype(float), 1)
        if not 0.8 < s < 1.2:
            return None
        return int(sel.sum()), float(a), float(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--example", default="/root/reference/example")
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--targets", type=int, default=0, help="0 = all reads; N = a cluster of N neighbouring reads")
    ap.add_argument("--anchor", type=int, default=100_000, help="reference position the target cluster starts at")
    args = ap.parse_args()

    refs = read_fasta(os.path.join(args.example, "ref.fa"))
    reads = read_fastq(os.path.join(args.example, "reads.fq.gz"))
    index = [RefIndex(s) for _, s in refs]
    # placement of every read: (ref, strand, a, s) of its best hit; interval on that reference
    place = []
    for name, s, q in reads:
        best = None
        for r, ix in enumerate(index):
            for strand, x in ((0, s), (1, s.translate(_COMP)[::-1])):
                p = ix.place(x)
                if p and (best is None or p[0] > best[0]):
                    best = (p[0], r, strand, p[1], p[2])
        place.append(best)
    placed = [i for i, p in enumerate(place) if p]
    print("reads %d, placed %d; per reference: %s" % (
        len(reads), len(placed), [sum(1 for i in placed if place[i][1] == r) for r in range(len(refs))]))

    def interval(i):
        _, r, strand, a, s = place[i]
        L = len(reads[i][1])
        return r, max(0.0, a), min(float(index[r].length), a + s * L)

    if args.targets:
        cand = sorted((interval(i)[1], i) for i in placed if place[i][1] == 0 and interval(i)[1] >= args.anchor)
        targets = sorted(i for _, i in cand[:args.targets])
    else:
        targets = list(range(len(reads)))
    tset = set(targets)

    def to_read(i, x0, x1):
        """reference interval -> (begin, end) on the read's forward strand"""
        _, r, strand, a, s = place[i]
        L = len(reads[i][1])
        b = int(round(min(max((x0 - a) / s, 0), L)))
        e = int(round(min(max((x1 - a) / s, 0), L)))
        return (L - e, L - b) if strand else (b, e)

    by_ref = {}
    for i in placed:
        by_ref.setdefault(place[i][1], []).append(i)
    lines, used = [], set(targets)
    for t in targets:
        if not place[t]:
            continue
        r, t0, t1 = interval(t)
        for qi in by_ref[r]:
            if qi == t:
                continue
            _, q0, q1 = interval(qi)
            x0, x1 = max(t0, q0), min(t1, q1)
            if x1 - x0 < MIN_OVL:
                continue
            qb, qe = to_read(qi, x0, x1)
            tb, te = to_read(t, x0, x1)
            if qe - qb < 500 or te - tb < 500:
                continue
            strand = "-" if place[qi][2] != place[t][2] else "+"
            alen = max(qe - qb, te - tb)
            lines.append((qi, t, "%s\t%d\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t255" % (
                reads[qi][0], len(reads[qi][1]), qb, qe, strand, reads[t][0], len(reads[t][1]), tb, te,
                min(qe - qb, te - tb), alen)))
            used.add(qi)
    lines.sort(key=lambda x: (x[0], x[1]))

    os.makedirs(args.out_dir, exist_ok=True)

    def write_fq(path, ids):
        with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as f:
            for i in ids:
                n, s, q = reads[i]
                f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + q + b"\n")

    write_fq(os.path.join(args.out_dir, "reads.fq.gz"), sorted(used))
    write_fq(os.path.join(args.out_dir, "targets.fq.gz"), targets)
    with open(os.path.join(args.out_dir, "overlaps.paf"), "w") as f:
        for _, _, l in lines:
            f.write(l + "\n")
    print("targets %d, reads %d, overlaps %d -> %s" % (len(targets), len(used), len(lines), args.out_dir))


if __name__ == "__main__":
    main()
