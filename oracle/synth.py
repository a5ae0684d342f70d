"""TEST / BENCH INFRASTRUCTURE: deterministic synthetic genomes and reads (numpy, seeded).

Used by oracle/make_fixtures.py (golden vectors), tests/ and bench.py.  Nothing here is on the
product path.  The shapes follow SURVEY.md section 8(d): i.i.d. uniform chromosomes with injected
short diverged repeats, optional N runs and short contigs; reads are substrings (either strand)
with a controlled number of substitutions, plus optional junk / N-bearing reads.
"""
from __future__ import annotations

import gzip
import numpy as np

BASES = np.frombuffer(b"ACGTN", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_genome(chrom_lens, seed=1, repeat_frac=0.05, repeat_len=(300, 5000),
                divergences=(0.0, 0.005, 0.01, 0.03), n_runs=0, n_run_len=60):
    """Return list of (name, uint8 codes 0..4).  repeat_frac of each chromosome is overwritten with
    copies of segments drawn from anywhere in the genome (either already written), mutated at one
    of `divergences`."""
    rng = np.random.default_rng(seed)
    chroms = [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in chrom_lens]
    total = sum(len(c) for c in chroms)
    if repeat_frac > 0 and total > 4 * repeat_len[0]:
        budget = int(total * repeat_frac)
        while budget > 0:
            ln = int(rng.integers(repeat_len[0], repeat_len[1] + 1))
            src = chroms[int(rng.integers(0, len(chroms)))]
            dst = chroms[int(rng.integers(0, len(chroms)))]
            if len(src) <= ln or len(dst) <= ln:
                ln = min(len(src), len(dst)) // 2
                if ln < 20:
                    budget -= 20
                    continue
            s = int(rng.integers(0, len(src) - ln))
            d = int(rng.integers(0, len(dst) - ln))
            seg = src[s:s + ln].copy()
            if rng.random() < 0.5:
                seg = _COMP[seg[::-1]]
            div = divergences[int(rng.integers(0, len(divergences)))]
            if div > 0:
                m = rng.random(ln) < div
                seg[m] = (seg[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            dst[d:d + ln] = seg
            budget -= ln
    for _ in range(n_runs):
        c = chroms[int(rng.integers(0, len(chroms)))]
        if len(c) > 4 * n_run_len:
            p = int(rng.integers(0, len(c) - n_run_len))
            c[p:p + n_run_len] = 4
    return [("chr%d" % (i + 1), c) for i, c in enumerate(chroms)]


def write_fasta(path, chroms, width=70):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "wb") as f:
        for name, codes in chroms:
            f.write(b">" + name.encode() + b"\n")
            s = BASES[codes].tobytes()
            for i in range(0, len(s), width * 1000):
                blk = s[i:i + width * 1000]
                f.write(b"\n".join(blk[j:j + width] for j in range(0, len(blk), width)) + b"\n")


def sim_reads(chroms, n, length, seed=2, subs=(0, 1, 2, 3, 4), junk_frac=0.02, n_frac=0.02,
              pe=False, insert=(300, 600), boundary_frac=0.0):
    """Simulate reads.  SE: returns (names, list of uint8 code arrays).  PE: returns
    (names1, reads1, names2, reads2); PE2 is the reverse complement of the fragment's 3' end."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for _, c in chroms], dtype=np.int64)
    cat = np.concatenate([c for _, c in chroms])
    starts = np.concatenate([[0], np.cumsum(lens)])[:-1]

    def mutate(seq, k):
        seq = seq.copy()
        if k > 0:
            pos = rng.choice(len(seq), size=min(k, len(seq)), replace=False)
            seq[pos] = np.where(seq[pos] < 4, (seq[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3,
                                rng.integers(0, 4, size=len(pos), dtype=np.uint8))
        return seq

    def draw(fraglen):
        while True:
            ci = int(rng.choice(len(chroms), p=lens / lens.sum()))
            if lens[ci] >= fraglen:
                break
        if boundary_frac > 0 and rng.random() < boundary_frac and ci + 1 < len(chroms):
            # deliberately straddle the boundary between chromosome ci and ci+1 (concatenated coords)
            p = int(starts[ci] + lens[ci] - rng.integers(1, fraglen))
            p = max(0, min(p, len(cat) - fraglen))
            return -1, p, cat[p:p + fraglen]
        p = int(rng.integers(0, lens[ci] - fraglen + 1))
        return ci, p, chroms[ci][1][p:p + fraglen]

    names1, reads1, names2, reads2 = [], [], [], []
    for i in range(n):
        k = int(subs[int(rng.integers(0, len(subs)))])
        r = rng.random()
        if not pe:
            if r < junk_frac:
                seq = rng.integers(0, 4, size=length, dtype=np.uint8)
                nm = "r%d|junk" % (i + 1)
            else:
                ci, p, seq = draw(length)
                strand = "+" if rng.random() < 0.5 else "-"
                seq = mutate(seq, k)
                if strand == "-":
                    seq = _COMP[seq[::-1]]
                nm = "r%d|%s|%d|%s|%d" % (i + 1, chroms[ci][0] if ci >= 0 else "span", p, strand, k)
            if rng.random() < n_frac:
                seq = seq.copy()
                nn = int(rng.integers(1, 4))
                seq[rng.choice(length, size=nn, replace=False)] = 4
                nm += "|N%d" % nn
            names1.append(nm)
            reads1.append(np.ascontiguousarray(seq))
        else:
            fl = int(rng.integers(insert[0], insert[1] + 1))
            fl = max(fl, length)
            ci, p, frag = draw(fl)
            strand = "+" if rng.random() < 0.5 else "-"
            if strand == "-":
                frag = _COMP[frag[::-1]]
            a = mutate(frag[:length], k)
            k2 = int(subs[int(rng.integers(0, len(subs)))])
            b = mutate(_COMP[frag[-length:][::-1]], k2)
            if r < junk_frac:  # orphan: replace mate with junk
                b = rng.integers(0, 4, size=length, dtype=np.uint8)
            nm = "p%d|%s|%d|%s|%d" % (i + 1, chroms[ci][0] if ci >= 0 else "span", p, strand, fl)
            names1.append(nm + "/1")
            reads1.append(np.ascontiguousarray(a))
            names2.append(nm + "/2")
            reads2.append(np.ascontiguousarray(b))
    if pe:
        return names1, reads1, names2, reads2
    return names1, reads1


def write_reads_fasta(path, names, reads):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "wb") as f:
        for nm, r in zip(names, reads):
            f.write(b">" + nm.encode() + b"\n" + BASES[r].tobytes() + b"\n")


def write_reads_fastq(path, names, reads, qual=b"I"):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "wb") as f:
        for nm, r in zip(names, reads):
            f.write(b"@" + nm.encode() + b"\n" + BASES[r].tobytes() + b"\n+\n" + qual * len(r) + b"\n")
