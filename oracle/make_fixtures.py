#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference binary (oracle/_ref/biokanga, built by oracle/build_ref.sh) on seeded synthetic inputs.

The reference ships no tests or fixtures, so these outputs are the pin for the
CPU oracle and, through it, for the CUDA path.  Run in the build container (needs /root/reference
only to build oracle/_ref once):

    python oracle/make_fixtures.py            # regenerates every case

Each case directory holds: genome.fa.gz, the .sfx written by `biokanga index` (gzip), read files,
and per run <tag>.sam.gz (-M6: every read with its YU:Z: class), <tag>.csv.gz (-M0: mismatch counts)
and <tag>.log (alignment summary).  Sizes are kept to a few hundred kB per case.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

REF = os.path.join(HERE, "_ref", "biokanga")
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run(args, cwd):
    r = subprocess.run([REF] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace")[-3000:])
        raise SystemExit("reference failed: %s" % " ".join(args))


def gz(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def strip_log(src, dst):
    """Keep the alignment-summary block of the log without timestamps (they differ per run)."""
    keep = []
    on = False
    for ln in open(src, errors="replace"):
        body = ln.split("](biokanga) ", 1)[1] if "](biokanga) " in ln else ln
        if "Alignment of" in body and "completed" in body:
            on = True
        if on and ("Reporting of aligned result set" in body or "Exit code" in body or "Total processing" in body):
            continue
        if on:
            keep.append(body.rstrip("\n"))
    open(dst, "w").write("\n".join(keep) + "\n")


def align_runs(case_dir, tmp, sfx, runs):
    meta = {}
    for tag, r in runs.items():
        common = ["align", "-I", sfx, "-i", r["reads"][0], "-T4"] + r["args"]
        if len(r["reads"]) > 1:
            common += ["-u", r["reads"][1]]
        run(common + ["-M6", "-o", tag + ".sam", "-F", tag + ".sam.log"], tmp)
        run(common + ["-M0", "-o", tag + ".csv", "-F", tag + ".log"], tmp)
        gz(os.path.join(tmp, tag + ".sam"), os.path.join(case_dir, tag + ".sam.gz"))
        gz(os.path.join(tmp, tag + ".csv"), os.path.join(case_dir, tag + ".csv.gz"))
        strip_log(os.path.join(tmp, tag + ".log"), os.path.join(case_dir, tag + ".log"))
        meta[tag] = {"reads": [os.path.basename(x) + ".gz" for x in r["reads"]], "args": r["args"]}
    json.dump(meta, open(os.path.join(case_dir, "runs.json"), "w"), indent=1, sort_keys=True)


def case_tiny():
    """40 kbp, 4 chromosomes (one 300 bp contig), diverged repeats, one 60-N run."""
    d = os.path.join(GOLD, "tiny")
    os.makedirs(d, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        g = synth.make_genome([20000, 15000, 5000, 300], seed=11, repeat_frac=0.15, repeat_len=(100, 800), n_runs=1)
        synth.write_fasta(os.path.join(tmp, "genome.fa"), g)
        run(["index", "-i", "genome.fa", "-o", "tiny.sfx", "-r", "tiny", "-F", "idx.log"], tmp)
        sets = {
            "r100": dict(n=3000, length=100, seed=12, subs=(0, 1, 2, 3, 4, 5), boundary_frac=0.01),
            "r50": dict(n=1500, length=50, seed=13, subs=(0, 1, 2, 3, 4), boundary_frac=0.01),
            "r150": dict(n=1500, length=150, seed=14, subs=(0, 1, 2, 3, 4, 5, 6, 8), boundary_frac=0.01),
            "r251": dict(n=600, length=251, seed=15, subs=(0, 2, 5, 9, 14), boundary_frac=0.01),
        }
        for nm, kw in sets.items():
            n, r = synth.sim_reads(g, **kw)
            synth.write_reads_fasta(os.path.join(tmp, nm + ".fa"), n, r)
            gz(os.path.join(tmp, nm + ".fa"), os.path.join(d, nm + ".fa.gz"))
        # mixed-length FASTQ set (exercises ragged batches + FASTQ parsing)
        rng = np.random.default_rng(16)
        names, reads = [], []
        for i in range(800):
            L = int(rng.integers(50, 301))
            n1, r1 = synth.sim_reads(g, 1, L, seed=1000 + i, subs=(0, 1, 2, 3), junk_frac=0.05, n_frac=0.05)
            names.append("m%d|%s" % (i + 1, n1[0]))
            reads.append(r1[0])
        synth.write_reads_fastq(os.path.join(tmp, "mixed.fq"), names, reads)
        gz(os.path.join(tmp, "mixed.fq"), os.path.join(d, "mixed.fq.gz"))
        # PE set
        n1, r1, n2, r2 = synth.sim_reads(g, 2000, 100, seed=17, subs=(0, 1, 2, 3), pe=True, insert=(150, 450),
                                         junk_frac=0.05)
        synth.write_reads_fasta(os.path.join(tmp, "pe1.fa"), n1, r1)
        synth.write_reads_fasta(os.path.join(tmp, "pe2.fa"), n2, r2)
        gz(os.path.join(tmp, "pe1.fa"), os.path.join(d, "pe1.fa.gz"))
        gz(os.path.join(tmp, "pe2.fa"), os.path.join(d, "pe2.fa.gz"))
        gz(os.path.join(tmp, "genome.fa"), os.path.join(d, "genome.fa.gz"))
        gz(os.path.join(tmp, "tiny.sfx"), os.path.join(d, "tiny.sfx.gz"))
        runs = {
            "r100_s3": {"reads": ["r100.fa"], "args": ["-s3"]},
            "r100_s5_e2": {"reads": ["r100.fa"], "args": ["-s5", "-e2"]},
            "r100_s0": {"reads": ["r100.fa"], "args": ["-s0"]},
            "r100_s3_n5": {"reads": ["r100.fa"], "args": ["-s3", "-n5"]},
            "r100_s4_m1": {"reads": ["r100.fa"], "args": ["-s4", "-m1"]},
            "r100_s4_m2": {"reads": ["r100.fa"], "args": ["-s4", "-m2"]},
            "r100_s4_m3": {"reads": ["r100.fa"], "args": ["-s4", "-m3"]},
            "r100_s3_Q1": {"reads": ["r100.fa"], "args": ["-s3", "-Q1"]},
            "r100_s3_Q2": {"reads": ["r100.fa"], "args": ["-s3", "-Q2"]},
            "r50_s8": {"reads": ["r50.fa"], "args": ["-s8"]},
            "r50_s8_e2": {"reads": ["r50.fa"], "args": ["-s8", "-e2"]},
            "r150_s3": {"reads": ["r150.fa"], "args": ["-s3"]},
            "r150_s5_e2": {"reads": ["r150.fa"], "args": ["-s5", "-e2"]},
            "r251_s6": {"reads": ["r251.fa"], "args": ["-s6"]},
            "mixed_s3": {"reads": ["mixed.fq"], "args": ["-s3", "-n2"]},
            "pe_U2": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U2", "-d100", "-D1000"]},
            "pe_U4": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U4", "-d100", "-D400"]},
            "pe_U1": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U1", "-d100", "-D600"]},
            "pe_U3": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U3", "-d120", "-D500"]},
            "pe_U1_far": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U1", "-d100", "-D1500"]},
        }
        align_runs(d, tmp, "tiny.sfx", runs)


def case_repeats():
    """~300 kbp with interspersed high-copy short repeats: exercises the 100-candidate probe,
    the MaxIter cap (2500 with -m3, 5000 default) and multi-loci classes."""
    d = os.path.join(GOLD, "repeats")
    os.makedirs(d, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(21)
        unit_a = rng.integers(0, 4, 64, dtype=np.uint8)     # ~3000 copies, ~2850 exact per 25-mer (> 2500, < 5000)
        unit_b = rng.integers(0, 4, 90, dtype=np.uint8)     # ~400 copies   (> 100)
        unit_c = rng.integers(0, 4, 120, dtype=np.uint8)    # ~60 copies    (< 100)
        parts = []

        def emit(unit, copies, div):
            for _ in range(copies):
                u = unit.copy()
                if div > 0:
                    m = rng.random(len(u)) < div
                    u[m] = (u[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
                parts.append(u)
                parts.append(rng.integers(0, 4, int(rng.integers(20, 50)), dtype=np.uint8))

        emit(unit_a, 3000, 0.002)
        emit(unit_b, 400, 0.02)
        emit(unit_c, 60, 0.0)
        order = rng.permutation(len(parts) // 2)
        shuffled = []
        for k in order:
            shuffled.append(parts[2 * k])
            shuffled.append(parts[2 * k + 1])
        body = np.concatenate(shuffled)
        uniq = rng.integers(0, 4, 40000, dtype=np.uint8)
        half = len(body) // 2
        g = [("rep1", np.concatenate([uniq[:20000], body[:half]])), ("rep2", np.concatenate([body[half:], uniq[20000:]]))]
        synth.write_fasta(os.path.join(tmp, "genome.fa"), g)
        run(["index", "-i", "genome.fa", "-o", "repeats.sfx", "-r", "repeats", "-F", "idx.log"], tmp)
        n, r = synth.sim_reads(g, 4000, 100, seed=22, subs=(0, 1, 2, 3), junk_frac=0.01, n_frac=0.01)
        synth.write_reads_fasta(os.path.join(tmp, "r100.fa"), n, r)
        n, r = synth.sim_reads(g, 2000, 60, seed=23, subs=(0, 1, 2), junk_frac=0.01, n_frac=0.01)
        synth.write_reads_fasta(os.path.join(tmp, "r60.fa"), n, r)
        for f in ("r100.fa", "r60.fa", "genome.fa"):
            gz(os.path.join(tmp, f), os.path.join(d, f + ".gz"))
        gz(os.path.join(tmp, "repeats.sfx"), os.path.join(d, "repeats.sfx.gz"))
        runs = {
            "r100_s3": {"reads": ["r100.fa"], "args": ["-s3"]},
            "r100_s5_e2": {"reads": ["r100.fa"], "args": ["-s5", "-e2"]},
            "r100_s3_m3": {"reads": ["r100.fa"], "args": ["-s3", "-m3"]},
            "r100_s3_m2": {"reads": ["r100.fa"], "args": ["-s3", "-m2"]},
            "r60_s5": {"reads": ["r60.fa"], "args": ["-s5"]},
            "r60_s5_e2_m3": {"reads": ["r60.fa"], "args": ["-s5", "-e2", "-m3"]},
        }
        align_runs(d, tmp, "repeats.sfx", runs)


def case_lowcopy():
    """~260 kbp with LOW-copy repeat families (2..9 copies of 150-400 bp units, exact and ~1 % diverged): the regime of the
    multi-loci options -- -r1 (distribution only) with -R limits below, at and above the family sizes, and -X."""
    d = os.path.join(GOLD, "lowcopy")
    os.makedirs(d, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(41)
        parts = []
        for fam in range(120):
            unit = rng.integers(0, 4, int(rng.integers(150, 400)), dtype=np.uint8)
            copies = int(rng.integers(2, 10))
            div = (0.0, 0.0, 0.01)[fam % 3]
            for _ in range(copies):
                u = unit.copy()
                if div > 0:
                    m = rng.random(len(u)) < div
                    u[m] = (u[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
                if rng.random() < 0.4:
                    u = (3 - u[::-1]).astype(np.uint8)
                parts.append(u)
        for _ in range(300):
            parts.append(rng.integers(0, 4, int(rng.integers(100, 500)), dtype=np.uint8))
        order = rng.permutation(len(parts))
        body = np.concatenate([parts[k] for k in order])
        third = len(body) // 3
        g = [("lc1", body[:third]), ("lc2", body[third:2 * third]), ("lc3", body[2 * third:])]
        synth.write_fasta(os.path.join(tmp, "genome.fa"), g)
        run(["index", "-i", "genome.fa", "-o", "lowcopy.sfx", "-r", "lowcopy", "-F", "idx.log"], tmp)
        n, r = synth.sim_reads(g, 5000, 100, seed=42, subs=(0, 0, 1, 2, 3), junk_frac=0.01, n_frac=0.01)
        synth.write_reads_fasta(os.path.join(tmp, "r100.fa"), n, r)
        for f in ("r100.fa", "genome.fa"):
            gz(os.path.join(tmp, f), os.path.join(d, f + ".gz"))
        gz(os.path.join(tmp, "lowcopy.sfx"), os.path.join(d, "lowcopy.sfx.gz"))
        runs = {
            "r0_s3": {"reads": ["r100.fa"], "args": ["-s3"]},
            "r1_R5_s3": {"reads": ["r100.fa"], "args": ["-s3", "-r1", "-R5"]},
            "r1_R2_s3": {"reads": ["r100.fa"], "args": ["-s3", "-r1", "-R2"]},
            "r1_R20_s5_e2": {"reads": ["r100.fa"], "args": ["-s5", "-e2", "-r1", "-R20"]},
            "r1_R4_X_s3": {"reads": ["r100.fa"], "args": ["-s3", "-r1", "-R4", "-X"]},
        }
        align_runs(d, tmp, "lowcopy.sfx", runs)
        # -r5: every locus of a multi-hit read becomes its own record, numbered in arrival order (Aligner.cpp:6668-6716)
        # -- single-threaded so that the numbering is the read order.  CSV (-M0), SAM incl. unaligned (-M6) and the log.
        meta = json.load(open(os.path.join(d, "runs.json")))
        for tag, args in {"r5_R5_s3": ["-s3", "-r5", "-R5"], "r5_R3_X_s3": ["-s3", "-r5", "-R3", "-X"],
                          "r5_R8_s5_e2": ["-s5", "-e2", "-r5", "-R8"]}.items():
            common = ["align", "-I", "lowcopy.sfx", "-i", "r100.fa", "-T1"] + args
            run(common + ["-M6", "-o", tag + ".sam", "-F", tag + ".sam.log"], tmp)
            run(common + ["-M0", "-o", tag + ".csv", "-F", tag + ".log"], tmp)
            gz(os.path.join(tmp, tag + ".sam"), os.path.join(d, tag + ".sam.gz"))
            gz(os.path.join(tmp, tag + ".csv"), os.path.join(d, tag + ".csv.gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            strip_log(os.path.join(tmp, tag + ".sam.log"), os.path.join(d, tag + ".sam.log"))
            meta[tag] = {"reads": ["r100.fa.gz"], "args": args, "all_loci": True}
        # -r3 / -r4: multi-loci reads are assigned one locus by clustering with other reads' loci -- needs depth, so a second
        # read set drawn from the first chromosome only (~14x): CSV + log (the clustering counters are in it)
        n, r = synth.sim_reads(g[:1], 12000, 100, seed=43, subs=(0, 0, 1, 2), junk_frac=0.0, n_frac=0.0)
        synth.write_reads_fasta(os.path.join(tmp, "deep.fa"), n, r)
        gz(os.path.join(tmp, "deep.fa"), os.path.join(d, "deep.fa.gz"))
        for tag, args in {"r3_R5_s3": ["-s3", "-r3", "-R5"], "r4_R5_s3": ["-s3", "-r4", "-R5"],
                          "r4_R8_X_s5": ["-s5", "-r4", "-R8", "-X"]}.items():
            common = ["align", "-I", "lowcopy.sfx", "-i", "deep.fa", "-T4"] + args
            run(common + ["-M6", "-o", tag + ".sam", "-F", tag + ".sam.log"], tmp)
            run(common + ["-M0", "-o", tag + ".csv", "-F", tag + ".log"], tmp)
            gz(os.path.join(tmp, tag + ".sam"), os.path.join(d, tag + ".sam.gz"))
            gz(os.path.join(tmp, tag + ".csv"), os.path.join(d, tag + ".csv.gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"reads": ["deep.fa.gz"], "args": args, "clustered": True}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_formats():
    """Output-format runs on the tiny index: CSV variants -M1..3, BED -M4, FASTQ qualities -g0/-g1 in SAM,
    gzip output.  One FASTQ read set with descriptors that carry trailing words."""
    d = os.path.join(GOLD, "formats")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(tiny, "tiny.sfx.gz"), "rb") as f, open(os.path.join(tmp, "tiny.sfx"), "wb") as g:
            shutil.copyfileobj(f, g)
        import pyoracle as po
        names, bases, offs = po.read_fasta_reads(os.path.join(tiny, "r100.fa.gz"))
        rng = np.random.default_rng(5)
        with open(os.path.join(tmp, "q.fq"), "wb") as f:
            for i in range(600):
                r = bases[offs[i]:offs[i + 1]]
                q = bytes(rng.integers(66, 105, len(r)).astype(np.uint8))   # valid for Sanger and Illumina 1.3+
                f.write(b"@" + names[i].encode() + b" extra words\n" + synth.BASES[r].tobytes() + b"\n+\n" + q + b"\n")
        gz(os.path.join(tmp, "q.fq"), os.path.join(d, "q.fq.gz"))
        # a read set with no two reads from the same (chromosome, position): BAM/BAI bytes then do not depend on how the
        # reference's unstable sort orders equal keys
        seen, n_u = set(), 0
        with open(os.path.join(tmp, "qu.fq"), "wb") as f:
            for i in range(len(names)):
                key = tuple(names[i].split("|")[1:3])
                if key in seen or n_u >= 500:
                    continue
                seen.add(key)
                n_u += 1
                r = bases[offs[i]:offs[i + 1]]
                q = bytes(rng.integers(66, 105, len(r)).astype(np.uint8))
                f.write(b"@" + names[i].encode() + b"\n" + synth.BASES[r].tobytes() + b"\n+\n" + q + b"\n")
        gz(os.path.join(tmp, "qu.fq"), os.path.join(d, "qu.fq.gz"))
        for tag, args, out in (("bam5", ["-s3", "-M5"], "out5.bam"), ("bam6", ["-s3", "-M6", "-g0"], "out6.bam"),
                               ("bamQ2", ["-s3", "-M6", "-Q2"], "out62.bam")):
            run(["align", "-I", "tiny.sfx", "-i", "qu.fq", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            shutil.copyfile(os.path.join(tmp, out), os.path.join(d, out))
            shutil.copyfile(os.path.join(tmp, out + ".bai"), os.path.join(d, out + ".bai"))
        runs = {
            "m1": (["-s3", "-M1"], "m1.csv"), "m2": (["-s3", "-M2"], "m2.csv"), "m3": (["-s3", "-M3"], "m3.csv"),
            "m4": (["-s3", "-M4"], "m4.bed"), "m4t": (["-s3", "-M4", "-tmytrack"], "m4t.bed"),
            "g0": (["-s3", "-M6", "-g0"], "g0.sam"), "g1": (["-s3", "-M5", "-g1"], "g1.sam"),
            "g2": (["-s3", "-M5", "-g2"], "g2.sam"), "gz": (["-s3", "-M0"], "z.csv.gz"),
            "trim": (["-s3", "-M0", "-y5", "-Y7", "-l60"], "trim.csv"),
        }
        meta = {}
        for tag, (args, out) in runs.items():
            run(["align", "-I", "tiny.sfx", "-i", "q.fq", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            if out.endswith(".gz"):
                shutil.copyfile(os.path.join(tmp, out), os.path.join(d, out))
            else:
                gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            meta[tag] = {"args": args, "out": out}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_post():
    """Post-alignment host passes on the tiny index: chromosome filters -Z / -z, flank
    auto-trimming -x (CSV variants, BED, SAM, BAM + BAI), and the non-aligned / multi-aligned read reports -j / -J."""
    d = os.path.join(GOLD, "post")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(tiny, "tiny.sfx.gz"), "rb") as f, open(os.path.join(tmp, "tiny.sfx"), "wb") as g:
            shutil.copyfileobj(f, g)
        import pyoracle as po
        names, bases, offs = po.read_fasta_reads(os.path.join(tiny, "r100.fa.gz"))
        # no two reads from the same (chromosome, position): byte-comparable BAM / BAI (see case_formats)
        seen, keep = set(), []
        for i in range(len(names)):
            key = tuple(names[i].split("|")[1:3])
            if key not in seen and len(keep) < 1200:
                seen.add(key)
                keep.append(i)
        synth.write_reads_fasta(os.path.join(tmp, "p.fa"), [names[i] for i in keep], [bases[offs[i]:offs[i + 1]] for i in keep])
        gz(os.path.join(tmp, "p.fa"), os.path.join(d, "p.fa.gz"))
        runs = {
            "Z2": (["-s3", "-M0", "-Zchr2"], "Z2.csv"), "Z2sam": (["-s3", "-M6", "-Zchr2"], "Z2.sam"),
            "z13": (["-s3", "-M0", "-zchr[13]"], "z13.csv"), "zZ": (["-s3", "-M0", "-zCHR[12]", "-Zchr2", "-Zchr4"], "zZ.csv"),
            "ZZ": (["-s3", "-M5", "-Zchr1$", "-Z^chr3"], "ZZ.sam"),
            "x5": (["-s5", "-M0", "-x5"], "x5.csv"), "x5m3": (["-s5", "-M3", "-x5"], "x5m3.csv"),
            "x5bed": (["-s5", "-M4", "-x5"], "x5.bed"), "x5sam": (["-s5", "-M6", "-x5"], "x5.sam"),
            "x7": (["-s8", "-M0", "-x7"], "x7.csv"), "x3Q1": (["-s6", "-M5", "-x3", "-Q1"], "x3Q1.sam"),
            "x6Z": (["-s6", "-M0", "-x6", "-Zchr3"], "x6Z.csv"),
            "jJ": (["-s3", "-M0", "-jnone.fa", "-Jmulti.fa"], "jJ.csv"),
            "jJr1": (["-s3", "-M0", "-r1", "-R3", "-jnone1.fa", "-Jmulti1.fa.gz"], "jJr1.csv"),
        }
        meta = {}
        for tag, (args, out) in runs.items():
            run(["align", "-I", "tiny.sfx", "-i", "p.fa", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out}
            for a in args:
                if a[:2] in ("-j", "-J"):
                    if a.endswith(".gz"):
                        shutil.copyfile(os.path.join(tmp, a[2:]), os.path.join(d, a[2:]))
                    else:
                        gz(os.path.join(tmp, a[2:]), os.path.join(d, a[2:] + ".gz"))
        # 50 bp reads with up to 4 substitutions: some cannot keep half their length between two 7-base exact flanks (eNARTrim)
        with gzip.open(os.path.join(tiny, "r50.fa.gz"), "rb") as f, open(os.path.join(tmp, "r50.fa"), "wb") as g:
            shutil.copyfileobj(f, g)
        for tag, args, out in (("x7r50", ["-s10", "-M0", "-x7"], "x7r50.csv"), ("x7r50sam", ["-s10", "-M6", "-x7"], "x7r50.sam")):
            run(["align", "-I", "tiny.sfx", "-i", "r50.fa", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": "../tiny/r50.fa.gz"}
        for tag, args, out in (("x5bam", ["-s5", "-M5", "-x5"], "x5.bam"), ("x5bam6", ["-s5", "-M6", "-x5", "-Zchr2"], "x56.bam")):
            run(["align", "-I", "tiny.sfx", "-i", "p.fa", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            shutil.copyfile(os.path.join(tmp, out), os.path.join(d, out))
            shutil.copyfile(os.path.join(tmp, out + ".bai"), os.path.join(d, out + ".bai"))
            meta[tag] = {"args": args, "out": out}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_dups():
    """PCR artefact reduction -k on the tiny index: a read set in which many reads stack on the same start locus (copies of
    a read, some with one more substitution so that the copies differ in LowMMCnt).  Which of several identical copies
    survives depends on the reference's unstable sort: compare rows without read id / name."""
    d = os.path.join(GOLD, "dups")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(tiny, "tiny.sfx.gz"), "rb") as f, open(os.path.join(tmp, "tiny.sfx"), "wb") as g:
            shutil.copyfileobj(f, g)
        import pyoracle as po
        names, bases, offs = po.read_fasta_reads(os.path.join(tiny, "r100.fa.gz"))
        rng = np.random.default_rng(77)
        out_n, out_r = [], []
        for i in range(700):
            r = bases[offs[i]:offs[i + 1]].copy()
            out_n.append(names[i]); out_r.append(r)
            if rng.random() < 0.5:
                for c in range(int(rng.integers(1, 8))):
                    v = r.copy()
                    if rng.random() < 0.3 and v[50] < 4:
                        v[50] = (v[50] + 1) & 3
                    out_n.append("d%d_%s" % (c, names[i])); out_r.append(v)
        perm = rng.permutation(len(out_n))
        synth.write_reads_fasta(os.path.join(tmp, "dup.fa"), [out_n[k] for k in perm], [out_r[k] for k in perm])
        gz(os.path.join(tmp, "dup.fa"), os.path.join(d, "dup.fa.gz"))
        meta = {}
        for tag, args, out in (("k0", ["-s3", "-M0", "-k0"], "k0.csv"), ("k50", ["-s3", "-M0", "-k50"], "k50.csv"),
                               ("k250", ["-s3", "-M0", "-k250"], "k250.csv"), ("k20sam", ["-s3", "-M6", "-k20"], "k20.sam"),
                               ("k0x4", ["-s3", "-M0", "-k0", "-x4", "-Zchr3"], "k0x4.csv")):
            run(["align", "-I", "tiny.sfx", "-i", "dup.fa", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_constraints():
    """Loci base constraints -5 on the tiny index: single-end and paired-end runs (the mate of a violating read goes with it)."""
    d = os.path.join(GOLD, "constraints")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "r100.fa", "pe1.fa", "pe2.fa"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        rows = ['"Chrom","Start","End","Bases"', '"chr1",1000,1010,R', "chr1,5000,5000,AC", 'chr2,3000,3200,"R"', "chr2,7000,7003,ACGT",
                "chr3,100,4000,R", "CHR1,12000,12100,RA", "chr1,15000,15050, g t", "chr2,9000,9400,R", "chr2,9100,9150,CG"]
        open(os.path.join(tmp, "cons.csv"), "w").write("\n".join(rows) + "\n")
        shutil.copyfile(os.path.join(tmp, "cons.csv"), os.path.join(d, "cons.csv"))
        meta = {}
        for tag, reads, args, out in (("c5", ["r100.fa"], ["-s3", "-M0", "-5cons.csv"], "c5.csv"),
                                      ("c5sam", ["r100.fa"], ["-s3", "-M6", "-5cons.csv"], "c5.sam"),
                                      ("c5k", ["r100.fa"], ["-s5", "-M0", "-5cons.csv", "-k0", "-x4"], "c5k.csv"),
                                      ("c5pe", ["pe1.fa", "pe2.fa"], ["-s3", "-M0", "-U2", "-5cons.csv"], "c5pe.csv"),
                                      ("c5pesam", ["pe1.fa", "pe2.fa"], ["-s3", "-M6", "-U1", "-D600", "-5cons.csv"], "c5pe.sam")):
            run(["align", "-I", "tiny.sfx", "-i", reads[0], "-T4", "-o", out, "-F", tag + ".log"] + (["-u", reads[1]] if len(reads) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in reads]}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_sample():
    """Read sampling -# on the tiny index: every Nth raw read (single-end) / read pair (paired-end) of a file."""
    d = os.path.join(GOLD, "sample")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "r100.fa", "pe1.fa", "pe2.fa"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        meta = {}
        # ... and the paired-end form of the -x flank trimming (scans kept inside the outer thirds, nothing sloughed)
        for tag, reads, args, out in (("s7", ["r100.fa"], ["-s3", "-M0", "-#7"], "s7.csv"),
                                      ("s3pe", ["pe1.fa", "pe2.fa"], ["-s3", "-M6", "-U1", "-D600", "-#3"], "s3pe.sam"),
                                      ("p6", ["r100.fa"], ["-s2", "-M0", "-63"], "p6.csv"),          # -6: 5' primer artefact correction
                                      ("p6sam", ["r100.fa"], ["-s2", "-M6", "-63", "-x3", "-k0"], "p6.sam"),
                                      # -6 in paired-end runs without orphan recovery (-U2 / -U4): pairing at -s plus -6, then the correction
                                      ("p6pe", ["pe1.fa", "pe2.fa"], ["-s2", "-M0", "-U2", "-D600", "-63"], "p6pe.csv"),
                                      ("p6pesam", ["pe1.fa", "pe2.fa"], ["-s2", "-M6", "-U4", "-D1500", "-65", "-x3"], "p6pe.sam"),
                                      # ... and with it (-U1 / -U3, -D1500: suffix-array seeded recovery, cores sized from -s, acceptance at -s plus -6)
                                      ("p6u1", ["pe1.fa", "pe2.fa"], ["-s2", "-M0", "-U1", "-D1500", "-63"], "p6u1.csv"),
                                      ("p6u3sam", ["pe1.fa", "pe2.fa"], ["-s5", "-M6", "-U3", "-d120", "-D1500", "-61", "-x4"], "p6u3.sam"),
                                      ("pex0", ["pe1.fa", "pe2.fa"], ["-s5", "-M0", "-U1", "-D600", "-x5"], "pex0.csv"),
                                      ("pex6", ["pe1.fa", "pe2.fa"], ["-s5", "-M6", "-U3", "-D500", "-x7", "-#2"], "pex6.sam")):
            run(["align", "-I", "tiny.sfx", "-i", reads[0], "-T4", "-o", out, "-F", tag + ".log"] + (["-u", reads[1]] if len(reads) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in reads]}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_stats():
    """Statistics file -O on the tiny index: quality-band / substitution / multi-hit distributions and hits per target
    (single-end, FASTQ qualities kept with -g0, also behind -x trimming), insert-length histogram first in paired-end runs."""
    d = os.path.join(GOLD, "stats")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "pe1.fa", "pe2.fa", "mixed.fq"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        meta = {}
        for tag, reads, args, out in (("o1", ["mixed.fq"], ["-s3", "-n2", "-M0", "-g0", "-r1", "-R4", "-Ost.csv"], "o1.csv"),
                                      ("o2", ["mixed.fq"], ["-s5", "-n2", "-M5", "-g1", "-x4", "-Zchr3", "-Ost.csv"], "o2.sam"),
                                      ("o3pe", ["pe1.fa", "pe2.fa"], ["-s3", "-M0", "-U1", "-D600", "-Ost.csv"], "o3pe.csv")):
            run(["align", "-I", "tiny.sfx", "-i", reads[0], "-T4", "-o", out, "-F", tag + ".log"] + (["-u", reads[1]] if len(reads) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            gz(os.path.join(tmp, "st.csv"), os.path.join(d, tag + ".st.csv.gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in reads]}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_interplay():
    """Option combinations found by tests/fuzz_host_cli.py, on the lowcopy index (many reads with several loci): -r5 with
    chromosome filters (loci filtered as they are recorded, incl. the reference's compaction quirk), -r5 with BED (track line
    twice), -j / -J dropped under -r5, -O with BED (histogram only)."""
    d = os.path.join(GOLD, "interplay")
    os.makedirs(d, exist_ok=True)
    low = os.path.join(GOLD, "lowcopy")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("lowcopy.sfx", "r100.fa"):
            with gzip.open(os.path.join(low, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        meta = {}
        for tag, args, out in (("i1", ["-s3", "-M0", "-r5", "-R5", "-Zlc2"], "i1.csv"),
                               ("i2", ["-s3", "-M4", "-r5", "-R5", "-jn.fa", "-Jm.fa", "-Ost.csv"], "i2.bed"),
                               ("i3", ["-s3", "-M5", "-r5", "-R8", "-X", "-zLC[13]", "-x4"], "i3.sam"),
                               ("i4", ["-s3", "-M0", "-r5", "-R3", "-zlc[12]", "-Zlc1", "-k0"], "i4.csv"),
                               ("i5", ["-s3", "-M4", "-r1", "-R4", "-Ost.csv"], "i5.bed")):
            run(["align", "-I", "lowcopy.sfx", "-i", "r100.fa", "-T1", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": ["r100.fa.gz"], "index": "lowcopy"}
            if "-Ost.csv" in args:
                gz(os.path.join(tmp, "st.csv"), os.path.join(d, tag + ".st.csv.gz"))
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_pefilter():
    """Chromosome filters -Z / -z in paired-end runs on the tiny index: they act inside the pairing (AcceptThisChromID) and
    again afterwards (FiltByChroms, which gives the include expressions priority).  Pins the ORACLE's filtered pairing; the
    CUDA pairing kernel does not take the filter yet."""
    d = os.path.join(GOLD, "pefilter")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "pe1.fa", "pe2.fa"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        runs = {
            "U1_Z2": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U1", "-d100", "-D600", "-Zchr2"]},
            "U2_z13": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U2", "-d100", "-D1000", "-zchr[13]"]},
            "U3_ZZ": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U3", "-d120", "-D500", "-Zchr1$", "-Zchr3"]},
            "U4_zZ": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U4", "-d100", "-D400", "-zchr[12]", "-Zchr2"]},
            "U1far_z2": {"reads": ["pe1.fa", "pe2.fa"], "args": ["-s3", "-U1", "-d100", "-D1500", "-zchr2"]},
        }
        align_runs(d, tmp, "tiny.sfx", runs)


def case_simreads():
    """Reads made by the reference's own `biokanga simreads` (descriptors lcl|usimreads|<n>|<chrom>|<start>|<end>|<len>|<strand>|...):
    `align` checks every accepted alignment against the origin named in the descriptor (ReportAlignStats, Aligner.cpp:3556-3657),
    logs "There are N (a 2 edge, b 1 edge) high confidence aligned simulated reads with M misaligned" and writes misaligned
    reads as "iar" in CSV / BED (Aligner.cpp:6405-6418).  On the `repeats` genome (so that misaligned reads occur), on the
    `lowcopy` genome under -r5, and on a small genome whose chromosome names hold '|' (the second descriptor form).
    simreads seeds itself from the clock: the read files it wrote are committed, not re-creatable."""
    d = os.path.join(GOLD, "simreads")
    os.makedirs(d, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for case, f in (("repeats", "repeats.sfx"), ("repeats", "genome.fa"), ("repeats", "r100.fa"), ("lowcopy", "lowcopy.sfx"),
                        ("lowcopy", "genome.fa")):
            dst = f if f != "genome.fa" else case + ".fa"
            with gzip.open(os.path.join(GOLD, case, f + ".gz"), "rb") as a, open(os.path.join(tmp, dst), "wb") as b:
                shutil.copyfileobj(a, b)
        # a genome whose chromosome names are three '|' separated parts
        g = synth.make_genome([9000, 7000, 5000], seed=71, repeat_frac=0.2, repeat_len=(100, 600), n_runs=0)
        g = [("gnl|UG|Ta%d" % (k + 1), seq) for k, (_, seq) in enumerate(g)]
        synth.write_fasta(os.path.join(tmp, "piped.fa"), g)
        run(["index", "-i", "piped.fa", "-o", "piped.sfx", "-r", "piped", "-F", "idx.log"], tmp)
        gz(os.path.join(tmp, "piped.sfx"), os.path.join(d, "piped.sfx.gz"))
        sim = lambda genome, out, extra: run(["simreads", "-i", genome, "-o", out, "-T1", "-F", out + ".log"] + extra, tmp)
        sim("repeats.fa", "se_g1.fa", ["-n", "3000", "-l", "100", "-g1", "-z2"])
        sim("repeats.fa", "se_g3.fa", ["-n", "2000", "-l", "80", "-g3", "-z0.15"])
        sim("repeats.fa", "pe1.fa", ["-O", "pe2.fa", "-p", "-n", "3000", "-l", "100", "-j", "250", "-J", "450", "-g1", "-z1"])
        sim("lowcopy.fa", "lc.fa", ["-n", "2000", "-l", "100", "-g1", "-z1"])
        sim("piped.fa", "piped_se.fa", ["-n", "1500", "-l", "100", "-g1", "-z2"])

        def fasta_records(path):
            return [">" + x for x in open(path).read().split(">")[1:]]
        plain = fasta_records(os.path.join(tmp, "r100.fa"))
        simr = fasta_records(os.path.join(tmp, "se_g1.fa"))
        # not simulated first: the first accepted read decides, nothing is checked; simulated first, then plain reads: the check
        # stops for good at the first accepted read whose descriptor does not parse
        open(os.path.join(tmp, "mix_a.fa"), "w").write("".join(plain[:300] + simr[:700]))
        open(os.path.join(tmp, "mix_b.fa"), "w").write("".join(simr[:500] + plain[:200] + simr[500:900]))
        reads = ["se_g1.fa", "se_g3.fa", "pe1.fa", "pe2.fa", "lc.fa", "piped_se.fa", "mix_a.fa", "mix_b.fa"]
        for f in reads:
            gz(os.path.join(tmp, f), os.path.join(d, f + ".gz"))
        meta = {}
        for tag, index, rds, args, out in (
                ("se_g1", "repeats", ["se_g1.fa"], ["-s3", "-M0"], "se_g1.csv"),
                ("se_g1bed", "repeats", ["se_g1.fa"], ["-s3", "-M4"], "se_g1.bed"),
                ("se_g1sam", "repeats", ["se_g1.fa"], ["-s3", "-M6"], "se_g1.sam"),
                ("se_g1x", "repeats", ["se_g1.fa"], ["-s5", "-M0", "-x6"], "se_g1x.csv"),
                ("se_g3", "repeats", ["se_g3.fa"], ["-s8", "-M0", "-e2"], "se_g3.csv"),
                ("se_g3r1", "repeats", ["se_g3.fa"], ["-s8", "-M2", "-r1", "-R4", "-Zrep1"], "se_g3r1.csv"),
                ("pe_U1", "repeats", ["pe1.fa", "pe2.fa"], ["-s3", "-M0", "-U1", "-d100", "-D600"], "pe_U1.csv"),
                ("pe_U2", "repeats", ["pe1.fa", "pe2.fa"], ["-s3", "-M0", "-U2", "-d100", "-D400"], "pe_U2.csv"),
                ("pe_U3sam", "repeats", ["pe1.fa", "pe2.fa"], ["-s3", "-M6", "-U3", "-d100", "-D600"], "pe_U3.sam"),
                ("lc_r5", "lowcopy", ["lc.fa"], ["-s3", "-M0", "-r5", "-R5"], "lc_r5.csv"),
                ("lc_r3", "lowcopy", ["lc.fa"], ["-s3", "-M0", "-r3", "-R5"], "lc_r3.csv"),
                ("piped", "piped", ["piped_se.fa"], ["-s3", "-M0"], "piped.csv"),
                ("pipedbed", "piped", ["piped_se.fa"], ["-s5", "-M4", "-x4"], "piped.bed"),
                ("mix_a", "repeats", ["mix_a.fa"], ["-s3", "-M0"], "mix_a.csv"),
                ("mix_b", "repeats", ["mix_b.fa"], ["-s3", "-M0"], "mix_b.csv")):
            threads = "-T1" if "-r5" in args else "-T4"
            run(["align", "-I", index + ".sfx", "-i", rds[0], threads, "-o", out, "-F", tag + ".log"] + (["-u", rds[1]] if len(rds) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in rds], "index": index}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_grammar():
    """Option grammar of the front end (kanga.cpp:194-298): a parameter file (@file: several options per line, comments,
    quoted values), long option names, and a single-end input specification with wildcards -- three read files whose
    case-insensitive name order (A_, b_, C_) differs from their byte order, so the read ids in the CSV pin the load order."""
    d = os.path.join(GOLD, "grammar")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "r100.fa"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        recs = [">" + x for x in open(os.path.join(tmp, "r100.fa")).read().split(">")[1:]]
        parts = {"b_part.fa": recs[:400], "A_part.fa": recs[400:900], "C_part.fa": recs[900:1300]}
        for nm, rr in parts.items():
            open(os.path.join(tmp, nm), "w").write("".join(rr))
            gz(os.path.join(tmp, nm), os.path.join(d, nm + ".gz"))
        params = ["# parameter file for the grammar fixture", "", "-s3 --format=0", "; another comment", "// and another",
                  "--editdelta 1   -T4", '-t "my track"', "--in=*_part.fa"]
        open(os.path.join(tmp, "params.txt"), "w").write("\n".join(params) + "\n")
        shutil.copyfile(os.path.join(tmp, "params.txt"), os.path.join(d, "params.txt"))
        run(["align", "@params.txt", "-I", "tiny.sfx", "--out", "g1.csv", "--log=g1.log"], tmp)
        gz(os.path.join(tmp, "g1.csv"), os.path.join(d, "g1.csv.gz"))
        strip_log(os.path.join(tmp, "g1.log"), os.path.join(d, "g1.log"))


def case_bestmatches():
    """-N (best matches, CSfxArrayV3::LocateBestMatches): the -R<n> loci with the fewest mismatches from one un-staged pass,
    behind every multi-loci mode that can be repeated (-r1 stats only, -r3 / -r4 clustering, -r5 all loci), on the lowcopy
    genome (2..9-copy repeat families).  -T1 so that the -r5 record numbering is the read order."""
    d = os.path.join(GOLD, "bestmatches")
    os.makedirs(d, exist_ok=True)
    low = os.path.join(GOLD, "lowcopy")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("lowcopy.sfx", "r100.fa", "deep.fa"):
            with gzip.open(os.path.join(low, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        meta = {}
        for tag, reads, args, out in (("n5_r5", "r100.fa", ["-s3", "-M0", "-r5", "-R5", "-N"], "n5_r5.csv"),
                                      ("n5_r5sam", "r100.fa", ["-s3", "-M6", "-r5", "-R5", "-N"], "n5_r5.sam"),
                                      ("n3_r5s8", "r100.fa", ["-s8", "-e2", "-M0", "-r5", "-R3", "-N"], "n3_r5s8.csv"),
                                      ("n8_r5Q1", "r100.fa", ["-s5", "-M0", "-r5", "-R8", "-N", "-Q1"], "n8_r5Q1.csv"),
                                      ("n2_r5s0", "r100.fa", ["-s0", "-M0", "-r5", "-R2", "-N"], "n2_r5s0.csv"),
                                      ("n4_r1", "r100.fa", ["-s5", "-M0", "-r1", "-R4", "-N"], "n4_r1.csv"),
                                      ("n5_r3", "deep.fa", ["-s3", "-M0", "-r3", "-R5", "-N"], "n5_r3.csv"),
                                      ("n6_r4", "deep.fa", ["-s5", "-M0", "-r4", "-R6", "-N"], "n6_r4.csv")):
            run(["align", "-I", "lowcopy.sfx", "-i", reads, "-T1", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [reads + ".gz"], "index": "lowcopy"}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_manyloci():
    """-R beyond a few loci (the reference takes up to 500 with -r1/-r3/-r4, more with -r5) on the `repeats` genome, whose
    reads from the 400-copy and 3000-copy units carry hundreds of equally good loci: -r5 with and without -X, -r4, and -N."""
    d = os.path.join(GOLD, "manyloci")
    os.makedirs(d, exist_ok=True)
    rep = os.path.join(GOLD, "repeats")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("repeats.sfx", "r100.fa", "r60.fa"):
            with gzip.open(os.path.join(rep, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        recs = [">" + x for x in open(os.path.join(tmp, "r60.fa")).read().split(">")[1:]]
        open(os.path.join(tmp, "r60a.fa"), "w").write("".join(recs[:400]))     # keeps the -r5 outputs small
        gz(os.path.join(tmp, "r60a.fa"), os.path.join(d, "r60a.fa.gz"))
        meta = {}
        # 60 bp reads inside the 64-base unit have ~2850 exact loci, inside the 90-base unit dozens to hundreds within -s5
        for tag, reads, args, out in (("r5_R300X", "r60a.fa", ["-s5", "-M0", "-r5", "-R300", "-X"], "r5_R300X.csv"),
                                      ("r5_R500", "r60a.fa", ["-s5", "-M0", "-r5", "-R500"], "r5_R500.csv"),
                                      ("r5_R100N", "r60a.fa", ["-s5", "-M0", "-r5", "-R100", "-N"], "r5_R100N.csv"),
                                      ("r4_R200X", "r60a.fa", ["-s5", "-M0", "-r4", "-R200", "-X"], "r4_R200X.csv"),
                                      ("r1_R500", "r60a.fa", ["-s5", "-M0", "-r1", "-R500"], "r1_R500.csv")):
            run(["align", "-I", "repeats.sfx", "-i", reads, "-T1", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [reads + ".gz"], "index": "repeats", "reads_dir": "manyloci"}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_priority():
    """-B / -V (priority regions, Aligner.cpp:9102-9186, 4126-4186) on the lowcopy genome (2..9-copy repeat families): reads
    with several equally good loci become unique when exactly one of them lies inside a region of the BED file; without -V
    the accepted alignments outside every region are then filtered (PR).  One BED file with tabs, upper-case chromosome names,
    comments, a header line and overlapping features; one with commas (which the reference's sscanf only takes with
    white space around them)."""
    d = os.path.join(GOLD, "priority")
    os.makedirs(d, exist_ok=True)
    low = os.path.join(GOLD, "lowcopy")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("lowcopy.sfx", "r100.fa", "deep.fa"):
            with gzip.open(os.path.join(low, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        open(os.path.join(tmp, "pri.bed"), "w").write(
            "# priority regions of the lowcopy genome\ntrack name=priority\nlc1\t0\t30000\tfirst\t0\t+\nlc2\t40000\t90861\n"
            "LC3\t1000\t2000\tup\nlc3\t50000\t50100\nlc1 29500 31000 overlap 5 -\nlc9\t5\t500\n")
        open(os.path.join(tmp, "pri_csv.bed"), "w").write("lc2 , 100 , 45000\nlc3 , 60000 , 90000\n")
        for f in ("pri.bed", "pri_csv.bed"):
            shutil.copy(os.path.join(tmp, f), os.path.join(d, f))
        # paired ends on the same genome (the pairing sees the records the priority regions decided)
        chroms, name, seq = [], None, []
        for ln in gzip.open(os.path.join(low, "genome.fa.gz"), "rt"):
            if ln.startswith(">"):
                if name:
                    chroms.append((name, np.array(["ACGT".index(c) for c in "".join(seq).upper()], dtype=np.uint8)))
                name, seq = ln[1:].split()[0], []
            else:
                seq.append(ln.strip())
        chroms.append((name, np.array(["ACGT".index(c) for c in "".join(seq).upper()], dtype=np.uint8)))
        n1, r1, n2, r2 = synth.sim_reads(chroms, 1200, 100, seed=77, subs=(0, 1, 2), pe=True, insert=(150, 450), junk_frac=0.03)
        synth.write_reads_fasta(os.path.join(tmp, "lpe1.fa"), n1, r1)
        synth.write_reads_fasta(os.path.join(tmp, "lpe2.fa"), n2, r2)
        gz(os.path.join(tmp, "lpe1.fa"), os.path.join(d, "lpe1.fa.gz"))
        gz(os.path.join(tmp, "lpe2.fa"), os.path.join(d, "lpe2.fa.gz"))
        meta = {}
        for tag, args, out in (("b_pe", ["-s3", "-M0", "-U1", "-D600", "-B", "pri.bed"], "b_pe.csv"),
                               ("bv_pe", ["-s3", "-M6", "-U2", "-D600", "-B", "pri.bed", "-V"], "bv_pe.sam"),
                               ("b_pe3", ["-s5", "-M0", "-U3", "-d120", "-D500", "-B", "pri_csv.bed", "-Zlc1"], "b_pe3.csv"),
                               ("bv_pe4", ["-s3", "-M0", "-U4", "-D600", "-B", "pri.bed", "-V", "-x3"], "bv_pe4.csv")):
            run(["align", "-I", "lowcopy.sfx", "-i", "lpe1.fa", "-u", "lpe2.fa", "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": ["lpe1.fa.gz", "lpe2.fa.gz"], "index": "lowcopy", "reads_dir": "priority"}
        for tag, reads, args, out in (("b_s3", "r100.fa", ["-s3", "-M0", "-B", "pri.bed"], "b_s3.csv"),
                                      ("bv_s3", "r100.fa", ["-s3", "-M0", "-B", "pri.bed", "-V"], "bv_s3.csv"),
                                      ("bv_s5e2", "r100.fa", ["-s5", "-e2", "-M0", "-B", "pri.bed", "-V"], "bv_s5e2.csv"),
                                      ("b_csv_sam", "r100.fa", ["-s3", "-M6", "-B", "pri_csv.bed"], "b_csv_sam.sam"),
                                      ("bv_deep", "deep.fa", ["-s3", "-M0", "-B", "pri_csv.bed", "-V"], "bv_deep.csv"),
                                      ("b_z", "r100.fa", ["-s3", "-M0", "-B", "pri.bed", "-Z", "lc2"], "b_z.csv"),
                                      ("bv_x2k", "r100.fa", ["-s4", "-M0", "-B", "pri.bed", "-V", "-x2", "-k1"], "bv_x2k.csv"),
                                      # behind the multi-loci modes: up to -R loci inside the regions stand for the read
                                      ("bv_r1", "r100.fa", ["-s3", "-M0", "-r1", "-R3", "-B", "pri.bed", "-V"], "bv_r1.csv"),
                                      ("b_r1x", "r100.fa", ["-s5", "-M0", "-r1", "-R2", "-X", "-B", "pri.bed"], "b_r1x.csv"),
                                      ("bv_r3", "deep.fa", ["-s3", "-M0", "-r3", "-R5", "-B", "pri.bed", "-V"], "bv_r3.csv"),
                                      ("b_r4x", "deep.fa", ["-s3", "-M0", "-r4", "-R3", "-X", "-B", "pri_csv.bed"], "b_r4x.csv"),
                                      ("bv_r5", "r100.fa", ["-s3", "-M0", "-r5", "-R4", "-B", "pri.bed", "-V"], "bv_r5.csv"),
                                      ("b_r5x", "r100.fa", ["-s3", "-M0", "-r5", "-R2", "-X", "-B", "pri_csv.bed"], "b_r5x.csv"),
                                      ("bv_r5n", "r100.fa", ["-s5", "-M0", "-r5", "-R3", "-N", "-B", "pri.bed", "-V"], "bv_r5n.csv")):
            run(["align", "-I", "lowcopy.sfx", "-i", reads, "-T1" if "-r5" in args else "-T4", "-o", out, "-F", tag + ".log"] + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [reads + ".gz"], "index": "lowcopy"}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_contam():
    """-H (adaptor trimming at load, CContaminants): reads of the tiny case with adaptor tails written over their ends (the
    last k bases of a 5' adaptor over the first k of the read, the first k of a 3' adaptor over its last k; one substitution
    in a tenth of them), a contaminant file with every kind of overlay code -- @1, @3, @1234, reverse complements (@57),
    no code (1, 2, 5, 6), an N inside, PE2 only, the minimum length of 4 --; single-end, with fixed trims and a length filter
    on top, FASTQ with qualities, SAM, paired ends with orphan recovery and with sampling."""
    d = os.path.join(GOLD, "contam")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    rng = np.random.default_rng(424242)
    ad5 = ["ACACTCTTTCCCTACACGACGCTCTTCCGATCT", "CTGTCTCTTATACACATCT", "AATGATACGGCGACCACCGA", "TTNGACCATG"]
    ad3 = ["AGATCGGAAGAGCACACGTCTGAACTCCAGTCA", "CTGTCTCTTATACACATCT", "TTNGACCATG", "ACACTCTTTCCCTACACGACGCTCTTCCGATC"]

    def spoil(seq):
        s = list(seq)
        if rng.random() < 0.5:
            a = ad5[int(rng.integers(0, len(ad5)))].replace("N", "A")
            k = int(rng.integers(1, min(len(a), 30) + 1))
            s[:k] = list(a[len(a) - k:])
            if rng.random() < 0.1:
                s[int(rng.integers(0, k))] = "ACGT"[int(rng.integers(0, 4))]
        if rng.random() < 0.5:
            a = ad3[int(rng.integers(0, len(ad3)))].replace("N", "C")
            k = int(rng.integers(1, min(len(a), 30) + 1))
            s[len(s) - k:] = list(a[:k])
            if rng.random() < 0.1:
                s[len(s) - 1 - int(rng.integers(0, k))] = "ACGT"[int(rng.integers(0, 4))]
        return "".join(s)

    with tempfile.TemporaryDirectory() as tmp:
        with gzip.open(os.path.join(tiny, "tiny.sfx.gz"), "rb") as a, open(os.path.join(tmp, "tiny.sfx"), "wb") as b:
            shutil.copyfileobj(a, b)
        for src, dst, limit in (("r100.fa", "c100.fa", 1500), ("pe1.fa", "cpe1.fa", 800), ("pe2.fa", "cpe2.fa", 800)):
            recs = gzip.open(os.path.join(tiny, src + ".gz"), "rt").read().split(">")[1:limit + 1]
            with open(os.path.join(tmp, dst), "w") as o:
                for r in recs:
                    name, seq = r.split("\n", 1)
                    o.write(">%s\n%s\n" % (name, spoil(seq.replace("\n", ""))))
        lines = gzip.open(os.path.join(tiny, "mixed.fq.gz"), "rt").read().splitlines()
        with open(os.path.join(tmp, "cmixed.fq"), "w") as o:
            for i in range(0, min(len(lines), 4 * 1200), 4):
                seq = spoil(lines[i + 1]) if len(lines[i + 1]) >= 40 else lines[i + 1]
                o.write("%s\n%s\n+\n%s\n" % (lines[i], seq, lines[i + 3]))
        open(os.path.join(tmp, "contam.fa"), "w").write(
            ">ad5@1 five prime adaptor\nACACTCTTTCCCTACACGACGCTCTTCCGATCT\n>ad3@3\nAGATCGGAAGAGCACACGTC\nTGAACTCCAGTCA\n"
            ">both@1234\nCTGTCTCTTATACACATCT\n>rc7@57\nGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT\n>nocode\nAATGATACGGCGACCACCGA\n"
            ">withn@13\nTTNGACCATG\n>pe2only@24\nCAAGCAGAAGACGGCATACGAGAT\n>tiny4@3\nGGCC\n")
        # vectors ('&' codes): reads cut out of a vector sequence -- as they are or reverse complemented, with 0..6
        # substitutions (100 / 25 = 4 are tolerated) -- lie inside it and fall to the length filter; vec2 is marked for PE2
        # reads only, yet the reference also tries it on the 3' end of SE / PE1 reads
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        vecs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, n)) for n in (420, 260)]

        def from_vector(v, length):
            p = int(rng.integers(0, len(v) - length + 1))
            s = list(v[p:p + length])
            for _ in range(int(rng.integers(0, 7))):
                i = int(rng.integers(0, length))
                s[i] = "ACGT"[("ACGT".index(s[i]) + 1 + int(rng.integers(0, 3))) % 4]
            s = "".join(s)
            return s if rng.random() < 0.5 else "".join(comp[c] for c in reversed(s))

        for src, dst in (("c100.fa", "cv100.fa"), ("cpe1.fa", "cvpe1.fa"), ("cpe2.fa", "cvpe2.fa")):
            recs = open(os.path.join(tmp, src)).read().split(">")[1:601]
            with open(os.path.join(tmp, dst), "w") as o:
                for r in recs:
                    name, seq = r.split("\n", 1)
                    seq = seq.strip()
                    if rng.random() < 0.25:
                        seq = from_vector(vecs[int(rng.integers(0, 2))], len(seq))
                    o.write(">%s\n%s\n" % (name, seq))
        open(os.path.join(tmp, "contam_v.fa"), "w").write(
            ">ad5@1\nACACTCTTTCCCTACACGACGCTCTTCCGATCT\n>vec1&15 cloning vector\n%s\n>ad3@34\nAGATCGGAAGAGCACACGTCTGAACTCCAGTCA\n>vec2&2\n%s\n"
            % ("\n".join(vecs[0][i:i + 70] for i in range(0, len(vecs[0]), 70)), vecs[1]))
        for fn in ("c100.fa", "cpe1.fa", "cpe2.fa", "cmixed.fq", "cv100.fa", "cvpe1.fa", "cvpe2.fa"):
            gz(os.path.join(tmp, fn), os.path.join(d, fn + ".gz"))
        shutil.copy(os.path.join(tmp, "contam.fa"), os.path.join(d, "contam.fa"))
        shutil.copy(os.path.join(tmp, "contam_v.fa"), os.path.join(d, "contam_v.fa"))
        meta = {}
        for tag, reads, args, out in (("h_s3", ["c100.fa"], ["-s3", "-M0", "-H", "contam.fa"], "h_s3.csv"),
                                      ("h_trim", ["c100.fa"], ["-s4", "-M0", "-H", "contam.fa", "-y3", "-Y2", "-l80"], "h_trim.csv"),
                                      ("h_sam", ["c100.fa"], ["-s3", "-M6", "-H", "contam.fa"], "h_sam.sam"),
                                      ("h_fq", ["cmixed.fq"], ["-s3", "-M6", "-g0", "-n2", "-H", "contam.fa"], "h_fq.sam"),
                                      ("h_pe", ["cpe1.fa", "cpe2.fa"], ["-s3", "-M0", "-U1", "-D600", "-H", "contam.fa"], "h_pe.csv"),
                                      ("h_pesam", ["cpe1.fa", "cpe2.fa"], ["-s5", "-M6", "-U3", "-D600", "-#2", "-H", "contam.fa", "-y1"], "h_pe.sam"),
                                      ("h_x", ["c100.fa"], ["-s5", "-M0", "-H", "contam.fa", "-x3", "-Zchr2"], "h_x.csv"),
                                      ("hv_s3", ["cv100.fa"], ["-s3", "-M0", "-H", "contam_v.fa"], "hv_s3.csv"),
                                      ("hv_sam", ["cv100.fa"], ["-s5", "-M6", "-H", "contam_v.fa", "-l40"], "hv_sam.sam"),
                                      ("hv_pe", ["cvpe1.fa", "cvpe2.fa"], ["-s3", "-M0", "-U1", "-D600", "-H", "contam_v.fa"], "hv_pe.csv")):
            run(["align", "-I", "tiny.sfx", "-i", reads[0], "-T4", "-o", out, "-F", tag + ".log"] + (["-u", reads[1]] if len(reads) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in reads], "index": "tiny"}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


def case_empty():
    """Runs in which no read survives the load filters (-l above every read length): the reference goes on and reports an empty
    run -- average length 0 with a minimum of -1, one unprocessed record in the class summary, an empty record in -M6."""
    d = os.path.join(GOLD, "empty")
    os.makedirs(d, exist_ok=True)
    tiny = os.path.join(GOLD, "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("tiny.sfx", "r50.fa", "pe1.fa", "pe2.fa"):
            with gzip.open(os.path.join(tiny, f + ".gz"), "rb") as a, open(os.path.join(tmp, f), "wb") as b:
                shutil.copyfileobj(a, b)
        meta = {}
        for tag, reads, args, out in (("e_m0", ["r50.fa"], ["-s3", "-M0", "-l60"], "e_m0.csv"),
                                      ("e_m6", ["r50.fa"], ["-s3", "-M6", "-l60"], "e_m6.sam"),
                                      ("e_m4x", ["r50.fa"], ["-s0", "-M4", "-l60", "-x3", "-k0"], "e_m4x.bed"),
                                      ("e_pe", ["pe1.fa", "pe2.fa"], ["-s3", "-M0", "-U1", "-l120"], "e_pe.csv"),
                                      ("e_r5", ["r50.fa"], ["-s3", "-M0", "-l60", "-r5", "-R3"], "e_r5.csv"),
                                      ("e_r4", ["r50.fa"], ["-s5", "-M6", "-l60", "-r4", "-R8", "-x7"], "e_r4.sam"),
                                      ("e_r1", ["r50.fa"], ["-s3", "-M0", "-l60", "-r1", "-R3", "-k0"], "e_r1.csv")):
            run(["align", "-I", "tiny.sfx", "-i", reads[0], "-T4", "-o", out, "-F", tag + ".log"] + (["-u", reads[1]] if len(reads) > 1 else []) + args, tmp)
            gz(os.path.join(tmp, out), os.path.join(d, out + ".gz"))
            strip_log(os.path.join(tmp, tag + ".log"), os.path.join(d, tag + ".log"))
            meta[tag] = {"args": args, "out": out, "reads": [r + ".gz" for r in reads], "index": "tiny"}
        json.dump(meta, open(os.path.join(d, "runs.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit("build oracle/_ref first: oracle/build_ref.sh")
    which = sys.argv[1:] or ["tiny", "repeats", "formats", "lowcopy", "post", "dups", "constraints", "sample", "stats", "interplay", "pefilter", "simreads", "grammar", "bestmatches", "manyloci", "priority", "contam", "empty"]
    if "tiny" in which:
        case_tiny()
    if "repeats" in which:
        case_repeats()
    if "formats" in which:
        case_formats()
    if "lowcopy" in which:
        case_lowcopy()
    if "post" in which:
        case_post()
    if "dups" in which:
        case_dups()
    if "constraints" in which:
        case_constraints()
    if "sample" in which:
        case_sample()
    if "stats" in which:
        case_stats()
    if "interplay" in which:
        case_interplay()
    if "pefilter" in which:
        case_pefilter()
    if "simreads" in which:
        case_simreads()
    if "grammar" in which:
        case_grammar()
    if "bestmatches" in which:
        case_bestmatches()
    if "manyloci" in which:
        case_manyloci()
    print("fixtures written under", GOLD)
    if "priority" in which:
        case_priority()
    if "contam" in which:
        case_contam()
    if "empty" in which:
        case_empty()
