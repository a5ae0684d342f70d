#!/usr/bin/env python
"""TEST INFRASTRUCTURE (manual tool, not collected by pytest): randomised differential run of the host front end against the
reference binary itself.  Draws option combinations (search options, multi-loci modes, post-alignment passes, output
formats, side files), runs oracle/_ref/biokanga and the front end on the same tiny-genome inputs and compares the main output,
the summary block of the log and every side file.  The front end is either the CPU harness of tests/test_host_cli_cpu.py
(default: built here against tests/bkx_cpu_double.cpp) or, with --cli, the real bkx-align on a GPU box.

    python tests/fuzz_host_cli.py --seeds 0:40 [--jobs 4] [--cli biokanga_b200/bkx-align]
"""
import argparse
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]
REF = os.path.join(ROOT, "oracle", "_ref", "biokanga")
GOLD = os.path.join(HERE, "golden")


def build_cpu_cli(dst):
    odir = os.path.join(ROOT, "oracle", "_build")
    csrc = os.path.join(ROOT, "biokanga_b200", "csrc")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_build/libbkoracle.so"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", dst, os.path.join(csrc, "host", "bkx_align_main.cpp"),
                    os.path.join(HERE, "bkx_cpu_double.cpp"), "-x", "c++", os.path.join(csrc, "bkx_cluster.cu"), "-x", "none",
                    "-L" + odir, "-lbkoracle", "-lz", "-lpthread", "-Wl,-rpath," + odir], check=True)


def summary_block(path):
    keep, on = [], False
    for ln in open(path, errors="replace"):
        body = ln.split("](biokanga) ", 1)[1] if "](biokanga) " in ln else ln
        if "Alignment of" in body and "completed" in body:
            on = True
        if on and body.startswith(("Reporting of aligned result set", "Exit code", "Sorting alignments", "Header written", "Reported SAM",
                                   "Completed reporting SAM", "Reported BAM", "Completed reporting BAM")):
            continue
        if on:
            keep.append(body.rstrip("\n"))
    return keep


def lines(path):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt", errors="replace") as f:
        return f.read().splitlines()


def draw(rng):
    pe = rng.random() < 0.25
    reads = ["pe1.fa", "pe2.fa"] if pe else [rng.choice(["r100.fa", "r50.fa", "mixed.fq", "r150.fa"])]
    a = ["-s%d" % rng.choice([2, 3, 5, 8]), "-e%d" % rng.choice([1, 1, 2]), "-m%d" % rng.choice([0, 0, 1, 3]), "-Q%d" % rng.choice([0, 0, 1, 2])]
    if reads[0] == "mixed.fq":
        a += ["-n%d" % rng.choice([1, 2, 5]), "-g%d" % rng.choice([0, 1, 3])]
    fmt = rng.choice([0, 0, 1, 2, 3, 4, 5, 6])
    ml = 0
    if pe:
        umode = rng.choice([1, 2, 3, 4])
        a += ["-U%d" % umode, "-d%d" % rng.choice([100, 120]), "-D%d" % rng.choice([400, 600, 1500])]
        if fmt in (1, 2, 3):
            fmt = 0
    elif rng.random() < 0.35:
        ml = rng.choice([1, 3, 4, 5])
        a += ["-r%d" % ml, "-R%d" % rng.choice([2, 3, 5, 8])] + (["-X"] if rng.random() < 0.3 else [])
        if ml == 5 and fmt in (1, 2, 3):
            fmt = 0
    a.append("-M%d" % fmt)
    dedup = False
    if rng.random() < 0.4:
        a.append("-x%d" % rng.choice([2, 4, 5, 7]))
    if ml != 5 and rng.random() < 0.2:
        a.append("-6%d" % rng.choice([1, 3, 5]))
    if not pe and rng.random() < 0.3:
        a.append("-k%d" % rng.choice([0, 20, 100, 250]))
        dedup = True
    if rng.random() < 0.3:
        a += rng.choice([["-Zchr2"], ["-zchr[13]"], ["-Zchr1$", "-Z4"], ["-zCHR2", "-Zchr2"]])
    if rng.random() < 0.3:
        a.append("-5cons.csv")
    if rng.random() < 0.25:
        a.append("-#%d" % rng.choice([2, 3, 7]))
    if rng.random() < 0.2:
        a += ["-y%d" % rng.choice([0, 3]), "-Y%d" % rng.choice([0, 4]), "-l%d" % (40 if reads[0] == "r50.fa" else rng.choice([40, 60]))]
    side = []
    if rng.random() < 0.3:
        a += ["-jnone.fa", "-Jmulti.fa"]
        side += ["none.fa", "multi.fa"]
    if fmt != 6 and rng.random() < 0.3:
        a.append("-Ost.csv")
        side.append("st.csv")
    if rng.random() < 0.3:   # drawn last: the earlier draws of a seed stay what they were
        a += ["-B", "pri.bed"] + (["-V"] if rng.random() < 0.5 else [])
    # adaptor trimming at load: every read loses at least a base at each end an adaptor is tried on (the 50-base reads then
    # all fall below the default minimum length of 50: a run without a read)
    if rng.random() < 0.2:
        a += ["-H", "contam.fa"]
    out = "out" + {0: ".csv", 1: ".csv", 2: ".csv", 3: ".csv", 4: ".bed", 5: ".sam", 6: ".sam"}[fmt]
    return reads, a, out, side, dedup, ml


def write_side_files(work):
    """The files some drawn options name: the loci constraints of the `constraints` case and a BED file of priority regions
    over the tiny genome (upper-case and unknown chromosome names, a feature with all six columns, a comment), the adaptor
    file of the `contam` case."""
    shutil.copyfile(os.path.join(GOLD, "constraints", "cons.csv"), os.path.join(work, "cons.csv"))
    shutil.copyfile(os.path.join(GOLD, "contam", "contam.fa"), os.path.join(work, "contam.fa"))
    open(os.path.join(work, "pri.bed"), "w").write("# priority regions\nchr1\t2000\t9000\nchr1 12000 12500\nCHR2\t0\t7000\n"
                                                   "chr3\t2500\t2600\nchr4\t0\t300\tsmall\t0\t-\nchr7\t1\t2\n")


def one(seed, cli, work):
    rng = random.Random(seed)
    reads, args, out, side, dedup, ml = draw(rng)
    d = {}
    for who, exe in (("ref", REF), ("bkx", cli)):
        d[who] = os.path.join(work, "s%d_%s" % (seed, who))
        os.makedirs(d[who], exist_ok=True)
        cmd = [exe, "align", "-I", os.path.join(work, "tiny.sfx"), "-i", os.path.join(work, reads[0])]
        if len(reads) > 1:
            cmd += ["-u", os.path.join(work, reads[1])]
        cmd += [os.path.join(work, x) if i and args[i - 1] in ("-B", "-H") else x if not x.startswith("-5") else "-5" + os.path.join(work, x[2:])
                for i, x in enumerate(args)] + ["-o", out, "-F", "log.txt"]
        if who == "ref":
            cmd.append("-T1" if ml == 5 else "-T4")
        r = subprocess.run(cmd, cwd=d[who], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        d[who + "_rc"] = r.returncode
    tag = "seed %d: %s %s" % (seed, " ".join(reads), " ".join(args))
    if d["ref_rc"] != 0 or d["bkx_rc"] != 0:
        return (tag, "exit codes ref=%d bkx=%d" % (d["ref_rc"], d["bkx_rc"])) if (d["ref_rc"] == 0) != (d["bkx_rc"] == 0) else None
    problems = []

    def key(ln):
        if not dedup:
            return ln
        if out.endswith(".sam"):
            c = ln.split("\t")
            return ln if ln.startswith("@") else "\t".join([c[1], c[2], c[3], c[5]] + c[11:])
        return ",".join(ln.split(",")[1:13]) if out.endswith(".csv") else ln
    a, b = lines(os.path.join(d["ref"], out)), lines(os.path.join(d["bkx"], out))
    if sorted(map(key, a)) != sorted(map(key, b)):
        problems.append("main output differs (%d vs %d lines)" % (len(a), len(b)))
    if summary_block(os.path.join(d["ref"], "log.txt")) != summary_block(os.path.join(d["bkx"], "log.txt")):
        la, lb = summary_block(os.path.join(d["ref"], "log.txt")), summary_block(os.path.join(d["bkx"], "log.txt"))
        diff = [(x, y) for x, y in zip(la, lb) if x != y][:3]
        problems.append("summary differs: %s (%d vs %d lines)" % (diff, len(la), len(lb)))
    for f in side:
        pa, pb = os.path.join(d["ref"], f), os.path.join(d["bkx"], f)
        if os.path.exists(pa) != os.path.exists(pb):
            problems.append("side file %s exists ref=%s bkx=%s" % (f, os.path.exists(pa), os.path.exists(pb)))
        elif os.path.exists(pa):
            ra, rb = lines(pa), lines(pb)
            if f.endswith(".fa"):
                if dedup:
                    continue
                ra, rb = sorted("\n".join(ra).split(">")), sorted("\n".join(rb).split(">"))
            if ra != rb:
                problems.append("side file %s differs" % f)
    if not problems:
        shutil.rmtree(d["ref"]); shutil.rmtree(d["bkx"])
    return (tag, "; ".join(problems)) if problems else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:20")
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--cli", default=None)
    ap.add_argument("--keep", default=None, help="work directory to keep (failing cases stay in it)")
    o = ap.parse_args()
    lo, hi = (int(x) for x in o.seeds.split(":"))
    work = o.keep or tempfile.mkdtemp(prefix="bkxfuzz")
    os.makedirs(work, exist_ok=True)
    for f in ("tiny.sfx", "r100.fa", "r50.fa", "r150.fa", "mixed.fq", "pe1.fa", "pe2.fa"):
        with gzip.open(os.path.join(GOLD, "tiny", f + ".gz"), "rb") as a, open(os.path.join(work, f), "wb") as b:
            shutil.copyfileobj(a, b)
    write_side_files(work)
    cli = o.cli
    if cli is None:
        cli = os.path.join(work, "bkx-align-cpu")
        build_cpu_cli(cli)
    bad = 0
    with ThreadPoolExecutor(o.jobs) as ex:
        for res in ex.map(lambda s: one(s, os.path.abspath(cli), work), range(lo, hi)):
            if res:
                bad += 1
                print("MISMATCH", res[0], "\n   ", res[1], flush=True)
    print("%d seeds, %d mismatches (work dir %s)" % (hi - lo, bad, work))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
