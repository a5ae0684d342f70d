"""Helpers for the golden vectors under tests/golden/ (made by oracle/make_fixtures.py from the
reference binary).  Test infrastructure only."""
from __future__ import annotations

import csv
import gzip
import io
import json
import os
import shutil

import numpy as np

import pyoracle as po
from biokanga_b200 import abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"tiny": "tiny.sfx", "repeats": "repeats.sfx", "lowcopy": "lowcopy.sfx"}


def sfx_path(case, scratch):
    dst = os.path.join(str(scratch), CASES[case])
    if not os.path.exists(dst):
        with gzip.open(os.path.join(GOLD, case, CASES[case] + ".gz"), "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
    return dst


def runs(case):
    return json.load(open(os.path.join(GOLD, case, "runs.json")))


def all_runs(all_loci=True):
    """Every (case, tag); all_loci=False leaves out the -r3..5 runs (hit lists / clustering: they have their own checks)."""
    out = []
    for case in CASES:
        rs = runs(case)
        for tag in sorted(rs):
            if all_loci or not (rs[tag].get("all_loci") or rs[tag].get("clustered")):
                out.append((case, tag))
    return out


def params_from_args(index, args):
    """Map the reference CLI flags used by the fixtures onto bkx_align_params / bkx_pe_params
    (kanga.cpp:194-294 option table)."""
    opt = {"s": 10, "e": 1, "n": 1, "m": 0, "Q": 0, "U": 0, "d": 100, "D": 1000}
    for a in args:
        opt[a[1]] = int(a[2:]) if len(a) > 2 else 1
    p = index.default_params(opt["m"], max_subs=opt["s"], min_edit_dist=opt["e"], max_ns=opt["n"],
                             align_strand=opt["Q"])
    if opt.get("r", 0):  # multi-loci modes: -r<mode> -R<limit> [-X]  (kanga.cpp:667-696)
        p.ml_mode = opt["r"]
        p.max_ml_matches = opt.get("R", 5)
        p.clamp_max_ml = 1 if "X" in opt else 0
    pe = None
    if opt["U"]:
        pe = abi.PEParams()
        pe.pe_proc, pe.pair_min_len, pe.pair_max_len = opt["U"], opt["d"], opt["D"]
        pe.pair_strand = 1 if "E" in opt else 0
    return p, pe


def load_reads(case, run):
    """Reads of a run in the reference's load order (PE: PE1/PE2 interleaved)."""
    files = [os.path.join(GOLD, case, f) for f in run["reads"]]
    n1, b1, o1 = po.read_fasta_reads(files[0])
    if len(files) == 1:
        return n1, b1, o1
    n2, b2, o2 = po.read_fasta_reads(files[1])
    names, reads = [], []
    for i in range(len(n1)):
        names += [n1[i], n2[i]]
        reads += [b1[o1[i]:o1[i + 1]], b2[o2[i]:o2[i + 1]]]
    bases, offs = po.pack_reads(reads)
    return names, bases, offs


def expected(case, tag):
    """Per read name: (nar_code, chrom, loci0, strand, mismatches|None, flag) from the reference's
    -M6 SAM (every read + YU:Z class) and -M0 CSV (mismatch count of accepted reads)."""
    exp = {}
    with gzip.open(os.path.join(GOLD, case, tag + ".sam.gz"), "rt") as f:
        for ln in f:
            if ln.startswith("@"):
                continue
            c = ln.rstrip("\n").split("\t")
            nar = "AA"
            for t in c[11:]:
                if t.startswith("YU:Z:"):
                    nar = t[5:]
            flag = int(c[1])
            if flag & 4:
                exp[c[0]] = (nar, None, None, None, None, flag)
            else:
                exp[c[0]] = (nar, c[2], int(c[3]) - 1, "-" if flag & 16 else "+", None, flag)
    with gzip.open(os.path.join(GOLD, case, tag + ".csv.gz"), "rt") as f:
        for row in csv.reader(f):
            name = row[13]
            e = exp[name]
            assert e[1] == row[3] and e[2] == int(row[4]) and e[3] == row[7], (name, e, row)
            exp[name] = (e[0], e[1], e[2], e[3], int(row[11]), e[5])
    return exp


def results_to_tuples(index_entries, names, res):
    """Same shape as expected() from a bkx_read_result array."""
    ent = {e.entry_id: e.name.decode() for e in index_entries}
    out = {}
    for nm, r in zip(names, res):
        code = abi.NAR_CODES[int(r["nar"])]
        if int(r["nar"]) == abi.NAR_ACCEPTED:
            out[nm] = (code, ent[int(r["chrom_id"])], int(r["match_loci"]), chr(int(r["strand"])), int(r["mismatches"]))
        else:
            out[nm] = (code, None, None, None, None)
    return out


def log_stats(case, tag):
    return open(os.path.join(GOLD, case, tag + ".log")).read()
