"""CPU: the oracle (oracle/bk_oracle.c) against every golden vector produced by the reference binary."""
import re

import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
from biokanga_b200 import abi


@pytest.mark.parametrize("case,tag", gu.all_runs())
def test_oracle_matches_reference(case, tag, golden_dir):
    run = gu.runs(case)[tag]
    idx = po.OracleIndex(gu.sfx_path(case, golden_dir))
    p, pe = gu.params_from_args(idx, run["args"])
    names, bases, offs = gu.load_reads(case, run)
    if run.get("all_loci"):
        return check_all_loci(case, tag, idx, p, names, bases, offs)
    if run.get("clustered"):
        return check_clustered(case, tag, idx, p, names, bases, offs)
    res, st = idx.align(p, bases, offs, nthreads=4)
    if pe is not None:
        idx.pair(p, pe, res, bases, offs)
    got = gu.results_to_tuples(idx.entries(), names, res)
    exp = gu.expected(case, tag)
    assert len(got) == len(exp)
    bad = [(n, got[n], exp[n][:5]) for n in names if got[n] != exp[n][:5]]
    assert not bad, "%d reads differ, first: %r" % (len(bad), bad[:5])
    # the NAR histogram of the reference's log must match too (EN is reported via NumSloughedNs)
    log = gu.log_stats(case, tag)
    hist = np.bincount(res["nar"], minlength=abi.NAR_COUNT)
    for code, n in re.findall(r"^\s+(\d+) \((\w\w)\)", log, flags=re.M):
        assert hist[abi.NAR_CODES.index(n)] == int(code), (n, code, hist)
    # "Provisionally accepted A aligned reads (U uniquely, M aligning to multiloci) aligning to a total of T loci":
    # with -r1 these carry the exact per-read LowHitInstances (Aligner.cpp:9357-9364, :535)
    m = re.search(r"Provisionally accepted (\d+) aligned reads \((\d+) uniquely, (\d+) aligning to multiloci\) "
                  r"aligning to a total of (\d+) loci", log)
    assert m, "summary line missing"
    assert (st.tot_accepted_aligned, st.tot_accepted_unique, st.tot_accepted_multi, st.tot_loci_aligned) == tuple(
        int(v) for v in m.groups())


def check_all_loci(case, tag, idx, p, names, bases, offs):
    """-r5 runs: one CSV row per reported locus, numbered in read order (the fixtures were made with one thread)."""
    import csv
    import gzip
    import os
    res, multi, st = idx.align_multi(p, bases, offs, nthreads=4)
    ents = {e.entry_id: e.name.decode() if isinstance(e.name, bytes) else e.name for e in idx.entries()}
    got = []
    for i in range(len(names)):
        if res["nar"][i] == abi.NAR_ACCEPTED:
            for h in multi[i][:res["num_hits"][i]]:
                got.append((names[i], ents[int(h["chrom_id"])], int(h["match_loci"]), chr(int(h["strand"])), int(h["mismatches"])))
    exp = {}
    with gzip.open(os.path.join(gu.GOLD, case, tag + ".csv.gz"), "rt") as f:
        for row in csv.reader(f):
            exp[int(row[0])] = (row[13], row[3], int(row[4]), row[7], int(row[11]))
    assert len(got) == len(exp)
    assert got == [exp[k] for k in sorted(exp)]          # same loci, same order, same numbering
    log = gu.log_stats(case, tag)
    m = re.search(r"Provisionally accepted (\d+) aligned reads \((\d+) uniquely, (\d+) aligning to multiloci\) "
                  r"aligning to a total of (\d+) loci", log)
    assert (st.tot_accepted_aligned, st.tot_accepted_unique, st.tot_accepted_multi, st.tot_loci_aligned) == tuple(
        int(v) for v in m.groups())
    m = re.search(r"Unable to align (\d+) source reads of which (\d+)", log)
    assert (st.tot_non_aligned + st.num_sloughed_ns, st.num_sloughed_ns) == tuple(int(v) for v in m.groups())


def check_clustered(case, tag, idx, p, names, bases, offs):
    """-r3 / -r4 runs: the oracle's records and hit lists go through the PRODUCT's host-side clustering
    (bkx_assign_multi_matches, no GPU involved) and must give the reference's CSV rows and log counters."""
    from biokanga_b200 import lib as bkx
    res, multi, st = idx.align_multi(p, bases, offs, nthreads=4)
    max_len = int((offs[1:] - offs[:-1]).max())
    cs = bkx.assign_multi_matches(res, multi, p.ml_mode, max_len)
    got = gu.results_to_tuples(idx.entries(), names, res)
    exp = gu.expected(case, tag)
    assert len(got) == len(exp)
    bad = [(n, got[n], exp[n][:5]) for n in names if got[n] != exp[n][:5]]
    assert not bad, "%d reads differ, first: %r" % (len(bad), bad[:5])
    log = gu.log_stats(case, tag)
    m = re.search(r"Assigning (\d+) reads which aligned to multiple loci", log)
    assert cs.multi_reads == int(m.group(1))
    m = re.search(r"removed (\d+) unclustered orphans from (\d+) putative resulting in (\d+) \((\d+) clustered near unique, "
                  r"(\d+) clustered near other multiloci reads\)", log)
    assert (cs.putative - cs.assigned, cs.putative, cs.assigned, cs.near_unique, cs.near_multi) == tuple(
        int(v) for v in m.groups())


@pytest.mark.parametrize("tag", ["U1_Z2", "U2_z13", "U3_ZZ", "U4_zZ", "U1far_z2"])
def test_oracle_pairing_with_chromosome_filters(tag, golden_dir):
    """-Z / -z in paired-end runs act inside the pairing (AcceptThisChromID: exclude expressions first, then the include
    ones) and once more afterwards (FiltByChroms: include expressions first).  Pins bko_pair_reads_filtered -- the oracle is
    ahead of the CUDA pairing kernel here, which does not take the filter yet (bkx-align refuses the combination)."""
    import json
    import os
    run = json.load(open(os.path.join(gu.GOLD, "pefilter", "runs.json")))[tag]
    excl = [a[2:] for a in run["args"] if a.startswith("-Z")]
    incl = [a[2:] for a in run["args"] if a.startswith("-z")]
    idx = po.OracleIndex(gu.sfx_path("tiny", golden_dir))
    p, pe = gu.params_from_args(idx, [a for a in run["args"] if a[1] not in "Zz"])
    names, bases, offs = gu.load_reads("tiny", run)
    ents = idx.entries()
    hit = lambda pats, name: any(re.search(x, name, re.I) for x in pats)
    keep = np.ones(len(ents) + 1, dtype=np.uint8)            # AcceptThisChromID, Aligner.cpp:2651-2710
    post = np.ones(len(ents) + 1, dtype=np.uint8)            # FiltByChroms, Aligner.cpp:4019-4124
    for e in ents:
        nm = e.name.decode()
        keep[e.entry_id] = (not hit(excl, nm)) and (not incl or hit(incl, nm))
        post[e.entry_id] = hit(incl, nm) or (not incl and not hit(excl, nm))
    res, st = idx.align(p, bases, offs, nthreads=4)
    pst = idx.pair(p, pe, res, bases, offs, keep=keep)
    drop = (res["nar"] == abi.NAR_ACCEPTED) & (post[res["chrom_id"]] == 0)
    res["nar"][drop] = abi.NAR_CODES.index("FC")
    got = gu.results_to_tuples(ents, names, res)
    exp = gu.expected("pefilter", tag)
    bad = [(n, got[n], exp[n][:5]) for n in names if got[n] != exp[n][:5]]
    assert not bad, "%d reads differ, first: %r" % (len(bad), bad[:5])
    log = gu.log_stats("pefilter", tag)
    hist = np.bincount(res["nar"], minlength=abi.NAR_COUNT)
    for code, n in re.findall(r"^\s+(\d+) \((\w\w)\)", log, flags=re.M):
        assert hist[abi.NAR_CODES.index(n)] == int(code), (n, code, hist)
    m = re.search(r"(\d+) Paired End aligned pairs were filtered out by chromosome", log)
    assert m and int(m.group(1)) == pst.num_filtered_by_chrom
    m = re.search(r"From \d+ Paired End pairs there were (\d+) accepted \(of which (\d+) pairs were from recovered orphans\)", log)
    assert m and (int(m.group(1)), int(m.group(2))) == (pst.accepted_num_paired, pst.partner_paired)
    m = re.search(r"Filtering by chromosome completed - removed (\d+)  matches", log)
    assert m and int(m.group(1)) == int(drop.sum())
