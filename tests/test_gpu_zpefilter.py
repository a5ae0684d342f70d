"""GPU: -Z / -z chromosome filters in paired-end runs.  In the reference they act inside the pairing (AcceptThisChromID,
Aligner.cpp:2651-2710, called from AcceptProvPE :2771-2786, the orphan-recovery arms :3296-3302 / :3411-3417 and the SE
fallback :3442-3477); here the pairing kernels take them as a keep map (bkx_set_chrom_filter).  Checked against the oracle
(bko_pair_reads_filtered) and against five runs of the reference binary (tests/golden/pefilter)."""
import json
import os
import re

import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
import test_gpu_cli as cli
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx
from test_gpu_parity import assert_same, indexes

pytestmark = pytest.mark.gpu

PE_FILTER_RUNS = ["U1_Z2", "U2_z13", "U3_ZZ", "U4_zZ", "U1far_z2"]


def keep_maps(ents, args):
    """(AcceptThisChromID, FiltByChroms) per entry id: exclude-first inside the pairing, include-first afterwards."""
    excl = [a[2:] for a in args if a.startswith("-Z")]
    incl = [a[2:] for a in args if a.startswith("-z")]
    hit = lambda pats, name: any(re.search(x, name, re.I) for x in pats)
    keep = np.ones(len(ents) + 1, dtype=np.uint8)
    post = np.ones(len(ents) + 1, dtype=np.uint8)
    for e in ents:
        nm = e.name.decode()
        keep[e.entry_id] = (not hit(excl, nm)) and (not incl or hit(incl, nm))
        post[e.entry_id] = hit(incl, nm) or (not incl and not hit(excl, nm))
    return keep, post


@pytest.mark.parametrize("tag", PE_FILTER_RUNS)
def test_pairing_kernels_with_chromosome_filters_match_oracle_and_reference(tag, golden_dir):
    run = json.load(open(os.path.join(gu.GOLD, "pefilter", "runs.json")))[tag]
    gidx, oidx = indexes("tiny", golden_dir)
    plain = [a for a in run["args"] if a[1] not in "Zz"]
    p, pe = gu.params_from_args(gidx, plain)
    names, bases, offs = gu.load_reads("tiny", run)
    keep, post = keep_maps(gidx.entries(), run["args"])
    exp, _ = oidx.align(p, bases, offs, nthreads=4)
    ope = oidx.pair(p, pe, exp, bases, offs, keep=keep)
    try:
        gidx.set_chrom_filter(keep)
        got, _ = gidx.align(p, bases, offs)                       # align, then pair (bkx_pair_reads)
        gpe = gidx.pair(p, pe, got, bases, offs)
        assert_same(names, got, exp)
        assert bytes(gpe) == bytes(ope)
        fused, _, fpe = gidx.align_pairs(p, pe, bases, offs)      # one pass (bkx_align_pairs)
        assert_same(names, fused, exp)
        assert bytes(fpe) == bytes(ope)
    finally:
        gidx.set_chrom_filter(None)
    # cleared: the pairing is the unfiltered one again
    exp0, _ = oidx.align(p, bases, offs, nthreads=4)
    ope0 = oidx.pair(p, pe, exp0, bases, offs)
    got0, _, gpe0 = gidx.align_pairs(p, pe, bases, offs)
    assert_same(names, got0, exp0)
    assert bytes(gpe0) == bytes(ope0)
    # and the filtered records straight against the reference's files, FiltByChroms applied as the host does
    drop = (got["nar"] == abi.NAR_ACCEPTED) & (post[got["chrom_id"]] == 0)
    got["nar"][drop] = abi.NAR_CODES.index("FC")
    tup = gu.results_to_tuples(gidx.entries(), names, got)
    ref = gu.expected("pefilter", tag)
    diff = [(n, tup[n], ref[n][:5]) for n in names if tup[n] != ref[n][:5]]
    assert not diff, "%d reads differ from the reference, first %r" % (len(diff), diff[:5])
    m = re.search(r"(\d+) Paired End aligned pairs were filtered out by chromosome", gu.log_stats("pefilter", tag))
    assert m and int(m.group(1)) == gpe.num_filtered_by_chrom


def test_chromosome_filter_argument_checks(golden_dir):
    gidx, _ = indexes("tiny", golden_dir)
    with pytest.raises(bkx.BkxError):
        gidx.set_chrom_filter(np.ones(len(gidx.entries()) + 5, dtype=np.uint8))   # one flag per id 0..num_entries, no more
    gidx.set_chrom_filter(None)


@pytest.mark.parametrize("tag", PE_FILTER_RUNS)
def test_cli_paired_end_chromosome_filters_match_reference(tag, golden_dir, tmp_path):
    """bkx-align -Z / -z with -U1..4: CSV rows, SAM records and the summary block (incl. 'pairs were filtered out by
    chromosome' and FiltByChroms' own count) against the reference's files.  The CPU suite runs the same check through the
    oracle-backed test double (tests/test_host_cli_cpu.py)."""
    cli.test_cli_outputs_match_reference("pefilter", tag, golden_dir, tmp_path, src_case="tiny")
