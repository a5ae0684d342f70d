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


def draw_filtered_pairing_case(seed, chroms):
    """Random paired-end reads, -U mode, insert range, -E and a random keep map (at least one chromosome filtered)."""
    import synth
    rng = np.random.default_rng(9000 + seed)
    L = int(rng.choice([50, 75, 100, 150]))
    ins_lo = int(rng.choice([L + 20, 150, 300]))
    usable = [c for c in chroms if len(c[1]) >= 1200]
    _, r1, _, r2 = synth.sim_reads(usable, int(rng.integers(400, 1000)), L, seed=int(rng.integers(1, 1 << 30)),
                                   subs=tuple(range(0, int(rng.integers(1, 5)))), junk_frac=0.08, n_frac=0.0, pe=True,
                                   insert=(ins_lo, ins_lo + int(rng.choice([50, 300, 900]))))
    # every seventh pair gets a mate from another pair: ends on different chromosomes, and orphans to recover
    r2 = [r2[(i + 1) % len(r2)] if i % 7 == 0 else r2[i] for i in range(len(r2))]
    reads = [x for pair in zip(r1, r2) for x in pair]
    bases = np.concatenate(reads)
    offs = np.arange(len(reads) + 1, dtype=np.uint64) * L
    pe = abi.PEParams()
    pe.pe_proc = 1 + seed % 4
    pe.pair_min_len = int(rng.choice([100, 150, 200]))
    pe.pair_max_len = pe.pair_min_len + int(rng.choice([300, 800, 1500]))
    pe.pair_strand = int(rng.random() < 0.15)
    keep = (rng.random(len(chroms) + 1) < 0.6).astype(np.uint8)
    keep[1 + int(rng.integers(0, len(chroms)))] = 0
    kw = dict(max_subs=int(rng.choice([3, 5, 8])), min_edit_dist=1)
    return bases, offs, pe, keep, kw


@pytest.mark.parametrize("seed", range(12))
def test_random_pairing_with_keep_maps_matches_oracle(seed, golden_dir):
    """The pairing and orphan-recovery kernels alone: the GPU's own alignment records go into both the CUDA pairing
    (bkx_pair_reads with a keep map set) and the oracle's (bko_pair_reads_filtered); records, counters and the insert-length
    histogram must agree.  Random modes -U1..4, insert ranges, -E, keep maps, swapped mates."""
    from test_gpu_fuzz import index_pair
    case = ["tiny", "repeats", "lowcopy"][seed % 3]
    gidx, oidx, chroms = index_pair(case, golden_dir)
    bases, offs, pe, keep, kw = draw_filtered_pairing_case(seed, chroms)
    p = gidx.default_params(0, **kw)
    rec, _ = gidx.align(p, bases, offs)
    exp = rec.copy()
    ld_o = np.zeros(100001, dtype=np.uint32)
    ost = oidx.pair(p, pe, exp, bases, offs, len_dist=ld_o, keep=keep)
    ld_g = np.zeros(100001, dtype=np.uint32)
    try:
        gidx.set_chrom_filter(keep)
        gst = gidx.pair(p, pe, rec, bases, offs, len_dist=ld_g)
    finally:
        gidx.set_chrom_filter(None)
    assert_same(None, rec, exp)
    assert bytes(gst) == bytes(ost)
    assert (ld_g == ld_o).all()


@pytest.mark.parametrize("seed", range(8))
def test_orphan_recovery_with_core_lengths_from_a_lower_rate_matches_oracle(seed, golden_dir):
    """-6 runs: search and acceptance at -s plus -6, the recovery's core lengths from -s (bkx_pe_params.rescue_core_subs_p1;
    Aligner.cpp:3256 against :3275).  Wide insert ranges, so that the suffix-array seeded branch of the recovery runs (on these draws the
    oracle recovers up to 68 fewer pairs with the field than without: the cases are sensitive to it)."""
    from test_gpu_fuzz import index_pair
    case = ["tiny", "repeats", "lowcopy"][seed % 3]
    gidx, oidx, chroms = index_pair(case, golden_dir)
    bases, offs, pe, _, kw = draw_filtered_pairing_case(100 + seed, chroms)
    pe.pe_proc = 1 if seed & 1 else 3
    pe.pair_max_len = pe.pair_min_len + 1400
    s = [2, 3, 5][seed % 3]
    pe.rescue_core_subs_p1 = s + 1
    p = gidx.default_params(0, max_subs=s + [1, 3, 5][(seed // 3) % 3], min_edit_dist=kw["min_edit_dist"])
    rec, _ = gidx.align(p, bases, offs)
    exp = rec.copy()
    ost = oidx.pair(p, pe, exp, bases, offs)
    gst = gidx.pair(p, pe, rec, bases, offs)
    assert_same(None, rec, exp)
    assert bytes(gst) == bytes(ost)


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


@pytest.mark.parametrize("tag", ["p6u1", "p6u3sam"])
def test_cli_primer_correction_with_orphan_recovery_matches_reference(tag, golden_dir, tmp_path):
    """bkx-align -6 with -U1 / -U3 and a wide insert range (the run that needs rescue_core_subs_p1) against the reference's
    files; the CPU suite runs the same check through the oracle-backed test double."""
    cli.test_cli_read_sampling_matches_reference(tag, golden_dir, tmp_path)
