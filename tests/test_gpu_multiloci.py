"""GPU: multi-loci options on the low-copy repeat genome (tests/golden/lowcopy, reads hitting 2..9 equally good loci).
-r5 (every locus returned, bkx_align_reads_multi) against the oracle field by field, hit lists included and in the
order LocateCoreMultiples discovers them (libbiokanga/SfxArrayV2.cpp:6157-6205); -r1 counts against the oracle.
Run the file again with BKX_NO_FAST=1 to put the general kernel on the same inputs (done in the round's GPU runs)."""
import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lowcopy(golden_dir):
    sfx = gu.sfx_path("lowcopy", golden_dir)
    run = gu.runs("lowcopy")["r0_s3"]
    names, bases, offs = gu.load_reads("lowcopy", run)
    return bkx.Index.open(sfx), po.OracleIndex(sfx), bases, offs


@pytest.mark.parametrize("max_subs,mmd,limit,clamp", [(3, 1, 5, 0), (3, 1, 3, 1), (5, 2, 8, 0), (3, 1, 2, 0), (10, 1, 64, 0),
                                                     (3, 1, 9, 1)])
def test_all_loci_match_oracle(lowcopy, max_subs, mmd, limit, clamp):
    gidx, oidx, bases, offs = lowcopy
    kw = dict(max_subs=max_subs, min_edit_dist=mmd, ml_mode=5, max_ml_matches=limit, clamp_max_ml=clamp)
    got, gm, gst = gidx.align_multi(gidx.default_params(0, **kw), bases, offs)
    exp, em, est = oidx.align_multi(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    acc = got["nar"] == abi.NAR_ACCEPTED
    # slots [0, num_hits) of accepted reads are defined (the rest may hold candidates of an abandoned higher level)
    valid = acc[:, None] & (np.arange(limit)[None, :] < got["num_hits"][:, None])
    assert gm[valid].tobytes() == em[valid].tobytes()
    assert gst.as_dict() == est.as_dict()
    assert (got["num_hits"][acc] > 1).sum() > 100        # the regime: many reads with several loci
    assert got["num_hits"][acc].max() <= limit
    # every returned locus really carries the read with the stated mismatches
    seq = np.array(oidx.seq())
    ents = {e.entry_id: e for e in oidx.entries()}
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    for i in np.nonzero(acc)[0][:400]:
        rd = bases[offs[i]:offs[i + 1]]
        for h in gm[i][:got["num_hits"][i]]:
            e = ents[int(h["chrom_id"])]
            g = seq[e.start_ofs + int(h["match_loci"]): e.start_ofs + int(h["match_loci"]) + len(rd)]
            q = comp[rd[::-1]] if chr(int(h["strand"])) == "-" else rd
            assert int((q != g).sum()) == int(h["mismatches"]) == int(got["low_mm"][i])


@pytest.mark.parametrize("mode,limit,clamp,max_subs", [(3, 5, 0, 3), (4, 5, 0, 3), (4, 8, 1, 5)])
def test_clustering_modes_match_oracle(lowcopy, mode, limit, clamp, max_subs):
    """-r3 / -r4: records and hit lists of the search equal the oracle's; the host-side assignment then gives the same
    records whichever side produced its input."""
    gidx, oidx, bases, offs = lowcopy
    kw = dict(max_subs=max_subs, ml_mode=mode, max_ml_matches=limit, clamp_max_ml=clamp)
    got, gm, gst = gidx.align_multi(gidx.default_params(0, **kw), bases, offs)
    exp, em, est = oidx.align_multi(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert gst.as_dict() == est.as_dict()
    multi = (got["nar"] == abi.NAR_MULTIALIGN) & (got["hit_rslt"] == 1)
    cnt = np.where(multi, got["low_hit_instances"], np.where(got["nar"] == abi.NAR_ACCEPTED, 1, 0))
    valid = np.arange(limit)[None, :] < cnt[:, None]
    assert gm[valid].tobytes() == em[valid].tobytes()
    assert multi.sum() > 100
    a, b = got.copy(), exp.copy()
    ca = bkx.assign_multi_matches(a, gm, mode, 100)
    cb = bkx.assign_multi_matches(b, em, mode, 100)
    assert a.tobytes() == b.tobytes() and bytes(ca) == bytes(cb)


def test_all_loci_mode_needs_its_own_call(lowcopy):
    gidx, oidx, bases, offs = lowcopy
    with pytest.raises(bkx.BkxError):
        gidx.align(gidx.default_params(0, ml_mode=5, max_ml_matches=5), bases, offs)
    with pytest.raises(bkx.BkxError):
        gidx.align_multi(gidx.default_params(0, ml_mode=5, max_ml_matches=501), bases, offs)
    with pytest.raises(bkx.BkxError):
        gidx.align(gidx.default_params(0, ml_mode=3, max_ml_matches=5), bases, offs)
    with pytest.raises(bkx.BkxError):
        gidx.align(gidx.default_params(0, ml_mode=2, max_ml_matches=5), bases, offs)


@pytest.mark.parametrize("limit,clamp", [(5, 0), (2, 0), (4, 1), (20, 0)])
def test_distribution_mode_matches_oracle(lowcopy, limit, clamp):
    gidx, oidx, bases, offs = lowcopy
    kw = dict(max_subs=3, ml_mode=1, max_ml_matches=limit, clamp_max_ml=clamp)
    got, gst = gidx.align(gidx.default_params(0, **kw), bases, offs)
    exp, est = oidx.align(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert gst.as_dict() == est.as_dict()
    assert gst.tot_accepted_multi > 100 and gst.tot_loci_aligned > gst.tot_accepted_aligned


@pytest.mark.parametrize("mode,limit,max_subs,mmd,strand", [(5, 5, 3, 1, 0), (5, 3, 8, 2, 0), (5, 8, 5, 1, 1), (5, 2, 0, 1, 0), (5, 64, 10, 1, 0),
                                                            (3, 5, 3, 1, 0), (4, 6, 5, 1, 2), (1, 4, 5, 1, 0), (1, 2, 0, 1, 0), (1, 300, 8, 1, 0)])
def test_best_matches_match_oracle(lowcopy, mode, limit, max_subs, mmd, strand):
    """-N (best_matches = 1, CSfxArrayV3::LocateBestMatches): one un-staged pass that keeps the `limit` loci with the fewest
    mismatches, equal ones in discovery order -- records, hit lists (where the mode returns them), seeds / cands and stats
    against the oracle, whose restatement is pinned on eight reference runs (tests/golden/bestmatches)."""
    gidx, oidx, bases, offs = lowcopy
    kw = dict(max_subs=max_subs, min_edit_dist=mmd, ml_mode=mode, max_ml_matches=limit, best_matches=1, align_strand=strand)
    if mode == 1:
        got, gst = gidx.align(gidx.default_params(0, **kw), bases, offs)
        exp, est = oidx.align(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    else:
        got, gm, gst = gidx.align_multi(gidx.default_params(0, **kw), bases, offs)
        exp, em, est = oidx.align_multi(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    for f in abi.RESULT_DTYPE.names:
        bad = np.nonzero(got[f] != exp[f])[0]
        assert len(bad) == 0, (f, int(bad[0]), got[bad[0]], exp[bad[0]])
    assert gst.as_dict() == est.as_dict()
    if mode != 1:
        hits = got["hit_rslt"] == 1
        cnt = np.where(hits & (got["nar"] == abi.NAR_ACCEPTED), np.maximum(got["num_hits"], 1),
                       np.where(hits & (got["nar"] == abi.NAR_MULTIALIGN), got["low_hit_instances"], 0))
        valid = np.arange(limit)[None, :] < cnt[:, None]
        assert gm[valid].tobytes() == em[valid].tobytes()
        # the kept loci come sorted by mismatches
        mmv = np.where(valid, gm["mismatches"].astype(np.int32), 255)
        assert bool((np.diff(mmv, axis=1) >= 0).all())
    assert int((got["hit_rslt"] == 1).sum()) > 1000


@pytest.mark.parametrize("mode,limit,clamp,best,max_subs", [(5, 300, 1, 0, 5), (5, 500, 0, 0, 5), (5, 100, 0, 1, 5), (4, 200, 1, 0, 5),
                                                            (1, 500, 0, 0, 5), (5, 64, 1, 0, 5)])
def test_hundreds_of_loci_match_oracle(golden_dir, mode, limit, clamp, best, max_subs):
    """-R up to 500 on the `repeats` genome: 60 bp reads with dozens to ~2850 equally good loci (tests/golden/manyloci)."""
    import os
    sfx = gu.sfx_path("repeats", golden_dir)
    gidx, oidx = bkx.Index.open(sfx), po.OracleIndex(sfx)
    names, bases, offs = po.read_fasta_reads(os.path.join(gu.GOLD, "manyloci", "r60a.fa.gz"))
    kw = dict(max_subs=max_subs, ml_mode=mode, max_ml_matches=limit, clamp_max_ml=clamp, best_matches=best)
    if mode == 1:
        got, gst = gidx.align(gidx.default_params(0, **kw), bases, offs)
        exp, est = oidx.align(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    else:
        got, gm, gst = gidx.align_multi(gidx.default_params(0, **kw), bases, offs)
        exp, em, est = oidx.align_multi(oidx.default_params(0, **kw), bases, offs, nthreads=4)
    for f in abi.RESULT_DTYPE.names:
        bad = np.nonzero(got[f] != exp[f])[0]
        assert len(bad) == 0, (f, len(bad), int(bad[0]), names[int(bad[0])], got[bad[0]], exp[bad[0]])
    assert gst.as_dict() == est.as_dict()
    if mode != 1:
        hits = got["hit_rslt"] == 1
        cnt = np.where(hits, got["low_hit_instances"], 0)
        valid = np.arange(limit)[None, :] < cnt[:, None]
        assert gm[valid].tobytes() == em[valid].tobytes()
        assert int(cnt.max()) > 64 or limit <= 64
    gidx.close()
