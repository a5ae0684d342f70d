"""GPU: randomised differential test.  Random option combinations (-s, -e, -m, -Q, -n, multi-loci modes with -R / -X)
and random read sets (lengths 16..400, fixed or ragged, substitutions, Ns, junk, chromosome-spanning reads) on the three
golden genomes: every record of the fast + general kernel pair, of the general kernel alone and (default multi-loci mode)
of the wave path against the oracle."""
import os

import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
import synth
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu

_IDX = {}


def index_pair(case, golden_dir):
    if case not in _IDX:
        sfx = gu.sfx_path(case, golden_dir)
        oi = po.OracleIndex(sfx)
        seq = np.array(oi.seq())
        chroms = [(e.name.decode() if isinstance(e.name, bytes) else e.name, seq[e.start_ofs:e.end_ofs + 1]) for e in oi.entries()]
        _IDX[case] = (bkx.Index.open(sfx), oi, chroms)
    return _IDX[case]


def random_config(rng):
    kw = dict(max_subs=int(rng.choice([0, 1, 2, 3, 3, 5, 7, 10, 15])), min_edit_dist=int(rng.choice([1, 1, 2])),
              max_ns=int(rng.choice([0, 1, 1, 2, 5])), align_strand=int(rng.choice([0, 0, 1, 2])))
    mode = int(rng.choice([0, 0, 0, 1, 3, 4, 5]))
    if mode:
        kw.update(ml_mode=mode, max_ml_matches=int(rng.choice([2, 3, 5, 8, 20, 64])), clamp_max_ml=int(rng.integers(0, 2)))
    return int(rng.integers(0, 4)), kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("BKX_FUZZ_SEEDS", "60"))))
def test_random_options_and_reads(seed, golden_dir):
    rng = np.random.default_rng(1000 + seed)
    case = ["tiny", "repeats", "lowcopy"][seed % 3]
    gidx, oidx, chroms = index_pair(case, golden_dir)
    pmode, kw = random_config(rng)
    reads = []
    for _ in range(int(rng.integers(1, 4))):  # up to three length classes in one batch
        L = int(rng.choice([16, 25, 36, 50, 64, 75, 100, 101, 150, 151, 250, 319, 320, 321, 400]))
        usable = [c for c in chroms if len(c[1]) >= L]
        _, r = synth.sim_reads(usable, int(rng.integers(200, 900)), L, seed=int(rng.integers(1, 1 << 30)),
                               subs=tuple(range(0, int(rng.integers(1, 9)))), junk_frac=0.03, n_frac=0.05, boundary_frac=0.02)
        reads += r
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    bases = np.concatenate(reads)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    multi_mode = kw.get("ml_mode", 0) >= 3
    op = oidx.default_params(pmode, **kw)
    if multi_mode:
        exp, em, est = oidx.align_multi(op, bases, offs, nthreads=4)
    else:
        exp, est = oidx.align(op, bases, offs, nthreads=4)
    for general_only in (False, True):
        if general_only:
            os.environ["BKX_NO_FAST"] = "1"
        try:
            gp = gidx.default_params(pmode, **kw)
            if multi_mode:
                got, gm, gst = gidx.align_multi(gp, bases, offs)
            else:
                got, gst = gidx.align(gp, bases, offs)
        finally:
            os.environ.pop("BKX_NO_FAST", None)
        for f in abi.RESULT_DTYPE.names:
            bad = np.nonzero(got[f] != exp[f])[0]
            assert len(bad) == 0, (case, pmode, kw, general_only, f, int(bad[0]), got[bad[0]], exp[bad[0]])
        assert gst.as_dict() == est.as_dict(), (case, pmode, kw, general_only)
        if multi_mode:
            hits = (got["hit_rslt"] == 1)
            cnt = np.where(hits & (got["nar"] == abi.NAR_ACCEPTED), got["num_hits"],
                           np.where(hits & (got["nar"] == abi.NAR_MULTIALIGN), got["low_hit_instances"], 0))
            valid = np.arange(gm.shape[1])[None, :] < cnt[:, None]
            assert gm[valid].tobytes() == em[valid].tobytes(), (case, pmode, kw, general_only)
    if kw.get("ml_mode", 0) == 0:
        # the default search once more through the wave path (bkx_wave.cuh; large launches take it on their own): the compact
        # host call with the path forced on -- 16-byte records carry no per-read seed / candidate counts, the statistics
        # carry their sums
        os.environ["BKX_WAVE"] = "1"
        try:
            got, gst = gidx.align_packed2(gidx.default_params(pmode, **kw), bases, offs)
        finally:
            os.environ.pop("BKX_WAVE", None)
        for f in abi.RESULT_DTYPE.names:
            if f in ("seeds", "cands", "reserved"):
                continue
            bad = np.nonzero(got[f] != exp[f])[0]
            assert len(bad) == 0, (case, pmode, kw, "wave", f, int(bad[0]), got[bad[0]], exp[bad[0]])
        assert gst.as_dict() == est.as_dict(), (case, pmode, kw, "wave")


@pytest.mark.parametrize("seed", range(int(os.environ.get("BKX_FUZZ_PE_SEEDS", "30"))))
def test_random_paired_end_options(seed, golden_dir):
    """Paired ends: random -U mode, insert range, -E, read length and substitutions; the fused host call (align + pair
    + orphan recovery per pipeline slice, tiny slices forced on odd seeds) against the oracle's align then pair."""
    rng = np.random.default_rng(5000 + seed)
    case = ["tiny", "repeats", "lowcopy"][seed % 3]
    gidx, oidx, chroms = index_pair(case, golden_dir)
    L = int(rng.choice([36, 50, 75, 100, 125, 150]))
    ins_lo = int(rng.choice([L, L + 20, 150, 200, 300]))
    usable = [c for c in chroms if len(c[1]) >= 1200]
    n1, r1, n2, r2 = synth.sim_reads(usable, int(rng.integers(300, 1200)), L, seed=int(rng.integers(1, 1 << 30)),
                                     subs=tuple(range(0, int(rng.integers(1, 7)))), junk_frac=0.08, n_frac=0.0, pe=True,
                                     insert=(max(ins_lo, L), max(ins_lo, L) + int(rng.choice([50, 300, 900]))))
    reads = [x for pair in zip(r1, r2) for x in pair]
    bases = np.concatenate(reads)
    offs = np.arange(len(reads) + 1, dtype=np.uint64) * L
    kw = dict(max_subs=int(rng.choice([2, 3, 5, 8])), min_edit_dist=int(rng.choice([1, 2])))
    pe = abi.PEParams()
    pe.pe_proc = int(rng.integers(1, 5))
    pe.pair_min_len = int(rng.choice([25, 100, 150, 200, 300]))
    pe.pair_max_len = pe.pair_min_len + int(rng.choice([100, 500, 800, 1500, 4000]))
    pe.pair_strand = int(rng.random() < 0.15)
    op = oidx.default_params(0, **kw)
    exp, _ = oidx.align(op, bases, offs, nthreads=4)
    ld = np.zeros(100001, dtype=np.uint32)
    ost = oidx.pair(op, pe, exp, bases, offs, len_dist=ld)
    if seed & 1:
        os.environ["BKX_SLICE_MIN"], os.environ["BKX_SLICE_MAX"] = "1024", "3000"
    try:
        packed = bool(seed & 2)
        ld2 = np.zeros(100001, dtype=np.uint32)
        got, _, gps = gidx.align_pairs(gidx.default_params(0, **kw), pe, bkx.pack_bases4(bases) if packed else bases, offs,
                                       packed=packed, len_dist=ld2)
    finally:
        os.environ.pop("BKX_SLICE_MIN", None)
        os.environ.pop("BKX_SLICE_MAX", None)
    for f in abi.RESULT_DTYPE.names:
        bad = np.nonzero(got[f] != exp[f])[0]
        if len(bad):
            i = int(bad[0])
            plain, _ = gidx.align(gidx.default_params(0, **kw), bases, offs)
            again, _, _ = gidx.align_pairs(gidx.default_params(0, **kw), pe, bases, offs)
            oalign, _ = oidx.align(op, bases, offs, nthreads=1)
            import subprocess
            print("DIAG GPU", subprocess.run(["nvidia-smi", "--query-gpu=serial,uuid,driver_version", "--format=csv,noheader"],
                                             capture_output=True, text=True).stdout.strip())
            print("DIAG self-check of the live index: %d elements outside their bucket" % gidx.self_check())
            os.environ["BKX_NO_FAST"] = "1"
            gen_only, _ = gidx.align(gidx.default_params(0, **kw), bases, offs)
            os.environ.pop("BKX_NO_FAST")
            fresh_idx = bkx.Index.open(gu.sfx_path(case, golden_dir))
            fresh, _ = fresh_idx.align(fresh_idx.default_params(0, **kw), bases, offs)
            fresh_idx.close()
            def nbad(a):
                return int(sum((a[q] != oalign[q]).sum() for q in ("nar", "cands", "nxt_low_mm", "match_loci")))
            print("DIAG mismatching fields vs oracle align: old index fast+general %d, old index general only %d, FRESH index object %d" % (
                nbad(plain), nbad(gen_only), nbad(fresh)))
            print("DIAG info old index: prefix_k %d device_bytes %d" % (gidx.info.prefix_k, gidx.info.device_bytes))
            for what, label in ((1, "overflow pool"), (2, "lane hash sets"), (4, "launch geometry")):
                gidx.debug_reset(what)
                again2, _ = gidx.align(gidx.default_params(0, **kw), bases, offs)
                print("DIAG after resetting %s of the old index: %d mismatching fields" % (label, nbad(again2)))
            print("DIAG field %s read %d: fused %r, fused again %r, plain GPU align %r, oracle align only %r, oracle align+pair %r" % (
                f, i, got[f][i], again[f][i], plain[f][i], oalign[f][i], exp[f][i]))
            raise AssertionError("%s L=%d %r U%d d%d D%d E%d field %s read %d (mate %d)\n got  %r\n exp  %r\n mate got %r\n mate exp %r" % (
                case, L, kw, pe.pe_proc, pe.pair_min_len, pe.pair_max_len, pe.pair_strand, f, i, i ^ 1, got[i], exp[i],
                got[i ^ 1], exp[i ^ 1]))
    assert bytes(gps) == bytes(ost)
    assert np.array_equal(ld, ld2)
