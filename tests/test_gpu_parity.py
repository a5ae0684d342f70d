"""GPU: the CUDA path, called through the C ABI (libbkx.so), against the CPU oracle on the same
inputs and against the golden vectors made by the reference binary.  Bit-exact, every field."""
import ctypes as C

import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu

FIELDS = ["nar", "hit_rslt", "strand", "num_hits", "low_mm", "nxt_low_mm", "low_hit_instances", "chrom_id",
          "match_loci", "match_len", "mismatches", "flags", "seeds", "cands"]


def assert_same(names, got, exp, fields=FIELDS):
    bad = np.zeros(len(got), dtype=bool)
    for f in fields:
        bad |= got[f] != exp[f]
    if bad.any():
        idx = np.nonzero(bad)[0]
        msg = ["%d of %d reads differ" % (len(idx), len(got))]
        for i in idx[:8]:
            msg.append("read %d %s\n   cuda   %s\n   oracle %s" % (i, names[i] if names else "", got[i], exp[i]))
        raise AssertionError("\n".join(msg))


_CACHE = {}


def indexes(case, golden_dir):
    if case not in _CACHE:
        path = gu.sfx_path(case, golden_dir)
        _CACHE[case] = (bkx.Index.open(path), po.OracleIndex(path))
    return _CACHE[case]


@pytest.mark.parametrize("case,tag", gu.all_runs(all_loci=False))
def test_cuda_matches_oracle_and_reference(case, tag, golden_dir):
    run = gu.runs(case)[tag]
    gidx, oidx = indexes(case, golden_dir)
    p, pe = gu.params_from_args(gidx, run["args"])
    op, _ = gu.params_from_args(oidx, run["args"])
    assert bytes(p) == bytes(op), "default parameter derivation differs"
    names, bases, offs = gu.load_reads(case, run)
    got, gst = gidx.align(p, bases, offs)
    exp, ost = oidx.align(op, bases, offs, nthreads=4)
    assert_same(names, got, exp)
    assert gst.as_dict() == ost.as_dict()
    if pe is not None:
        gpe = gidx.pair(p, pe, got, bases, offs)
        ope = oidx.pair(op, pe, exp, bases, offs)
        assert_same(names, got, exp)
        assert bytes(gpe) == bytes(ope)
    # and straight against the reference's own outputs
    tup = gu.results_to_tuples(gidx.entries(), names, got)
    ref = gu.expected(case, tag)
    diff = [(n, tup[n], ref[n][:5]) for n in names if tup[n] != ref[n][:5]]
    assert not diff, "%d reads differ from the reference, first %r" % (len(diff), diff[:5])


def test_index_info_and_getters(golden_dir):
    gidx, oidx = indexes("tiny", golden_dir)
    for f in ("concat_len", "tot_seq_len", "num_entries", "sfx_el_size", "version"):
        assert getattr(gidx.info, f) == getattr(oidx.info, f), f
    assert gidx.info.dataset_name == b"tiny"
    seq = oidx.seq()
    for e, oe in zip(gidx.entries(), oidx.entries()):
        assert (e.entry_id, e.seq_len, e.start_ofs, e.end_ofs, e.name) == (oe.entry_id, oe.seq_len, oe.start_ofs, oe.end_ofs, oe.name)
        got = gidx.get_seq(e.entry_id, 0, e.seq_len)
        assert np.array_equal(got, seq[e.start_ofs:e.end_ofs + 1])
        assert gidx.ident(e.name.decode()) == e.entry_id
    part = gidx.get_seq(1, 19990, 100)
    assert len(part) == 10
    with pytest.raises(bkx.BkxError):
        gidx.entry(99)


def test_align_one_shim(golden_dir):
    gidx, oidx = indexes("tiny", golden_dir)
    p = gidx.default_params(0, max_subs=3)
    names, bases, offs = gu.load_reads("tiny", gu.runs("tiny")["r100_s3"])
    exp, _ = oidx.align(oidx.default_params(0, max_subs=3), bases, offs)
    for i in (0, 1, 5, 17, 59, 107):
        hr, (inst, low, nxt), hit = gidx.align_one(p, bases[offs[i]:offs[i + 1]])
        assert hr == exp[i]["hit_rslt"]
        assert (inst, low, nxt) == (exp[i]["low_hit_instances"], exp[i]["low_mm"], exp[i]["nxt_low_mm"])
        for f in FIELDS:
            assert hit[f] == exp[i][f], (i, f)


def test_edge_inputs(golden_dir):
    gidx, oidx = indexes("tiny", golden_dir)
    p = gidx.default_params(0, max_subs=5)
    op = oidx.default_params(0, max_subs=5)
    seq = oidx.seq()
    reads = []
    ents = oidx.entries()
    # flush against chromosome starts / ends, spanning boundaries, whole short contig, all-N, 1-base, junk codes
    for e in ents:
        reads.append(seq[e.start_ofs:e.start_ofs + 60].copy())
        reads.append(seq[e.end_ofs - 59:e.end_ofs + 1].copy())
        reads.append(seq[e.end_ofs - 30:e.end_ofs + 31].copy())       # crosses the EOS
    reads.append(seq[ents[3].start_ofs:ents[3].end_ofs + 1].copy())   # the entire 300 bp contig
    reads.append(np.full(50, 4, np.uint8))
    reads.append(np.array([2], np.uint8))
    reads.append(np.array([0, 1, 2, 3, 5, 0, 1] * 10, np.uint8))        # code 5 (undefined)
    reads.append(np.zeros(2000, np.uint8))
    nrun = np.nonzero(seq == 4)[0]
    reads.append(seq[nrun[0] - 40:nrun[0] + 60].copy())               # runs into the N block
    reads.append(seq[nrun[0] - 99:nrun[0] + 1].copy())                # exactly one genome N at the end
    for r in reads:
        r[r == 7] = 0
    bases, offs = po.pack_reads(reads)
    got, _ = gidx.align(p, bases, offs)
    exp, _ = oidx.align(op, bases, offs)
    assert_same(None, got, exp)
    # empty batch is a no-op
    out, st = gidx.align(p, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(out) == 0 and st.reads == 0


def test_parameter_errors(golden_dir):
    gidx, _ = indexes("tiny", golden_dir)
    p = gidx.default_params(0)
    p.max_subs = 99
    with pytest.raises(bkx.BkxError):
        gidx.align(p, np.zeros(10, np.uint8), np.array([0, 10], np.uint64))
    with pytest.raises(bkx.BkxError):
        bkx.Index.open("/nonexistent.sfx")


def test_device_resident_variant(golden_dir):
    torch = pytest.importorskip("torch")
    gidx, oidx = indexes("tiny", golden_dir)
    run = gu.runs("tiny")["r150_s3"]
    p, _ = gu.params_from_args(gidx, run["args"])
    names, bases, offs = gu.load_reads("tiny", run)
    exp, _ = oidx.align(gu.params_from_args(oidx, run["args"])[0], bases, offs)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(offs.astype(np.int64)).cuda()
    d_out = torch.zeros(len(exp) * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(C.sizeof(abi.AlignStats), dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    ts.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ts):
        gidx.align_device(p, d_b.data_ptr(), d_o.data_ptr(), len(exp), 150, d_out.data_ptr(), d_st.data_ptr(),
                          ts.cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(abi.RESULT_DTYPE)
    assert_same(names, got, exp)
    assert gidx.last_kernel_ms() > 0
    assert gidx.kernel_launches() > 0


def test_five_byte_suffix_elements_and_wide_prefix_table(golden_dir, tmp_path):
    """5-byte SA elements (what `biokanga index` writes for >= 4e9 symbols) and the u64 prefix table, forced on a
    small genome: results must equal the 4-byte run.  Also round-trips a 5-byte .sfx through the writer/reader."""
    import os
    _, oidx = indexes("tiny", golden_dir)
    seq = np.array(oidx.seq())
    sa4 = np.array(oidx.sa_bytes()).view(np.uint32)
    sa5 = np.zeros((len(sa4), 5), dtype=np.uint8)
    sa5[:, :4] = sa4.view(np.uint8).reshape(-1, 4)
    ents = np.zeros(oidx.info.num_entries, dtype=abi.ENTRY_DTYPE)
    for i, e in enumerate(oidx.entries()):
        ents[i] = (e.entry_id, e.seq_len, e.start_ofs, e.end_ofs, e.name)
    run = gu.runs("tiny")["r100_s5_e2"]
    names, bases, offs = gu.load_reads("tiny", run)
    exp, _ = oidx.align(gu.params_from_args(oidx, run["args"])[0], bases, offs, nthreads=4)
    os.environ["BKX_FORCE_WIDE_PT"] = "1"
    try:
        g5 = bkx.Index.from_host(seq, sa5.reshape(-1), 5, ents, name="tiny5")
        bkx.write_sfx(str(tmp_path / "t5.sfx"), seq, sa5.reshape(-1), 5, ents, name="tiny5")
        g5f = bkx.Index.open(str(tmp_path / "t5.sfx"))
    finally:
        del os.environ["BKX_FORCE_WIDE_PT"]
    assert g5.info.sfx_el_size == 5 and g5f.info.sfx_el_size == 5
    for g in (g5, g5f):
        got, _ = g.align(gu.params_from_args(g, run["args"])[0], bases, offs)
        assert_same(names, got, exp)
    o5 = po.OracleIndex(str(tmp_path / "t5.sfx"))
    got, _ = o5.align(gu.params_from_args(o5, run["args"])[0], bases, offs)
    assert_same(names, got, exp)


def test_index_self_check_rejects_a_damaged_suffix_array(golden_dir):
    """Opening an index ends with a device self-check (every suffix-array element inside the prefix-table bucket of its
    suffix): a suffix array with one 4 kB page zeroed, or belonging to another sequence, is refused -- unless
    BKX_NO_VERIFY asks for it."""
    import os
    oi = po.OracleIndex(gu.sfx_path("repeats", golden_dir))
    seq = np.array(oi.seq())
    sa = np.array(oi.sa_bytes()).view(np.uint32).copy()
    ents = np.zeros(len(oi.entries()), dtype=abi.ENTRY_DTYPE)
    for i, e in enumerate(oi.entries()):
        ents[i] = (e.entry_id, e.seq_len, e.start_ofs, e.end_ofs, e.name if isinstance(e.name, bytes) else e.name.encode())
    good = bkx.Index.from_host(seq, sa, 4, ents, name="ok")
    good.close()
    bad = sa.copy()
    bad[50_000:51_024] = 0
    with pytest.raises(bkx.BkxError, match="self-check"):
        bkx.Index.from_host(seq, bad, 4, ents, name="damaged")
    os.environ["BKX_NO_VERIFY"] = "1"
    try:
        bkx.Index.from_host(seq, bad, 4, ents, name="unchecked").close()
    finally:
        os.environ.pop("BKX_NO_VERIFY", None)
