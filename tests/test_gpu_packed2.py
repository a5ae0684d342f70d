"""GPU: the compact host interface (bkx_align_reads_packed2 / bkx_align_pairs_packed2: 2 bits per base + exception list in,
16-byte records out) gives the records of the one-byte-per-base call -- golden runs with ragged reads, Ns, junk codes,
paired ends with orphan recovery, tiny pipeline slices (so that slices start inside a packed byte) and both length forms."""
import os

import numpy as np
import pytest

import goldutil as gu
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu

CASES = [("tiny", "r100_s3"), ("tiny", "mixed_s3"), ("tiny", "r251_s6"), ("tiny", "r100_s3_n5"), ("repeats", "r100_s3_m3"),
         ("repeats", "r60_s5"), ("tiny", "pe_U1"), ("tiny", "pe_U3"), ("tiny", "pe_U2"), ("tiny", "pe_U1_far")]


def same_records(a, b):
    for f in abi.RESULT_DTYPE.names:
        if f in ("seeds", "cands", "reserved"):
            continue
        bad = np.nonzero(a[f] != b[f])[0]
        assert len(bad) == 0, (f, int(bad[0]), a[bad[0]], b[bad[0]])


@pytest.mark.parametrize("tiny_slices", [False, True])
@pytest.mark.parametrize("case,tag", CASES)
def test_packed2_records_equal_byte_per_base_records(case, tag, tiny_slices, golden_dir):
    run = gu.runs(case)[tag]
    idx = bkx.Index.open(gu.sfx_path(case, golden_dir))
    names, bases, offs = gu.load_reads(case, run)
    p, pe = gu.params_from_args(idx, run["args"])
    if pe is None:
        exp, est = idx.align(p, bases, offs)
    else:
        ld = np.zeros(100001, dtype=np.uint32)
        exp, est, eps = idx.align_pairs(p, pe, bases, offs, len_dist=ld)
    if tiny_slices:
        os.environ["BKX_SLICE_MIN"], os.environ["BKX_SLICE_MAX"] = "1024", "1999"
    try:
        if pe is None:
            got, gst = idx.align_packed2(p, bases, offs)
            forms = [(got, gst)]
            got2, gst2 = idx.align_packed2(p, bases, offs, fixed=False)   # lengths given even when they are all equal
            forms.append((got2, gst2))
            got3, gst3 = idx.align_packed2(p, bases, offs, first_base=1237)   # a shard that starts inside a longer stream
            forms.append((got3, gst3))
        else:
            ld2 = np.zeros(100001, dtype=np.uint32)
            got, gst, gps = idx.align_packed2(p, bases, offs, pe=pe, len_dist=ld2)
            assert bytes(gps) == bytes(eps) and np.array_equal(ld, ld2)
            forms = [(got, gst)]
    finally:
        os.environ.pop("BKX_SLICE_MIN", None)
        os.environ.pop("BKX_SLICE_MAX", None)
    for g, st in forms:
        same_records(g, exp)
        assert st.as_dict() == est.as_dict()
    idx.close()


@pytest.mark.parametrize("tiny_slices", [False, True])
@pytest.mark.parametrize("case,tag", CASES)
def test_wave_path_gives_the_same_records(case, tag, tiny_slices, golden_dir):
    """BKX_WAVE=1: the default search laid out as streaming kernels per phase (bkx_wave.cuh), the lane-per-read kernel
    redoing what that path hands on -- the records and the statistics (with the seed and candidate totals) equal those of
    the default path.  (profiles/ab_kernel.py compares the 32-byte records of both paths byte for byte at full size.)"""
    run = gu.runs(case)[tag]
    idx = bkx.Index.open(gu.sfx_path(case, golden_dir))
    names, bases, offs = gu.load_reads(case, run)
    p, pe = gu.params_from_args(idx, run["args"])
    ld, ld2 = np.zeros(100001, dtype=np.uint32), np.zeros(100001, dtype=np.uint32)
    if pe is None:
        exp, est = idx.align(p, bases, offs)
    else:
        exp, est, eps = idx.align_pairs(p, pe, bases, offs, len_dist=ld)
    os.environ["BKX_WAVE"] = "1"
    if tiny_slices:
        os.environ["BKX_SLICE_MIN"], os.environ["BKX_SLICE_MAX"] = "1024", "1999"
    try:
        if pe is None:
            got, gst = idx.align_packed2(p, bases, offs)
        else:
            got, gst, gps = idx.align_packed2(p, bases, offs, pe=pe, len_dist=ld2)
            assert bytes(gps) == bytes(eps) and np.array_equal(ld, ld2)
    finally:
        for v in ("BKX_WAVE", "BKX_SLICE_MIN", "BKX_SLICE_MAX"):
            os.environ.pop(v, None)
    same_records(got, exp)                  # the 16-byte records carry no per-read seed / candidate counts ...
    assert gst.as_dict() == est.as_dict()   # ... their sums are in the statistics
    idx.close()


def test_packed2_non_acgt_codes_and_odd_offsets(golden_dir):
    """Reads of every length 20..90 back to back (every slice start phase), Ns and InDel / undefined codes sprinkled in."""
    idx = bkx.Index.open(gu.sfx_path("tiny", golden_dir))
    g = idx.get_seq(1, 0, 15000)
    rng = np.random.default_rng(77)
    reads = []
    for i in range(3000):
        L = 20 + (i % 71)
        s = int(rng.integers(0, len(g) - L))
        r = g[s:s + L].copy()
        if i % 5 == 0:
            r[int(rng.integers(0, L))] = 4
        if i % 97 == 0:
            r[int(rng.integers(0, L))] = int(rng.choice([5, 6, 7]))
        reads.append(r)
    bases = np.concatenate(reads)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    p = idx.default_params(0, max_subs=5, max_ns=2)
    p.min_core_len = 8
    exp, est = idx.align(p, bases, offs)
    os.environ["BKX_SLICE_MIN"], os.environ["BKX_SLICE_MAX"] = "1024", "1025"
    try:
        got, gst = idx.align_packed2(p, bases, offs)
    finally:
        os.environ.pop("BKX_SLICE_MIN", None)
        os.environ.pop("BKX_SLICE_MAX", None)
    same_records(got, exp)
    assert gst.as_dict() == est.as_dict()
    assert int((exp["nar"] == abi.NAR_NS).sum()) > 0 and int((exp["nar"] == abi.NAR_ACCEPTED).sum()) > 1000
    idx.close()


def test_device_resident_packed2_equals_device_resident_bytes(golden_dir):
    """bkx_align_reads_device_packed2 (reads resident in both layouts, the fast kernel takes the 2-bit words) against
    bkx_align_reads_device, ragged reads with Ns and junk codes."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    from biokanga_b200 import workload as wl
    case, tag = "tiny", "mixed_s3"
    run = gu.runs(case)[tag]
    idx = bkx.Index.open(gu.sfx_path(case, golden_dir))
    names, bases, offs = gu.load_reads(case, run)
    bases = np.concatenate([bases, bases[::-1].copy()[:0]])
    p, _ = gu.params_from_args(idx, run["args"])
    n = len(offs) - 1
    d_bases = torch.from_numpy(bases).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    d_pk2, d_flags = wl.pack2_device(d_bases, d_offs)
    assert int(d_flags.sum()) > 0
    max_len = int(np.diff(offs).max())
    outs = []
    for packed in (False, True):
        d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
        d_st = torch.zeros(C.sizeof(abi.AlignStats) // 8, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        if packed:
            idx.align_device_packed2(p, d_bases.data_ptr(), d_pk2.data_ptr(), d_flags.data_ptr(), d_offs.data_ptr(), n, max_len,
                                     d_out.data_ptr(), d_st.data_ptr(), None)
        else:
            idx.align_device(p, d_bases.data_ptr(), d_offs.data_ptr(), n, max_len, d_out.data_ptr(), d_st.data_ptr(), None)
        torch.cuda.synchronize()
        outs.append((d_out.cpu().numpy().view(abi.RESULT_DTYPE).copy(), d_st.cpu().numpy().copy()))
    assert outs[0][0].tobytes() == outs[1][0].tobytes()
    assert np.array_equal(outs[0][1], outs[1][1])
    exp, _ = idx.align(p, bases, offs)
    assert exp.tobytes() == outs[1][0].tobytes()
    idx.close()
