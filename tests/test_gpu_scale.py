"""GPU: a mid-size end-to-end run on a GPU-built index (30 Mbp genome with injected repeats, 400 k x 150 bp
reads): every record against the oracle, plus size-independent properties that also hold at the bench's full
size -- truth recovery of uniquely placed reads, idempotence, device-resident == host-buffer results, stats =
histogram of the records."""
import numpy as np
import pytest

import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx
from biokanga_b200 import workload as wl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def world():
    lens = wl.chrom_layout(30_000_000, n_chrom=6)
    d_seq, ents = wl.make_genome(lens, seed=7, device="cuda")
    n = int(d_seq.numel())
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0)
    torch.cuda.synchronize()
    gidx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="scale")
    d_bases, d_offs = wl.sim_reads(d_seq, ents, 400_000, 150, seed=8, subs=(0, 1, 2, 3, 4, 6))
    seq = d_seq.cpu().numpy()
    sa = d_sa.cpu().numpy().view(np.uint32)
    oidx = po.OracleIndex(seq=seq, sa=sa, el_size=4, entries=ents)
    return gidx, oidx, d_bases, d_offs, seq


def test_every_record_matches_oracle_and_properties_hold(world):
    gidx, oidx, d_bases, d_offs, seq = world
    bases = d_bases.cpu().numpy()
    offs = d_offs.cpu().numpy().astype(np.uint64)
    n = len(offs) - 1
    p = gidx.default_params(0, max_subs=3)
    got, gst = gidx.align(p, bases, offs)
    exp, ost = oidx.align(oidx.default_params(0, max_subs=3), bases, offs, nthreads=8)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert gst.as_dict() == ost.as_dict()
    # stats are the histogram of the records
    hist = np.bincount(got["nar"], minlength=abi.NAR_COUNT)
    assert [int(x) for x in gst.nar] == [int(x) for x in hist]
    assert gst.reads == n and gst.seeds == int(got["seeds"].astype(np.int64).sum())
    # idempotence and device-resident == host-buffer path
    again, _ = gidx.align(p, bases, offs)
    assert again.tobytes() == got.tobytes()
    d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    ts.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ts):
        gidx.align_device(p, d_bases.data_ptr(), d_offs.data_ptr(), n, 150, d_out.data_ptr(), None, ts.cuda_stream)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().view(abi.RESULT_DTYPE).tobytes() == got.tobytes()
    # every accepted alignment really has the reported number of mismatches at the reported locus
    acc = np.nonzero(got["nar"] == abi.NAR_ACCEPTED)[0][:20000]
    ents = {e.entry_id: e for e in oidx.entries()}
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    for i in acc[::50]:
        r = got[i]
        e = ents[int(r["chrom_id"])]
        g = seq[e.start_ofs + int(r["match_loci"]): e.start_ofs + int(r["match_loci"]) + 150]
        rd = bases[offs[i]:offs[i + 1]]
        if chr(int(r["strand"])) == "-":
            rd = comp[rd[::-1]]
        assert int((rd != g).sum()) == int(r["mismatches"]) <= 5
    # reads with <= 4 substitutions from a repeat-poor genome overwhelmingly come back accepted
    assert hist[abi.NAR_ACCEPTED] > 0.75 * n


def test_long_and_ragged_reads_take_the_general_kernel(world):
    gidx, oidx, d_bases, d_offs, seq = world
    rng = np.random.default_rng(3)
    ents = oidx.entries()
    reads = []
    for L in list(rng.integers(50, 900, 300)) + [1999, 2000, 321, 320, 15]:
        e = ents[int(rng.integers(0, 6))]
        s = int(rng.integers(0, e.seq_len - L))
        r = seq[e.start_ofs + s: e.start_ofs + s + L].copy()
        k = int(rng.integers(0, 1 + L // 40))
        pos = rng.choice(L, k, replace=False)
        r[pos] = (r[pos] + 1) & 3
        if rng.random() < 0.1:
            r[int(rng.integers(0, L))] = 4
        reads.append(r)
    bases, offs = po.pack_reads(reads)
    p = gidx.default_params(0, max_subs=5)
    got, _ = gidx.align(p, bases, offs)
    exp, _ = oidx.align(oidx.default_params(0, max_subs=5), bases, offs, nthreads=8)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f


def test_packed4_host_call_equals_byte_per_base_call(world):
    """bkx_align_reads_packed4: 1.1 M ragged reads of odd and even lengths (so slices of the internal pipeline start on
    both nibble phases) give the records of the one-byte-per-base call, bit for bit."""
    gidx, oidx, d_bases, d_offs, seq = world
    rng = np.random.default_rng(11)
    n = 1_100_000
    lens = rng.choice([51, 64, 75, 100, 33], n).astype(np.uint64)
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    ents = oidx.entries()
    e = ents[0]
    starts = rng.integers(e.start_ofs, e.start_ofs + e.seq_len - 200, n)
    total = int(offs[-1])
    idx = np.repeat(starts - offs[:-1].astype(np.int64), lens.astype(np.int64)) + np.arange(total)
    bases = seq[idx].copy()
    flip = rng.random(total) < 0.01
    bases[flip] = (bases[flip] + 1) & 3
    p = gidx.default_params(0, max_subs=3)
    a, sa_ = gidx.align(p, bases, offs)
    packed = bkx.pack_bases4(bases)
    assert packed.size == (total + 1) // 2
    b, sb_ = gidx.align_packed4(p, packed, offs)
    assert a.tobytes() == b.tobytes() and sa_.as_dict() == sb_.as_dict()
    assert (a["nar"] == abi.NAR_ACCEPTED).mean() > 0.5
