"""GPU: the device-resident PE path (align -> pair -> orphan rescue without leaving the GPU) on simulated
pairs from a GPU-built index, every record and counter against the oracle."""
import ctypes as C

import numpy as np
import pytest

import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx
from biokanga_b200 import workload as wl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("mode,dmin,dmax", [(abi.PE_UNIQUE, 100, 1000), (abi.PE_ORPHAN, 200, 1000), (abi.PE_ORPHAN_SE, 200, 1400),
                                            (abi.PE_UNIQUE_SE, 150, 500)])
def test_device_pe_pipeline_matches_oracle(mode, dmin, dmax):
    lens = wl.chrom_layout(8_000_000, n_chrom=4)
    d_seq, ents = wl.make_genome(lens, seed=17, device="cuda", repeat_frac=0.2, repeat_len=(200, 2000))
    n = int(d_seq.numel())
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0)
    torch.cuda.synchronize()
    gidx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="pe")
    n_pairs = 60_000
    d_bases, d_offs = wl.sim_pairs(d_seq, ents, n_pairs, 150, seed=18, subs=(0, 1, 2, 3, 4, 6), junk_frac=0.05)
    p = gidx.default_params(0, max_subs=3)
    pe = abi.PEParams()
    pe.pe_proc, pe.pair_min_len, pe.pair_max_len = mode, dmin, dmax
    d_out = torch.zeros(2 * n_pairs * 32, dtype=torch.uint8, device="cuda")
    d_pst = torch.zeros(C.sizeof(abi.PEStats) // 8, dtype=torch.int64, device="cuda")
    d_ld = torch.zeros(100001, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    ts.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ts):
        gidx.align_device(p, d_bases.data_ptr(), d_offs.data_ptr(), 2 * n_pairs, 150, d_out.data_ptr(), None, ts.cuda_stream)
        gidx.pair_device(p, pe, d_out.data_ptr(), n_pairs, d_bases.data_ptr(), d_offs.data_ptr(), 150, d_pst.data_ptr(),
                         d_ld.data_ptr(), ts.cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(abi.RESULT_DTYPE)
    oidx = po.OracleIndex(seq=d_seq.cpu().numpy(), sa=d_sa.cpu().numpy().view(np.uint32), el_size=4, entries=ents)
    bases = d_bases.cpu().numpy()
    offs = d_offs.cpu().numpy().astype(np.uint64)
    exp, _ = oidx.align(oidx.default_params(0, max_subs=3), bases, offs, nthreads=8)
    ld = np.zeros(100001, dtype=np.uint32)
    ost = oidx.pair(oidx.default_params(0, max_subs=3), pe, exp, bases, offs, len_dist=ld)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert d_pst.cpu().numpy().tolist() == np.frombuffer(bytes(ost), dtype=np.uint64).astype(np.int64).tolist()
    assert np.array_equal(d_ld.cpu().numpy().view(np.uint32), ld)
    if mode in (abi.PE_ORPHAN, abi.PE_ORPHAN_SE):
        assert ost.partner_paired > 0  # recovery really ran
    assert ost.accepted_num_paired > 0.25 * n_pairs
    # the fused host calls (align + pair per pipeline slice, reads cross PCIe once): same records, counters, histogram
    for packed in (False, True):
        ld2 = np.zeros(100001, dtype=np.uint32)
        rd = bkx.pack_bases4(bases) if packed else bases
        fused, fst, fps = gidx.align_pairs(p, pe, rd, offs, packed=packed, len_dist=ld2)
        assert fused.tobytes() == got.tobytes()
        assert bytes(fps) == bytes(ost)
        assert np.array_equal(ld2, ld)
        assert fst.reads == 2 * n_pairs
