"""GPU: the homeolog-rich regime of BASELINE.json configs[3] at a size the oracle handles in seconds -- a 3 x 7
chromosome hexaploid whose B and D sub-genomes are 2-5 % diverged copies of A, suffix array built by the
bounded-memory builder as 5-byte elements and served that way.  Every record against the oracle; multi-map classes must
actually occur (LocateCoreMultiples' MMDelta / hit-instance rules, libbiokanga/SfxArrayV2.cpp:6157-6261)."""
import numpy as np
import pytest

import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx
from biokanga_b200 import workload as wl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def hexa():
    d_seq, ents = wl.make_hexaploid(1_200_000, seed=4, device="cuda", block=20_000)
    n = int(d_seq.numel())
    # 5-byte elements back to back (the .sfx element layout; one fetch per element), read through the two-level prefix
    # table that indexes of >= 2^32 symbols get (forced here)
    d_sa5 = torch.zeros(n * 5 + 16, dtype=torch.uint8, device="cuda")
    bkx.build_suffix_array_packed5(d_seq.data_ptr(), n, d_sa5.data_ptr(), 0, 5_000_000)
    torch.cuda.synchronize()
    import os
    os.environ["BKX_FORCE_WIDE_PT"] = "1"
    try:
        gidx = bkx.Index.from_packed5(d_seq.data_ptr(), n, d_sa5.data_ptr(), ents, name="hexa")
    finally:
        del os.environ["BKX_FORCE_WIDE_PT"]
    sa5 = d_sa5[:n * 5].cpu().numpy()
    seq = d_seq.cpu().numpy()
    oidx = po.OracleIndex(seq=seq, sa=sa5, el_size=5, entries=ents)
    d_bases, d_offs = wl.sim_reads(d_seq, ents, 60_000, 150, seed=9, subs=(0, 1, 2, 3, 4))
    keep = (d_seq, d_sa5)
    return gidx, oidx, d_bases.cpu().numpy(), d_offs.cpu().numpy().astype(np.uint64), keep


@pytest.mark.parametrize("max_subs,mmd", [(3, 1), (3, 2), (6, 1), (10, 1)])
def test_hexaploid_records_match_oracle(hexa, max_subs, mmd):
    gidx, oidx, bases, offs, _ = hexa
    assert gidx.info.sfx_el_size == 5
    got, gst = gidx.align(gidx.default_params(0, max_subs=max_subs, min_edit_dist=mmd), bases, offs)
    exp, ost = oidx.align(oidx.default_params(0, max_subs=max_subs, min_edit_dist=mmd), bases, offs, nthreads=8)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert gst.as_dict() == ost.as_dict()
    hist = np.bincount(got["nar"], minlength=abi.NAR_COUNT)
    # homeologs: a real share of the reads is rejected as multi-mapping or for lack of a Hamming margin
    assert hist[abi.NAR_ACCEPTED] > 0.3 * len(got)
    assert hist[abi.NAR_MULTIALIGN] + hist[abi.NAR_MMDELTA] > 0.01 * len(got)
