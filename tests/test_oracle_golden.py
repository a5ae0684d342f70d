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
