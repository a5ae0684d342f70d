"""CPU: the inputs of the randomised GPU check of the pairing kernels' chromosome filter (tests/test_gpu_zpefilter.py) are
drawn here and run through the oracle alone -- they must reach every arm of the filter (pairs dropped, single ends dropped,
SE fallback, orphan arms), or the GPU check would prove little."""
import numpy as np

import goldutil as gu
import pyoracle as po
import test_gpu_zpefilter as z
from biokanga_b200 import abi


def test_drawn_keep_map_cases_reach_every_filter_arm(golden_dir):
    filtered_pairs = fc_single = se_dropped = recovered = 0
    for seed in range(12):
        case = ["tiny", "repeats", "lowcopy"][seed % 3]
        oidx = po.OracleIndex(gu.sfx_path(case, golden_dir))
        seq = np.array(oidx.seq())
        chroms = [(e.name.decode(), seq[e.start_ofs:e.end_ofs + 1]) for e in oidx.entries()]
        bases, offs, pe, keep, kw = z.draw_filtered_pairing_case(seed, chroms)
        p = oidx.default_params(0, **kw)
        rec, _ = oidx.align(p, bases, offs, nthreads=4)
        plain = rec.copy()
        st0 = oidx.pair(p, pe, plain, bases, offs)
        st = oidx.pair(p, pe, rec, bases, offs, keep=keep)
        fc = abi.NAR_CODES.index("FC")
        f, r = rec[0::2], rec[1::2]
        filtered_pairs += st.num_filtered_by_chrom
        fc_single += int(((f["nar"] == fc) != (r["nar"] == fc)).sum())
        se_dropped += st0.accepted_num_se - st.accepted_num_se
        recovered += st.partner_paired
        assert st0.num_filtered_by_chrom == 0
        assert st.accepted_num_paired <= st0.accepted_num_paired
    assert filtered_pairs > 100 and fc_single > 10 and se_dropped > 10 and recovered > 10, (filtered_pairs, fc_single, se_dropped, recovered)
