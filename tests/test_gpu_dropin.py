"""GPU: mid-size drop-in run -- the unmodified reference binary (oracle/_ref/biokanga, all host threads) and this repo's
`bkx-align` on the SAME files: a .sfx of a 200 Mbp genome (suffix array built on the GPU, written in the reference's
container format) and a FASTA of 2 M x 150 bp reads (1 M pairs in the paired-end case, with orphan recovery).  Every CSV
row (md5 of the sorted files) and the alignment-summary block of the logs must be identical.  The full-size form of the
same check is `python bench.py --dropin` (3.1 Gbp, 20 M reads)."""
import json
import os
import subprocess
import sys

import pytest

import pyoracle as po
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(os.path.dirname(bkx.LIB_PATH), "bkx-align")


@pytest.mark.parametrize("workload", ["se", "pe"])
def test_reference_binary_and_bkx_align_write_identical_files(workload):
    if not os.path.exists(po.REF_BIN):
        pytest.skip("oracle/_ref/biokanga is not built (oracle/build_ref.sh)")
    assert os.path.exists(CLI), "bkx-align is not built"
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--dropin", "--genome-mbp", "200", "--ref-sample", "2000000",
           "--workload", workload]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["csv_rows"][0] == line["csv_rows"][1] and line["csv_rows"][0] > 1_500_000, line
    assert line["csv_identical"], line
    assert line["summary_identical"], line
