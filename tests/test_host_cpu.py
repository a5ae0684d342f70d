"""CPU: host-side code of the front end that needs no GPU -- the parallel BGZF (BAM) block writer against the sequential
writer it replaces (byte-identical files and virtual addresses over 60 random record streams)."""
import os
import shutil
import subprocess

import pytest

from biokanga_b200 import lib as bkx

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_parallel_bgzf_writer_equals_sequential_writer(tmp_path):
    libdir = os.path.dirname(bkx.LIB_PATH)
    assert os.path.exists(bkx.LIB_PATH), "libbkx.so is not built"
    exe = tmp_path / "bgzf_harness"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(HERE, "bgzf_harness.cpp"), "-L" + libdir, "-lbkx",
                    "-lz", "-lpthread", "-Wl,-rpath," + libdir], check=True, cwd=HERE)
    out = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0 and "all 60 trials identical" in out.stdout, out.stdout[-2000:]


def test_pack_bases4_host_function():
    """bkx_pack_bases4 (pure host code): two bases per byte, even base in the low nibble, odd tail padded with zero."""
    import numpy as np
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 7, 150, 1001):
        b = rng.integers(0, 5, n).astype(np.uint8)
        p = bkx.pack_bases4(b | (rng.integers(0, 16, n).astype(np.uint8) << 4))  # quality nibbles must be dropped
        assert p.size == (n + 1) // 2
        un = np.empty(2 * p.size, np.uint8)
        un[0::2], un[1::2] = p & 15, p >> 4
        assert np.array_equal(un[:n], b)
        if n & 1:
            assert un[n] == 0


def test_clustering_rejects_bad_arguments():
    import numpy as np
    from biokanga_b200 import abi
    res = np.zeros(4, dtype=abi.RESULT_DTYPE)
    multi = np.zeros((4, 5), dtype=abi.MULTI_DTYPE)
    with pytest.raises(bkx.BkxError):
        bkx.assign_multi_matches(res, multi, 1, 100)      # -r1 has no clustering
    cs = bkx.assign_multi_matches(res, multi, 3, 100)      # nothing aligned: nothing to do
    assert (cs.multi_reads, cs.putative, cs.assigned) == (0, 0, 0)
