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
