"""CPU: the C-ABI library loads and exports every symbol include/bkx.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bkx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bkx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(bkx.LIB_PATH):
        bkx.build()
    L = C.CDLL(bkx.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libbkx.so does not export %s" % n
    assert sorted(bkx.EXPORTS) == names
    assert bkx.lib().bkx_abi_version() == 1


def test_pod_sizes_match_header():
    assert abi.RESULT_DTYPE.itemsize == 32
    assert C.sizeof(abi.AlignParams) == 64 and C.sizeof(abi.PEParams) == 32
    assert C.sizeof(abi.Entry) == 112 and C.sizeof(abi.IndexInfo) == 136
    assert C.sizeof(abi.AlignStats) == 256 and C.sizeof(abi.PEStats) == 64


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly (never route to the oracle)."""
    if bkx.lib().bkx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bkx.BkxError) as ei:
        bkx.Index.open(os.path.join(ROOT, "tests", "golden", "tiny", "tiny.sfx.gz"))
    assert ei.value.code < 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "biokanga_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in txt and "bk_oracle" not in txt and "libbkoracle" not in txt, f
