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


def test_pack_bases2_and_expand_results16_are_host_functions():
    """The two host-side helpers of the compact interface need no device: 2-bit stream + exception list, and the
    16-byte -> 32-byte record expansion."""
    import numpy as np
    rng = np.random.default_rng(3)
    for n in (0, 1, 3, 4, 5, 1003):
        b = rng.integers(0, 4, n).astype(np.uint8)
        if n > 5:
            b[[0, 5, n - 1]] = [4, 6, 4]
        pk, pos, code = bkx.pack_bases2(b | 0x30)   # quality bits in the high nibble are ignored
        dec = np.array([(pk[i // 4] >> (2 * (i % 4))) & 3 for i in range(n)], dtype=np.uint8)
        exp = b.copy()
        exp[b > 3] = 0
        assert np.array_equal(dec, exp)
        assert np.array_equal(pos, np.nonzero(b > 3)[0]) and np.array_equal(code, b[b > 3])
    r16 = np.zeros(3, dtype=abi.RESULT16_DTYPE)
    r16["nar_hr"] = [1 | (1 << 5), 5 | (3 << 5), 3]
    r16["strand_flags"] = [2 | (1 << 2), 3, 0]
    r16["chrom_id"], r16["match_loci"], r16["mismatches"], r16["low_hit_instances"] = [7, 0, 0], [123, 0, 0], [2, 0, 0], [1, 2, 0]
    r = bkx.expand_results16(r16, lens=np.array([100, 150, 75], dtype=np.uint16))
    assert list(r["nar"]) == [1, 5, 3] and list(r["hit_rslt"]) == [1, 3, 0] and list(r["strand"]) == [ord("-"), ord("?"), 0]
    assert list(r["match_len"]) == [100, 150, 0] and list(r["flags"]) == [1, 0, 0] and list(r["mismatches"]) == [2, 0, 0]
    assert abi.RESULT16_DTYPE.itemsize == 16
