"""GPU, two or more devices in ONE process (skipped on a single-GPU box): the multi-GPU product path of the library and
of the front end -- `bkx_clone_index` (peer copies of the index to another GPU) and `bkx-align --gpus N` (one host
thread per GPU, each aligning a contiguous, even-sized read range).  Records of the clone must equal device 0's, and a
2-GPU run of the front end must write the same bytes as the 1-GPU run."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

import goldutil as gu
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu
CLI = os.path.join(os.path.dirname(bkx.LIB_PATH), "bkx-align")


def _n_gpus():
    try:
        return int(bkx.lib().bkx_device_count())
    except Exception:
        return 0


need2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs in one process")


def clone(idx, device):
    h = C.c_void_p()
    bkx.check(bkx.lib().bkx_clone_index(idx._h, device, C.byref(h)))
    return bkx.Index(h)


@need2
@pytest.mark.parametrize("case,tag", [("tiny", "r150_s3"), ("repeats", "r100_s3"), ("tiny", "pe_U1"), ("tiny", "r251_s6")])
def test_clone_on_second_gpu_gives_identical_records(case, tag, golden_dir):
    """Every record and the stats of the clone on device 1 equal device 0's; r150 / r251 need the > 48 KB dynamic
    shared-memory opt-in of the fast kernel, which belongs to the device (not the process)."""
    run = gu.runs(case)[tag]
    idx0 = bkx.Index.open(gu.sfx_path(case, golden_dir), device=0)
    idx1 = clone(idx0, 1)
    assert idx1.info.device == 1 and idx1.info.concat_len == idx0.info.concat_len and idx1.self_check() == 0
    names, bases, offs = gu.load_reads(case, run)
    p, pe = gu.params_from_args(idx0, run["args"])
    if pe is None:
        a, sa = idx0.align(p, bases, offs)
        b, sb = idx1.align(p, bases, offs)
    else:
        a, sa, pa = idx0.align_pairs(p, pe, bases, offs)
        b, sb, pb = idx1.align_pairs(p, pe, bases, offs)
        assert bytes(pa) == bytes(pb)
    assert a.tobytes() == b.tobytes()
    assert sa.as_dict() == sb.as_dict()
    idx1.close()
    idx0.close()


@need2
def test_two_gpus_driven_from_two_host_threads_at_once(golden_dir):
    """What `bkx-align --gpus 2` does: one host thread per GPU, both calling into the library at the same time (first use
    of every kernel on both devices happens concurrently)."""
    case, tag = "tiny", "r150_s3"
    run = gu.runs(case)[tag]
    idx0 = bkx.Index.open(gu.sfx_path(case, golden_dir), device=0)
    idx1 = clone(idx0, 1)
    names, bases, offs = gu.load_reads(case, run)
    p, _ = gu.params_from_args(idx0, run["args"])
    ref, _ = idx0.align(p, bases, offs)
    half = (len(offs) - 1) // 2
    out = np.zeros(len(offs) - 1, dtype=abi.RESULT_DTYPE)
    errs = []

    def work(ix, lo, hi):
        try:
            for _ in range(3):
                got, _ = ix.align(p, bases, offs[lo:hi + 1])
                out[lo:hi] = got
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(idx0, 0, half)), threading.Thread(target=work, args=(idx1, half, len(offs) - 1))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert out.tobytes() == ref.tobytes()
    idx1.close()
    idx0.close()


@need2
@pytest.mark.parametrize("case,tag,fmt", [("tiny", "pe_U1", "-M0"), ("tiny", "pe_U3", "-M6"), ("tiny", "r150_s3", "-M0"),
                                          ("repeats", "r100_s3", "-M6"), ("lowcopy", "r5_R5_s3", "-M0")])
def test_front_end_on_two_gpus_writes_the_same_bytes(case, tag, fmt, golden_dir, tmp_path):
    assert os.path.exists(CLI), "bkx-align is not built"
    run = gu.runs(case)[tag]
    sfx = gu.sfx_path(case, golden_dir)
    files = [os.path.join(gu.GOLD, case, f) for f in run["reads"]]
    base = [CLI, "align", "-I", sfx, "-i", files[0]] + (["-u", files[1]] if len(files) > 1 else []) + run["args"] + [fmt]
    outs = []
    for g in (1, 2):
        o = tmp_path / ("o%d" % g)
        subprocess.run(base + ["--gpus", str(g), "-o", str(o), "-F", str(o) + ".log"], check=True, stdout=subprocess.DEVNULL)
        outs.append(open(o, "rb").read())
    assert outs[0] == outs[1]
    import test_gpu_cli as g
    assert g.summary_block(str(tmp_path / "o1") + ".log") == g.summary_block(str(tmp_path / "o2") + ".log")
