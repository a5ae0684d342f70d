"""GPU: the many-candidates regime of the fast kernel.  A genome carrying a 90-copy diverged repeat family, reads
drawn from the family, high substitution allowance (short cores): every strand of every phase meets more distinct
candidate loci than the shared-memory key list holds, so the lane-private hash sets in HBM and the one-probe
interval of cores no longer than the prefix-table key are what runs.  Every record against the oracle, and the fast
and general kernels against each other (LocateCoreMultiples, libbiokanga/SfxArrayV2.cpp:5693-6262)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pyoracle as po
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx
from biokanga_b200 import workload as wl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def family_world(seed=11, copies=90, elem_len=400, divergence=0.04, n_reads=6000, read_len=100):
    rng = np.random.default_rng(seed)
    lens = [1_500_000, 900_000, 400_000]
    chroms = [rng.integers(0, 4, L).astype(np.uint8) for L in lens]
    elem = rng.integers(0, 4, elem_len).astype(np.uint8)
    sites = []
    for c in range(copies):
        ci = int(rng.integers(0, len(chroms)))
        pos = int(rng.integers(1000, lens[ci] - elem_len - 1000))
        cp = elem.copy()
        mut = rng.random(elem_len) < divergence
        cp[mut] = (cp[mut] + rng.integers(1, 4, int(mut.sum()))) & 3
        if c % 3 == 0:
            cp = (3 - cp[::-1]).astype(np.uint8)
        chroms[ci][pos:pos + elem_len] = cp
        sites.append((ci, pos))
    seq = np.concatenate([np.concatenate([c, np.array([7], np.uint8)]) for c in chroms])
    ents, _ = wl.entries_for(lens)
    reads = []
    for i in range(n_reads):
        ci, pos = sites[int(rng.integers(0, copies))]
        o = int(rng.integers(0, elem_len - read_len))
        rd = chroms[ci][pos + o:pos + o + read_len].copy()
        k = int(rng.integers(0, 9))
        at = rng.choice(read_len, k, replace=False)
        rd[at] = (rd[at] + rng.integers(1, 4, k)) & 3
        if i & 1:
            rd = (3 - rd[::-1]).astype(np.uint8)
        reads.append(rd)
    bases = np.concatenate(reads)
    offs = np.arange(n_reads + 1, dtype=np.uint64) * read_len
    return seq, ents, bases, offs


@pytest.fixture(scope="module")
def fam():
    seq, ents, bases, offs = family_world()
    d_seq = torch.from_numpy(seq).cuda()
    n = len(seq)
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0)
    torch.cuda.synchronize()
    sa = d_sa.cpu().numpy().view(np.uint32)
    gidx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="family")
    oidx = po.OracleIndex(seq=seq, sa=sa, el_size=4, entries=ents)
    return gidx, oidx, bases, offs


@pytest.mark.parametrize("max_subs,mmd", [(5, 1), (8, 1), (10, 1), (10, 2), (15, 1)])
def test_family_reads_match_oracle(fam, max_subs, mmd):
    gidx, oidx, bases, offs = fam
    got, gst = gidx.align(gidx.default_params(0, max_subs=max_subs, min_edit_dist=mmd), bases, offs)
    exp, ost = oidx.align(oidx.default_params(0, max_subs=max_subs, min_edit_dist=mmd), bases, offs, nthreads=8)
    for f in abi.RESULT_DTYPE.names:
        assert np.array_equal(got[f], exp[f]), f
    assert gst.as_dict() == ost.as_dict()
    # the regime this file is about: reads meet far more candidate loci than the 24 in-smem keys
    assert np.percentile(got["cands"], 90) > 100


def test_general_kernel_alone_gives_the_same_records(fam, tmp_path):
    """BKX_NO_FAST is read at launch time, so the comparison runs in a child process."""
    gidx, oidx, bases, offs = fam
    got, _ = gidx.align(gidx.default_params(0, max_subs=10), bases, offs)
    np.save(tmp_path / "fast.npy", got.view(np.uint8))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import torch\n"
        "import test_gpu_repeat_family as t\n"
        "from biokanga_b200 import lib as bkx\n"
        "seq, ents, bases, offs = t.family_world()\n"
        "d_seq = torch.from_numpy(seq).cuda(); n = len(seq)\n"
        "d_sa = torch.empty(n, dtype=torch.int32, device='cuda')\n"
        "bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0); torch.cuda.synchronize()\n"
        "g = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name='family')\n"
        "got, _ = g.align(g.default_params(0, max_subs=10), bases, offs)\n"
        "ref = np.load(%r)\n"
        "assert got.view(np.uint8).tobytes() == ref.tobytes()\n"
        "print('same')\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path / "fast.npy"))
    env = dict(os.environ, BKX_NO_FAST="1", PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "oracle")]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "same" in out.stdout, out.stderr[-2000:]
