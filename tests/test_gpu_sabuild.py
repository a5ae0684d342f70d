"""GPU: suffix-array construction and the .sfx writer against what the reference's `biokanga index`
wrote (golden .sfx files, and a fresh run of the reference binary when oracle/_ref is present)."""
import os
import subprocess

import numpy as np
import pytest

import goldutil as gu
import pyoracle as po
import synth
from biokanga_b200 import abi
from biokanga_b200 import lib as bkx

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def gpu_sa(seq):
    d_seq = torch.from_numpy(np.ascontiguousarray(seq)).cuda()
    d_sa = torch.empty(len(seq), dtype=torch.int32, device="cuda")
    bkx.build_suffix_array_device(d_seq.data_ptr(), len(seq), d_sa.data_ptr(), 0)
    torch.cuda.synchronize()
    return d_sa.cpu().numpy().view(np.uint32)


def check_against(seq, ref_sa, sa, tail=64):
    """Same order as the reference for every suffix except the last few: where two suffixes tie up to the
    end of the concatenation the reference's comparator reads past its buffer (no defined order); here
    the shorter suffix sorts first.  So: drop the last `tail` suffix positions from both and compare."""
    n = len(seq)
    assert np.array_equal(np.sort(sa), np.arange(n, dtype=np.uint32))
    a, b = sa[sa < n - tail], ref_sa[ref_sa < n - tail]
    assert np.array_equal(a, b), "suffix order differs at %d places" % int((a != b).sum())
    # and the full array is sorted under symbol order with shorter-first ties (spot check)
    rng = np.random.default_rng(0)
    for i in rng.integers(0, n - 1, 2000):
        x, y = int(sa[i]), int(sa[i + 1])
        m = min(n - x, n - y, 6000)
        d = np.nonzero(seq[x:x + m] != seq[y:y + m])[0]
        if len(d):
            assert seq[x + d[0]] < seq[y + d[0]], (i, x, y)
        else:
            assert x > y or m == 6000, (i, x, y)


@pytest.mark.parametrize("case", ["tiny", "repeats"])
def test_sa_matches_golden_index(case, golden_dir):
    oidx = po.OracleIndex(gu.sfx_path(case, golden_dir))
    seq = np.array(oidx.seq())
    ref_sa = np.array(oidx.sa_bytes()).view(np.uint32)
    sa = gpu_sa(seq)
    check_against(seq, ref_sa, sa)


def entries_of(chroms):
    ents = np.zeros(len(chroms), dtype=abi.ENTRY_DTYPE)
    ofs = 0
    for i, (nm, c) in enumerate(chroms):
        ents[i] = (i + 1, len(c), ofs, ofs + len(c) - 1, nm.encode())
        ofs += len(c) + 1
    return ents


def concat(chroms):
    parts = []
    for _, c in chroms:
        parts += [c, np.array([7], np.uint8)]
    return np.concatenate(parts)


@pytest.mark.skipif(not os.path.exists(po.REF_BIN), reason="reference binary not built")
def test_sa_and_sfx_writer_against_fresh_reference_index(tmp_path):
    g = synth.make_genome([900000, 600000, 400000, 5000, 700], seed=5, repeat_frac=0.08, repeat_len=(200, 3000))
    fa = tmp_path / "g.fa"
    synth.write_fasta(str(fa), g)
    subprocess.run([po.REF_BIN, "index", "-i", "g.fa", "-o", "ref.sfx", "-r", "g", "-F", "i.log", "-T8"], cwd=tmp_path,
                   check=True, stdout=subprocess.DEVNULL)
    oidx = po.OracleIndex(str(tmp_path / "ref.sfx"))
    seq = concat(g)
    assert np.array_equal(np.array(oidx.seq()), seq)
    ref_sa = np.array(oidx.sa_bytes()).view(np.uint32)
    sa = gpu_sa(seq)
    check_against(seq, ref_sa, sa)
    # the writer produces a container the reference itself loads and aligns with, with identical results
    ents = entries_of(g)
    bkx.write_sfx(str(tmp_path / "ours.sfx"), seq, sa, 4, ents, name="g")
    assert os.path.getsize(tmp_path / "ours.sfx") == os.path.getsize(tmp_path / "ref.sfx")
    n, r = synth.sim_reads(g, 3000, 100, seed=6, subs=(0, 1, 2, 3))
    synth.write_reads_fasta(str(tmp_path / "r.fa"), n, r)
    outs = []
    for sfx in ("ref.sfx", "ours.sfx"):
        subprocess.run([po.REF_BIN, "align", "-I", sfx, "-i", "r.fa", "-s3", "-M0", "-o", sfx + ".csv", "-F",
                        sfx + ".log", "-T8"], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL)
        outs.append(sorted(open(tmp_path / (sfx + ".csv")).read().replace('"g"', '"x"').splitlines()))
    assert outs[0] == outs[1] and len(outs[0]) > 2000
    # and our own aligner on our own index agrees with the oracle on the reference's index
    gidx = bkx.Index.from_host(seq, sa, 4, ents, name="g")
    bases, offs = po.pack_reads(r)
    got, _ = gidx.align(gidx.default_params(0, max_subs=3), bases, offs)
    exp, _ = oidx.align(oidx.default_params(0, max_subs=3), bases, offs, nthreads=4)
    for f in ("nar", "strand", "chrom_id", "match_loci", "mismatches", "low_hit_instances"):
        assert np.array_equal(got[f], exp[f]), f


def gpu_sa_planes(seq, max_batch, with_hi):
    d_seq = torch.from_numpy(np.ascontiguousarray(seq)).cuda()
    n = len(seq)
    d_lo = torch.empty(n, dtype=torch.int32, device="cuda")
    d_hi = torch.full((n,), 0x5a, dtype=torch.uint8, device="cuda") if with_hi else None
    bkx.build_suffix_array_planes(d_seq.data_ptr(), n, d_lo.data_ptr(), d_hi.data_ptr() if with_hi else None, 0, max_batch)
    torch.cuda.synchronize()
    if with_hi:
        assert int(d_hi.max()) == 0  # every position of a small genome has zero high bits, and all were written
    return d_lo.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("case", ["tiny", "repeats"])
def test_bounded_memory_builder_matches_golden_index(case, golden_dir):
    """The batch-wise builder (the one for >= 4e9 symbols) on the golden genomes, forced into many small batches:
    same array as `biokanga index` wrote."""
    oi = po.OracleIndex(gu.sfx_path(case, golden_dir))
    seq = np.array(oi.seq())
    ref_sa = np.array(oi.sa_bytes()).view(np.uint32)
    for max_batch, with_hi in ((0, False), (len(seq) // 13, True)):
        check_against(seq, ref_sa, gpu_sa_planes(seq, max_batch, with_hi))


def test_bounded_memory_builder_equals_doubling_builder_and_serves_an_index():
    """30 Mbp genome with exact and diverged repeats: the two builders agree element for element; an index opened
    over the planes (borrowed, 5-byte elements forced) aligns like one opened over the u32 array."""
    from biokanga_b200 import workload as wl
    lens = wl.chrom_layout(30_000_000, n_chrom=6)
    d_seq, ents = wl.make_genome(lens, seed=21, device="cuda")
    n = int(d_seq.numel())
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0)
    d_lo = torch.empty(n, dtype=torch.int32, device="cuda")
    d_hi = torch.empty(n, dtype=torch.uint8, device="cuda")
    bkx.build_suffix_array_planes(d_seq.data_ptr(), n, d_lo.data_ptr(), d_hi.data_ptr(), 0, 4_000_000)
    torch.cuda.synchronize()
    assert torch.equal(d_sa, d_lo) and int(d_hi.max()) == 0
    g4 = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="u32")
    g5 = bkx.Index.from_planes(d_seq.data_ptr(), n, d_lo.data_ptr(), d_hi.data_ptr(), ents, name="planes")
    assert g5.info.sfx_el_size == 5 and g4.info.sfx_el_size == 4
    d_bases, d_offs = wl.sim_reads(d_seq, ents, 50_000, 120, seed=5, subs=(0, 1, 2, 3, 5, 7))
    bases, offs = d_bases.cpu().numpy(), d_offs.cpu().numpy().astype(np.uint64)
    r4, s4 = g4.align(g4.default_params(0, max_subs=5), bases, offs)
    r5, s5 = g5.align(g5.default_params(0, max_subs=5), bases, offs)
    assert r4.tobytes() == r5.tobytes() and s4.as_dict() == s5.as_dict()
    g5.close()
    assert int(d_lo[0]) == int(d_sa[0])  # the borrowed planes survive the index
