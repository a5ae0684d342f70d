"""Synthetic workloads generated on the GPU with torch (plumbing for bench.py and the scale tests):
human-sized random genomes with injected diverged repeats and reads
drawn from them with a controlled number of substitutions.  Deterministic for a given seed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import abi


def chrom_layout(total_bp, n_chrom=24, short_contigs=(50000, 20000, 5000, 2000, 700)):
    """Human-like: n_chrom large chromosomes of decreasing size plus a few short contigs."""
    w = np.linspace(2.0, 0.6, n_chrom)
    big = total_bp - sum(short_contigs)
    lens = [int(big * x / w.sum()) for x in w]
    lens[0] += big - sum(lens)
    lens += [int(x) for x in short_contigs if x < total_bp // 50]
    return lens


def entries_for(lens):
    ents = np.zeros(len(lens), dtype=abi.ENTRY_DTYPE)
    ofs = 0
    for i, ln in enumerate(lens):
        ents[i] = (i + 1, ln, ofs, ofs + ln - 1, ("chr%d" % (i + 1)).encode())
        ofs += ln + 1
    return ents, ofs


def make_genome(lens, seed=1, device="cuda", repeat_frac=0.05, repeat_len=(300, 5000),
                divergences=(0.0, 0.005, 0.01, 0.03), n_runs=20, n_run_len=60):
    """Returns (d_seq uint8[concat_len] with one EOS(7) after every chromosome, entries)."""
    ents, n = entries_for(lens)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    seq = torch.randint(0, 4, (n,), dtype=torch.uint8, device=device, generator=g)
    if repeat_frac > 0 and n > 40 * repeat_len[1]:
        avg = (repeat_len[0] + repeat_len[1]) // 2
        S = max(1, int(n * repeat_frac / avg))
        stride = n // S
        lens_t = torch.randint(repeat_len[0], repeat_len[1] + 1, (S,), device=device, generator=g)
        jitter = (torch.rand(S, device=device, generator=g) * (stride - repeat_len[1] - 1)).long()
        dst0 = torch.arange(S, device=device) * stride + jitter
        src0 = (torch.rand(S, device=device, generator=g) * (n - repeat_len[1] - 1)).long()
        rc = torch.rand(S, device=device, generator=g) < 0.5
        div = torch.tensor(divergences, device=device)[torch.randint(0, len(divergences), (S,), device=device, generator=g)]
        orig = seq.clone()  # sources are read from the pristine random sequence (deterministic, EOS-free)
        CH = 4096  # segments per chunk keeps the flat index tensors small
        for s0 in range(0, S, CH):
            sl = slice(s0, min(S, s0 + CH))
            ln = lens_t[sl]
            seg = torch.repeat_interleave(torch.arange(ln.numel(), device=device), ln)
            start = torch.cumsum(ln, 0) - ln
            within = torch.arange(int(ln.sum()), device=device) - start[seg]
            r = rc[sl][seg]
            sidx = torch.where(r, src0[sl][seg] + ln[seg] - 1 - within, src0[sl][seg] + within)
            v = orig[sidx]
            v = torch.where(r, 3 - v, v)
            mut = torch.rand(v.numel(), device=device, generator=g) < div[sl][seg]
            add = torch.randint(1, 4, (v.numel(),), dtype=torch.uint8, device=device, generator=g)
            v = torch.where(mut, (v + add) & 3, v)
            seq[dst0[sl][seg] + within] = v
        del orig
    for i in range(n_runs):
        p = int(torch.randint(0, n - n_run_len, (1,), generator=g, device=device).item())
        seq[p:p + n_run_len] = 4
    eos = torch.from_numpy((ents["end_ofs"] + 1).astype(np.int64)).to(device)
    seq[eos] = 7
    return seq, ents


def make_hexaploid(chrom_len, n_chrom=7, seed=1, device="cuda", divergence=((0.02, 0.035), (0.035, 0.05)),
                   block=50_000, shuffle_frac=0.05, n_runs=12, n_run_len=60):
    """Wheat-like hexaploid: sub-genome A is random; B and D are copies of A with
    2-5 % substitutions (one rate per chromosome, drawn from `divergence`) and local shuffles (neighbouring blocks
    swapped).  Entries are ordered 1A..7A, 1B..7B, 1D..7D.  Returns (d_seq, entries) like make_genome."""
    lens = [int(chrom_len)] * (3 * n_chrom)
    ents, n = entries_for(lens)
    for i in range(3 * n_chrom):
        ents[i]["name"] = ("chr%d%s" % (i % n_chrom + 1, "ABD"[i // n_chrom])).encode()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    seq = torch.empty(n, dtype=torch.uint8, device=device)
    CH = 1 << 26
    for c in range(n_chrom):
        a0 = int(ents[c]["start_ofs"])
        for s0 in range(0, chrom_len, CH):
            m = min(CH, chrom_len - s0)
            seq[a0 + s0:a0 + s0 + m] = torch.randint(0, 4, (m,), dtype=torch.uint8, device=device, generator=g)
        for sub in (1, 2):
            d0 = int(ents[sub * n_chrom + c]["start_ofs"])
            lo, hi = divergence[sub - 1]
            rate = lo + (hi - lo) * float(torch.rand(1, device=device, generator=g).item())
            for s0 in range(0, chrom_len, CH):
                m = min(CH, chrom_len - s0)
                v = seq[a0 + s0:a0 + s0 + m]
                mut = torch.rand(m, device=device, generator=g) < rate
                add = torch.randint(1, 4, (m,), dtype=torch.uint8, device=device, generator=g)
                seq[d0 + s0:d0 + s0 + m] = torch.where(mut, (v + add) & 3, v)
            nblk = chrom_len // block
            if nblk > 2 and shuffle_frac > 0:
                k = max(1, int(nblk * shuffle_frac))
                pick = torch.randperm((nblk - 1) // 2, device=device, generator=g)[:k] * 2  # disjoint neighbour pairs
                for b in pick.tolist():
                    x = d0 + b * block
                    t = seq[x:x + block].clone()
                    seq[x:x + block] = seq[x + block:x + 2 * block]
                    seq[x + block:x + 2 * block] = t
    for i in range(n_runs):
        p = int(torch.randint(0, n - n_run_len, (1,), generator=g, device=device).item())
        seq[p:p + n_run_len] = 4
    eos = torch.from_numpy((ents["end_ofs"] + 1).astype(np.int64)).to(device)
    seq[eos] = 7
    return seq, ents


def sim_reads(d_seq, ents, n_reads, length, seed=2, subs=(0, 1, 2, 3, 4), device="cuda", chunk=1 << 21):
    """Returns (d_bases uint8[n_reads*length], d_offsets int64[n_reads+1]).  Each read is a genome
    substring (either strand) carrying k substitutions, k drawn uniformly from `subs`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.from_numpy(ents["seq_len"].astype(np.int64)).to(device)
    starts = torch.from_numpy(ents["start_ofs"].astype(np.int64)).to(device)
    ok = lens >= length
    w = torch.where(ok, lens.double(), torch.zeros_like(lens, dtype=torch.double))
    out = torch.empty(n_reads * length, dtype=torch.uint8, device=device)
    subs_t = torch.tensor(subs, device=device)
    kmax = int(max(subs))
    ar = torch.arange(length, device=device)
    for s0 in range(0, n_reads, chunk):
        m = min(chunk, n_reads - s0)
        c = torch.multinomial(w, m, replacement=True, generator=g)
        p = starts[c] + (torch.rand(m, device=device, generator=g, dtype=torch.double) * (lens[c] - length + 1).double()).long()
        b = d_seq[p[:, None] + ar[None, :]]
        if kmax > 0:
            k = subs_t[torch.randint(0, len(subs), (m,), device=device, generator=g)]
            pos = torch.rand(m, length, device=device, generator=g).topk(kmax, dim=1).indices
            add = torch.randint(1, 4, (m, kmax), dtype=torch.uint8, device=device, generator=g)
            cur = torch.gather(b, 1, pos)
            new = torch.where((torch.arange(kmax, device=device)[None, :] < k[:, None]) & (cur < 4), (cur + add) & 3, cur)
            b = b.scatter(1, pos, new)
        rcm = torch.rand(m, device=device, generator=g) < 0.5
        br = torch.flip(b, dims=[1])
        br = torch.where(br < 4, 3 - br, br)
        b = torch.where(rcm[:, None], br, b)
        out[s0 * length:(s0 + m) * length] = b.reshape(-1)
    offs = torch.arange(n_reads + 1, device=device, dtype=torch.int64) * length
    return out, offs


def sim_pairs(d_seq, ents, n_pairs, length, seed=3, subs=(0, 1, 2, 3, 4), insert=(300, 600), junk_frac=0.02,
              device="cuda", chunk=1 << 20):
    """Paired-end reads: returns (d_bases uint8[2*n_pairs*length], d_offsets int64[2*n_pairs+1]) with PE1 / PE2 of a pair
    adjacent.  PE1 is the first `length` bases of a fragment of length U[insert] (either strand), PE2 the reverse
    complement of its last `length` bases; junk_frac of the mates are replaced by random sequence (orphans)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.from_numpy(ents["seq_len"].astype(np.int64)).to(device)
    starts = torch.from_numpy(ents["start_ofs"].astype(np.int64)).to(device)
    w = torch.where(lens >= insert[1], lens.double(), torch.zeros_like(lens, dtype=torch.double))
    out = torch.empty(2 * n_pairs * length, dtype=torch.uint8, device=device).view(n_pairs, 2, length)
    subs_t = torch.tensor(subs, device=device)
    kmax = int(max(subs))
    ar = torch.arange(length, device=device)

    def mutate(b, m):
        if kmax == 0:
            return b
        k = subs_t[torch.randint(0, len(subs), (m,), device=device, generator=g)]
        pos = torch.rand(m, length, device=device, generator=g).topk(kmax, dim=1).indices
        add = torch.randint(1, 4, (m, kmax), dtype=torch.uint8, device=device, generator=g)
        cur = torch.gather(b, 1, pos)
        new = torch.where((torch.arange(kmax, device=device)[None, :] < k[:, None]) & (cur < 4), (cur + add) & 3, cur)
        return b.scatter(1, pos, new)

    def revcomp(b):
        br = torch.flip(b, dims=[1])
        return torch.where(br < 4, 3 - br, br)

    for s0 in range(0, n_pairs, chunk):
        m = min(chunk, n_pairs - s0)
        c = torch.multinomial(w, m, replacement=True, generator=g)
        fl = torch.randint(insert[0], insert[1] + 1, (m,), device=device, generator=g)
        p = starts[c] + (torch.rand(m, device=device, generator=g, dtype=torch.double) * (lens[c] - fl + 1).double()).long()
        left = d_seq[p[:, None] + ar[None, :]]                      # first `length` bases of the fragment (+ strand)
        right = d_seq[(p + fl - length)[:, None] + ar[None, :]]     # last `length` bases
        minus = torch.rand(m, device=device, generator=g) < 0.5     # fragment taken from the - strand
        pe1 = torch.where(minus[:, None], revcomp(right), left)
        pe2 = torch.where(minus[:, None], left, revcomp(right))
        pe1, pe2 = mutate(pe1, m), mutate(pe2, m)
        junk = torch.rand(m, device=device, generator=g) < junk_frac
        rnd = torch.randint(0, 4, (m, length), dtype=torch.uint8, device=device, generator=g)
        pe2 = torch.where(junk[:, None], rnd, pe2)
        out[s0:s0 + m, 0] = pe1
        out[s0:s0 + m, 1] = pe2
    offs = torch.arange(2 * n_pairs + 1, device=device, dtype=torch.int64) * length
    return out.view(-1), offs


def algorithmic_bytes(results, concat_len, el_size, read_len):
    """W(read) = seeds*S*(E+8) + cands*(E+ceil(L/4)) + ceil(L/4) + 32, summed."""
    S = int(np.ceil(np.log2(concat_len)))
    q = (read_len + 3) // 4
    seeds = int(results["seeds"].astype(np.int64).sum())
    cands = int(results["cands"].astype(np.int64).sum())
    return seeds * S * (el_size + 8) + cands * (el_size + q) + len(results) * (q + 32)


def pack2_device(d_bases, d_offs, chunk_reads=1 << 21):
    """The 2-bit copy of device-resident reads that bkx_align_reads_device_packed2 takes beside the one-byte-per-base layout:
    (uint8 tensor of the packed stream with 16 bytes of slack -- base i at bits [2(i%4), +2) of byte i/4, non-ACGT bases as
    0 --, uint8 per-read flags: the read holds a non-ACGT base).  Works through the reads a chunk at a time: beside a
    144 GB index there is no room for whole-stream temporaries."""
    import torch
    dev = d_bases.device
    nb = int(d_bases.numel())
    n_reads = int(d_offs.numel()) - 1
    out = torch.zeros((((nb + 3) // 4 + 16 + 7) // 8) * 8, dtype=torch.uint8, device=dev)
    flags = torch.zeros(n_reads, dtype=torch.uint8, device=dev)
    offs = d_offs.to(torch.int64)
    r0 = 0
    while r0 < n_reads:
        r1 = min(n_reads, r0 + chunk_reads)
        b0, b1 = int(offs[r0]), int(offs[r1])
        a0 = b0 & ~3                          # start the chunk on a packed byte; the bases before b0 are packed again
        codes = d_bases[a0:b1] & 7
        exc = codes > 3
        c2 = torch.where(exc, torch.zeros_like(codes), codes)
        pad = (-int(c2.numel())) % 4
        if pad:                               # only the last chunk can end inside a byte (the next one starts at b1 & ~3)
            c2 = torch.cat([c2, torch.zeros(pad, dtype=c2.dtype, device=dev)])
        q = c2.view(-1, 4)
        pk = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).to(torch.uint8)
        if r1 < n_reads and (b1 & 3):         # the byte shared with the next chunk is written by that chunk, in full
            pk = pk[:-1]
        out[a0 // 4:a0 // 4 + pk.numel()] = pk
        csum = torch.cumsum(exc.to(torch.int32), 0, dtype=torch.int32)
        csum = torch.cat([torch.zeros(1, dtype=torch.int32, device=dev), csum])
        o = offs[r0:r1 + 1] - a0
        flags[r0:r1] = ((csum[o[1:]] - csum[o[:-1]]) > 0).to(torch.uint8)
        del codes, exc, c2, q, pk, csum, o
        r0 = r1
    return out, flags
