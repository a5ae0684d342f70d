import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(__file__)); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import tempfile
import goldutil as gu, pyoracle as po, synth
from biokanga_b200 import abi, lib as bkx
tmp = tempfile.mkdtemp()
sfx = gu.sfx_path("repeats", tmp)
oi = po.OracleIndex(sfx); seq = np.array(oi.seq()); sa = np.array(oi.sa_bytes()).view(np.uint32)
inv = np.empty(len(sa), np.int64); inv[sa] = np.arange(len(sa))
ents = {e.entry_id: e for e in oi.entries()}
chroms = [(e.name.decode() if isinstance(e.name, bytes) else e.name, seq[e.start_ofs:e.end_ofs + 1]) for e in oi.entries()]
gidx = bkx.Index.open(sfx)
print("info", gidx.info.concat_len, gidx.info.prefix_k, gidx.info.device_bytes)
tot_bad = 0
for seed in range(1, 200, 3):
    rng = np.random.default_rng(5000 + seed)
    L = int(rng.choice([36, 50, 75, 100, 125, 150])); ins_lo = int(rng.choice([L, L + 20, 150, 200, 300]))
    usable = [c for c in chroms if len(c[1]) >= 1200]
    n1, r1, n2, r2 = synth.sim_reads(usable, int(rng.integers(300, 1200)), L, seed=int(rng.integers(1, 1 << 30)),
                                     subs=tuple(range(0, int(rng.integers(1, 7)))), junk_frac=0.08, n_frac=0.0, pe=True,
                                     insert=(max(ins_lo, L), max(ins_lo, L) + int(rng.choice([50, 300, 900]))))
    reads = [x for pair in zip(r1, r2) for x in pair]
    bases = np.concatenate(reads); offs = np.arange(len(reads) + 1, dtype=np.uint64) * L
    kw = dict(max_subs=int(rng.choice([2, 3, 5, 8])), min_edit_dist=int(rng.choice([1, 2])))
    exp, _ = oi.align(oi.default_params(0, **kw), bases, offs, nthreads=4)
    got, _ = gidx.align(gidx.default_params(0, **kw), bases, offs)
    bad = np.nonzero((got["nar"] != exp["nar"]) | (got["cands"] != exp["cands"]) | (got["nxt_low_mm"] != exp["nxt_low_mm"]))[0]
    if len(bad):
        tot_bad += len(bad)
        print("seed", seed, "L", L, kw, "bad reads", len(bad), "of", len(reads))
        for i in bad[:6]:
            e, g = exp[i], got[i]
            line = "   read %d exp nar %d cands %d nxt %d | got nar %d cands %d nxt %d" % (i, e["nar"], e["cands"], e["nxt_low_mm"], g["nar"], g["cands"], g["nxt_low_mm"])
            if e["nar"] == 1 and g["nar"] != 1:
                en = ents[int(e["chrom_id"])]; p = en.start_ofs + int(e["match_loci"])
                line += " | lost locus concat %d (mod64 %d) chrom %d loci %d strand %s mm %d; SA ranks of p+0,+12,+24: %s" % (
                    p, p % 64, e["chrom_id"], e["match_loci"], chr(int(e["strand"])), e["mismatches"], [int(inv[p + o]) for o in (0, 12, 24)])
            print(line)
print("TOTAL BAD", tot_bad)
