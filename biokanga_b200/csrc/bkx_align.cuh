// The read-alignment kernel: one warp per read, persistent grid.
//
// Reference loops replaced (file:line under /root/reference):
//   CAligner::ProcCoredApprox        biokanga/Aligner.cpp:9027-9504   per-read driver + classification
//   CSfxArrayV3::AlignReads          libbiokanga/SfxArrayV2.cpp:7666-7760  staged phases
//   CSfxArrayV3::LocateCoreMultiples libbiokanga/SfxArrayV2.cpp:5693-6262  one phase (cores, walk, Hamming)
//   LocateFirstExact/LocateLastExact libbiokanga/SfxArrayV2.cpp:7765-8027  SA interval of a core
//
// Mapping to the warp:
//   * the read is 2-bit packed (both strands) into the warp's shared-memory slot once;
//   * SEEDS: lane c owns core c of the current strand/phase: k-mer prefix-table bucket, then a
//     lower/upper-bound refinement over the reference's suffix array -> interval [first, first+cnt);
//   * WALK: the SA entries of all cores are flattened in the reference's processing order
//     (core, then SA index) and handed out 32 per step, lane = one candidate locus: entry check,
//     "already seen" test, XOR+popc Hamming over packed 64-bit words, then warp reductions keep
//     (LowMMCnt, NxtLowMMCnt, LowHitInstances, first hit) exactly as the sequential loop would,
//     including the ordered early exit, the 100th-candidate copy probe and the MaxIter cap.
#pragma once
#include "../../include/bkx.h"
#include "bkx_index.cuh"

namespace bkx {

constexpr int kSeenCap = 128;        // "already processed" keys kept in shared memory per warp
constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;

struct KParams {
  int max_subs, mmd, max_ns, strand_mode, max_hits, min_core_len, slides_per100, max_iter, max_nodes;
};

struct WarpCtx {
  // read, both strands, in shared memory
  uint64_t* s2[2];
  uint32_t* sx[2];
  uint32_t* seen;
  int* pre;
  int L;
  bool hasN;
  int lane;
  // "already processed" set of the current strand
  int seen_n;
  bool overflow;
  uint64_t* hash;
  uint32_t hmask;
  uint32_t epoch;
  int nodes;
  // phase state (LowHitInstances, LowMMCnt, NxtLowMMCnt) and first hit
  int inst, low, nxt;
  int hit_strand, hit_ent, hit_mm;
  uint64_t hit_p;
  uint32_t seeds, cands;
};

__device__ __forceinline__ uint64_t read_word(const WarpCtx& c, int s, int pos) {
  int w = pos >> 5;
  unsigned sh = (unsigned)(pos & 31) * 2;
  uint64_t a = c.s2[s][w];
  if (sh == 0) return a;
  return (a >> sh) | (c.s2[s][w + 1] << (64 - sh));
}

__device__ __forceinline__ int rsym(const WarpCtx& c, int s, int i) {
  if ((c.sx[s][i >> 5] >> (i & 31)) & 1) return 4;
  return (int)((c.s2[s][i >> 5] >> ((i & 31) * 2)) & 3);
}

// probe-vs-suffix comparison of SfxArrayV2.cpp:7792-7811: -1 / 0 / +1, target EOS => -1.
static __device__ __noinline__ int cmp_core_slow(const DevIndex& I, const WarpCtx& c, int s, int ofs, int len, uint64_t g) {
  for (int i = 0; i < len; ++i) {
    int gs = gsym(I, g + i);
    if (gs == 7) return -1;
    int rs = rsym(c, s, ofs + i);
    if (rs > gs) return 1;
    if (rs < gs) return -1;
  }
  return 0;
}

__device__ __forceinline__ int cmp_core(const DevIndex& I, const WarpCtx& c, int s, int ofs, int len, uint64_t g) {
  if (c.hasN || span_has_exc(I, g, (uint32_t)len)) return cmp_core_slow(I, c, s, ofs, len, g);
  uint64_t w = g >> 5;
  unsigned sh = (unsigned)(g & 31) * 2;
  uint64_t prev = __ldg(I.g2 + w);
  for (int b = 0; b < len; b += 32) {
    uint64_t next = __ldg(I.g2 + (++w));
    uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    uint64_t rw = read_word(c, s, ofs + b);
    uint64_t x = rw ^ gw;
    int rem = len - b;
    if (rem < 32) x &= (1ull << (2 * rem)) - 1;
    if (x) {
      int pos = (__ffsll((long long)x) - 1) >> 1;
      int rc = (int)((rw >> (2 * pos)) & 3), gc = (int)((gw >> (2 * pos)) & 3);
      return rc > gc ? 1 : -1;
    }
  }
  return 0;
}

// SA interval [first, first+cnt) of suffixes starting with core (strand s, offset ofs, length len);
// cnt == 0 if absent.  Equivalent to LocateFirstExact + LocateLastExact over the whole array.
__device__ __forceinline__ void locate_core(const DevIndex& I, const WarpCtx& c, int s, int ofs, int len,
                                            uint64_t& first, uint64_t& cnt) {
  const int k = I.k;
  int eff = len < k ? len : k;  // leading symbols usable as table key
  bool n_at_eff = false;
  if (c.hasN) {
    for (int j = 0; j < eff; ++j)
      if (rsym(c, s, ofs + j) == 4) { eff = j; n_at_eff = true; break; }
  }
  uint64_t key = rev2(read_word(c, s, ofs)) >> (64 - 2 * k);
  uint64_t blo, bhi;
  if (eff == k) {
    blo = pt_get(I, key);
    bhi = pt_get(I, key + 1);
  } else {
    int sh = 2 * (k - eff);
    uint64_t p = key >> sh;
    if (n_at_eff) {
      uint64_t x = (p << sh) | ((1ull << sh) - 1);
      blo = pt_get(I, x);
      bhi = pt_get(I, x + 1);
    } else {
      blo = pt_get(I, p << sh);
      bhi = pt_get(I, (p + 1) << sh);
    }
  }
  cnt = 0;
  first = 0;
  if (blo >= bhi) return;
  // lower bound; remember whether the element finally at l compared equal
  uint64_t l = blo, h = bhi;
  bool h_equal = false;
  while (l < h) {
    uint64_t m = l + ((h - l) >> 1);
    int r = cmp_core(I, c, s, ofs, len, sa_get(I, m));
    if (r > 0) l = m + 1; else { h = m; h_equal = (r == 0); }
  }
  if (l >= bhi || !h_equal) return;
  first = l;
  // upper bound, probing first+1 first (most intervals hold a single suffix)
  uint64_t ul = l + 1, uh = bhi;
  if (ul < uh) {
    if (cmp_core(I, c, s, ofs, len, sa_get(I, ul)) < 0) uh = ul;
    else ul = ul + 1;
  }
  while (ul < uh) {
    uint64_t m = ul + ((uh - ul) >> 1);
    if (cmp_core(I, c, s, ofs, len, sa_get(I, m)) >= 0) ul = m + 1; else uh = m;
  }
  cnt = ul - first;
}

// Mismatch count of the whole read (strand s) against the concatenation at p; the window is known
// to lie inside one chromosome.  Returns 255 once the count exceeds max_mm (SfxArrayV2.cpp:6093-6152).
__device__ __forceinline__ int hamming(const DevIndex& I, const WarpCtx& c, int s, uint64_t p, int max_mm) {
  const int L = c.L;
  if (span_has_exc(I, p, (uint32_t)L)) {  // genome N inside the window: symbol-wise (N matches N)
    int mm = 0;
    for (int i = 0; i < L; ++i) {
      int gs = gsym(I, p + i);
      if (gs == 7) return 255;
      if (gs != rsym(c, s, i) && ++mm > max_mm) return 255;
    }
    return mm;
  }
  uint64_t w = p >> 5;
  unsigned sh = (unsigned)(p & 31) * 2;
  uint64_t prev = __ldg(I.g2 + w);
  int mm = 0;
  for (int b = 0, wi = 0; b < L; b += 32, ++wi) {
    uint64_t next = __ldg(I.g2 + (++w));
    uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    uint64_t x = c.s2[s][wi] ^ gw;
    uint64_t m = ((x | (x >> 1)) & 0x5555555555555555ull) | spread32(c.sx[s][wi]);
    int rem = L - b;
    if (rem < 32) m &= (1ull << (2 * rem)) - 1;
    mm += __popcll(m);
    if (mm > max_mm) return 255;
  }
  return mm;
}

// ---- "already processed" set (SfxArrayV2.cpp:5931-5950): 32-bit keys, reset per strand ----------
__device__ __forceinline__ bool seen_contains(const WarpCtx& c, uint32_t key) {
  if (!c.overflow) {
    bool f = false;
    for (int i = 0; i < c.seen_n; ++i) f |= (c.seen[i] == key);
    return f;
  }
  uint64_t want = ((uint64_t)c.epoch << 32) | key;
  uint32_t h = (key * 2654435761u) & c.hmask;
  for (;;) {
    uint64_t v = __ldcg(c.hash + h);  // L2 view: other lanes of this warp insert with atomics
    if (v == want) return true;
    if ((uint32_t)(v >> 32) != c.epoch) return false;
    h = (h + 1) & c.hmask;
  }
}

__device__ __forceinline__ void hash_insert(WarpCtx& c, uint32_t key) {
  uint64_t want = ((uint64_t)c.epoch << 32) | key;
  uint32_t h = (key * 2654435761u) & c.hmask;
  for (;;) {
    uint64_t v = __ldcg(c.hash + h);
    if ((uint32_t)(v >> 32) != c.epoch) {
      uint64_t old = atomicCAS((unsigned long long*)&c.hash[h], (unsigned long long)v, (unsigned long long)want);
      if (old == v) return;
      continue;  // another lane took the slot: re-read it
    }
    if (v == want) return;
    h = (h + 1) & c.hmask;
  }
}

// all lanes call; lanes with ins==true add their (distinct) key
__device__ __forceinline__ void seen_insert(WarpCtx& c, bool ins, uint32_t key) {
  unsigned m = __ballot_sync(kFull, ins);
  int cnt = __popc(m);
  if (cnt == 0) return;
  if (!c.overflow && c.seen_n + cnt <= kSeenCap) {
    if (ins) c.seen[c.seen_n + __popc(m & ((1u << c.lane) - 1))] = key;
    c.seen_n += cnt;
    __syncwarp();
    return;
  }
  if (!c.overflow) {  // migrate the shared-memory list into this warp's global hash set
    c.overflow = true;
    c.epoch += 1;
    __syncwarp();
    for (int i = c.lane; i < c.seen_n; i += 32) hash_insert(c, c.seen[i]);
    __syncwarp();
  }
  if (ins) hash_insert(c, key);
  c.seen_n += cnt;
  __threadfence_block();
  __syncwarp();
}

// ---- one step of the interval walk: up to 32 SA entries, one per lane, in processing order ------
// valid/cofs/sidx are lane-private.  capped => all lanes belong to one core whose interval ends at
// hi_idx and iter_cnt counts its new candidates so far (the 100th-candidate probe and MaxIter cap
// of SfxArrayV2.cpp:5857-5875 apply).  Returns the lane at which processing stopped (or -1).
__device__ __forceinline__ int walk_step(const DevIndex& I, const KParams& P, WarpCtx& c, int s, int max_mm,
                                         bool valid, int cofs, uint64_t sidx, bool capped, int& iter_cnt,
                                         uint64_t hi_idx, bool& stop_core, bool& stop_strand, bool& stop_all) {
  const unsigned lt = (1u << c.lane) - 1;
  uint64_t loci = valid ? sa_get(I, sidx) : 0;
  bool ok = valid && loci >= (uint64_t)cofs;
  uint64_t p = loci - (uint64_t)cofs;
  int ent = -1;
  if (ok) {
    ent = find_entry(I, p);
    ok = ent >= 0 && (p + (uint64_t)c.L - 1) <= __ldg(I.ent_end + ent);
  }
  uint32_t key = (uint32_t)(1u + (uint32_t)loci - (uint32_t)cofs);
  bool dup = ok && seen_contains(c, key);
  unsigned okm = __ballot_sync(kFull, ok);
  unsigned same = __match_any_sync(kFull, key);
  if (ok && (same & okm & lt)) dup = true;  // an earlier lane of this step already claims the key
  bool isnew = ok && !dup;
  unsigned newm = __ballot_sync(kFull, isnew);
  int stop_lane = -1;
  // caps, in processing order
  if (newm) {
    int rank = iter_cnt + __popc(newm & (lt | (1u << c.lane)));  // 1-based index among new candidates
    int cut = 32;
    if (capped) {
      bool at100 = isnew && rank == 100 && sidx < hi_idx && (hi_idx - sidx + 2) > (uint64_t)P.max_iter;
      bool atmax = isnew && rank == P.max_iter;
      unsigned cm = __ballot_sync(kFull, at100 || atmax);
      if (cm) { cut = __ffs(cm) - 1; stop_core = true; }
    }
    // identifier-node budget (cMaxNumIdentNodes): the candidate that fills it is still processed
    int allowed = P.max_nodes - c.nodes;
    if (__popc(newm & (cut >= 31 ? kFull : ((2u << cut) - 1))) >= allowed) {
      int nl = (int)__fns(newm, 0, allowed);  // lane of the allowed-th new candidate
      if (nl < cut || cut == 32) { cut = nl; }
      stop_core = true;
      stop_strand = true;
    }
    if (cut < 32) {
      stop_lane = cut;
      if (c.lane > cut) isnew = false;
      newm = __ballot_sync(kFull, isnew);
    }
  }
  int mm = 255;
  if (isnew) mm = hamming(I, c, s, p, max_mm);
  bool acc = isnew && mm <= max_mm;
  // ordered early exit: the (MaxHits+1)-th exact match ends the whole search (SfxArrayV2.cpp:6206-6214)
  unsigned zm = __ballot_sync(kFull, acc && mm == 0);
  if (zm) {
    int before = (c.low == 0) ? c.inst : 0;
    if (before + __popc(zm) > P.max_hits) {
      int need = P.max_hits + 1 - before;
      int el = (int)__fns(zm, 0, need);
      if (c.lane > el) { isnew = false; acc = false; }
      newm = __ballot_sync(kFull, isnew);
      stop_all = true;
      stop_core = true;
      stop_lane = el;
    }
  }
  int nnew = __popc(newm);
  seen_insert(c, isnew, key);
  c.nodes += nnew;
  iter_cnt += nnew;
  c.cands += (uint32_t)nnew;
  // merge (min, count of min, second distinct min, first lane at min) into the running state
  unsigned accm = __ballot_sync(kFull, acc);
  if (accm) {
    int v = acc ? mm : 255;
    int bmin = v;
#pragma unroll
    for (int o = 16; o; o >>= 1) bmin = min(bmin, __shfl_xor_sync(kFull, bmin, o));
    unsigned minm = __ballot_sync(kFull, acc && mm == bmin);
    int v2 = (acc && mm > bmin) ? mm : 255;
#pragma unroll
    for (int o = 16; o; o >>= 1) v2 = min(v2, __shfl_xor_sync(kFull, v2, o));
    int bcnt = __popc(minm);
    if (bmin < c.low) {
      int fl = __ffs(minm) - 1;
      c.nxt = min(c.low, v2);
      c.low = bmin;
      c.inst = bcnt;
      c.hit_p = __shfl_sync(kFull, p, fl);
      c.hit_ent = __shfl_sync(kFull, ent, fl);
      c.hit_mm = bmin;
      c.hit_strand = s;
    } else if (bmin == c.low) {
      c.inst += bcnt;
      c.nxt = min(c.nxt, v2);
    } else {
      c.nxt = min(c.nxt, bmin);
    }
  }
  return stop_lane;
}

// Position of core number idx in the slide loop of SfxArrayV2.cpp:5836-5847; false if the loop
// would have ended before reaching it.
__device__ __forceinline__ bool core_layout(int L, int CL, int delta, int max_slides, int idx, int& ofs_out) {
  int cur = delta, ofs = 0;
  for (int i = 0;; ++i) {
    if (!(i < max_slides && ofs <= L - CL && cur > CL / 3)) return false;
    if (ofs + CL + cur > L) cur = L - (ofs + CL);
    if (i == idx) { ofs_out = ofs; return true; }
    ofs += cur;
  }
}

// One phase = LocateCoreMultiples(max_mm, CL, delta).  Returns tHRslt.
__device__ __forceinline__ int run_phase(const DevIndex& I, const KParams& P, WarpCtx& c, int max_mm, int CL,
                                         int delta, int max_slides) {
  if (c.inst > P.max_hits && c.low == 0) return BKX_HR_HITINSTS;
  if (c.inst >= 1 && c.low == 0 && (c.nxt - c.low) < P.mmd) return BKX_HR_MMDELTA;
  if (c.inst <= 0 || c.low < 0 || c.nxt < 0) {
    c.inst = 0;
    c.low = c.nxt = max_mm + P.mmd + 1;
  }
  const int inst0 = c.inst, low0 = c.low, nxt0 = c.nxt;
  // number of cores of this phase (same for both strands)
  int n_cores = 0;
  {
    int o;
    bool v = core_layout(c.L, CL, delta, max_slides, c.lane, o);
    unsigned m = __ballot_sync(kFull, v);
    n_cores = __popc(m);
    if (n_cores == 32) {  // long reads: count the rest
      int i = 32;
      while (core_layout(c.L, CL, delta, max_slides, i, o)) ++i;
      n_cores = i;
    }
  }
  bool stop_all = false;
  const int s_begin = (P.strand_mode == BKX_STRAND_CRICK) ? 1 : 0;
  const int s_end = (P.strand_mode == BKX_STRAND_WATSON) ? 0 : 1;
  for (int s = s_begin; s <= s_end && !stop_all; ++s) {
    c.seen_n = 0;
    c.overflow = false;
    c.nodes = 0;
    bool stop_strand = false;
    for (int base = 0; base < n_cores && !stop_strand && !stop_all; base += 32) {
      // ---- seeds: lane = core
      int my = base + c.lane;
      int cofs = 0;
      uint64_t first = 0, cnt = 0;
      bool have = my < n_cores && core_layout(c.L, CL, delta, max_slides, my, cofs);
      if (have) locate_core(I, c, s, cofs, CL, first, cnt);
      int nc = min(32, n_cores - base);
      // ---- walk
      uint64_t cmax = cnt;
#pragma unroll
      for (int o = 16; o; o >>= 1) cmax = max(cmax, __shfl_xor_sync(kFull, cmax, o));
      int cores_done = nc;  // cores of this chunk whose LocateFirstExact the reference would issue
      if (cmax == 0) {
        // nothing to walk
      } else if (cmax <= 100) {
        // flattened: every interval is short enough that no cap can trigger
        int incl = (int)cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(kFull, incl, o);
          if (c.lane >= o) incl += t;
        }
        int total = __shfl_sync(kFull, incl, 31);
        c.pre[c.lane + 1] = incl;
        if (c.lane == 0) c.pre[0] = 0;
        __syncwarp();
        for (int e0 = 0; e0 < total; e0 += 32) {
          int e = e0 + c.lane;
          bool valid = e < total;
          int ci = 0;
          if (valid) {  // core owning flattened entry e: largest ci with pre[ci] <= e
            int lo = 0, hi = nc - 1;
            while (lo < hi) {
              int mid = (lo + hi + 1) >> 1;
              if (c.pre[mid] <= e) lo = mid; else hi = mid - 1;
            }
            ci = lo;
          }
          uint64_t cf = __shfl_sync(kFull, first, ci);
          int co = __shfl_sync(kFull, cofs, ci);
          uint64_t sidx = cf + (uint64_t)(e - c.pre[ci]);
          int iter_dummy = 0;
          bool sc = false;
          int sl = walk_step(I, P, c, s, max_mm, valid, co, sidx, false, iter_dummy, 0, sc, stop_strand, stop_all);
          if (stop_all || stop_strand) {
            cores_done = __shfl_sync(kFull, ci, sl < 0 ? 0 : sl) + 1;
            break;
          }
        }
        __syncwarp();
      } else {
        // some core of this chunk has > 100 copies: strictly core by core
        for (int ci = 0; ci < nc && !stop_all && !stop_strand; ++ci) {
          uint64_t cf = __shfl_sync(kFull, first, ci);
          uint64_t cc = __shfl_sync(kFull, cnt, ci);
          int co = __shfl_sync(kFull, cofs, ci);
          if (cc == 0) continue;
          int iter_cnt = 0;
          bool stop_core = false;
          const uint64_t hi_idx = cf + cc - 1;
          for (uint64_t e0 = 0; e0 < cc && !stop_core; e0 += 32) {
            uint64_t e = e0 + (uint64_t)c.lane;
            walk_step(I, P, c, s, max_mm, e < cc, co, cf + e, true, iter_cnt, hi_idx, stop_core, stop_strand, stop_all);
          }
          if (stop_all || stop_strand) cores_done = ci + 1;
        }
      }
      c.seeds += (uint32_t)cores_done;
    }
  }
  // return-code logic, SfxArrayV2.cpp:6237-6261
  if (c.low == low0 && c.inst == inst0) {
    if (nxt0 > c.nxt) return (c.nxt - c.low) < P.mmd ? BKX_HR_MMDELTA : BKX_HR_RMMDELTA;
    c.nxt = nxt0;
    return BKX_HR_NONE;
  }
  if (c.inst >= 1 && (c.nxt - c.low) < P.mmd) return BKX_HR_MMDELTA;
  if (c.inst > P.max_hits) return BKX_HR_HITINSTS;
  return BKX_HR_HITS;
}

}  // namespace bkx
