// The read-alignment kernel body: one G-lane group per read (G = 8: four reads per warp), persistent grid.
//
// Reference loops replaced (file:line under /root/reference):
//   CAligner::ProcCoredApprox        biokanga/Aligner.cpp:9027-9504   per-read driver + classification
//   CSfxArrayV3::AlignReads          libbiokanga/SfxArrayV2.cpp:7666-7760  staged phases
//   CSfxArrayV3::LocateCoreMultiples libbiokanga/SfxArrayV2.cpp:5693-6262  one phase (cores, walk, Hamming)
//   LocateFirstExact/LocateLastExact libbiokanga/SfxArrayV2.cpp:7765-8027  SA interval of a core
//
// Mapping to the lane group:
//   * the read is 2-bit packed (both strands) into the group's shared-memory slot once;
//   * SEEDS: the cores of BOTH strands of the current phase form one list; lane j owns seed j of the
//     current chunk of G: k-mer prefix-table bucket, then a lower/upper-bound refinement over the
//     reference's suffix array -> interval [first, first+cnt);
//   * WALK: per strand, the SA entries of the chunk's cores are flattened in the reference's
//     processing order (core, then SA index) and handed out G per step, lane = one candidate locus:
//     entry check, "already seen" test, XOR+popc Hamming over packed 64-bit words, then group
//     reductions keep (LowMMCnt, NxtLowMMCnt, LowHitInstances, first hit) exactly as the sequential
//     loop would, including the ordered early exit, the 100th-candidate copy probe and the MaxIter cap.
//
// Why the parallel walk is exact.  For one phase the reference's final (LowMMCnt, LowHitInstances,
// NxtLowMMCnt) are functions of the MULTISET of accepted mismatch counts of the candidates it
// processes: low = min, instances = multiplicity of min, next = second distinct value (else the initial
// MaxTotMM+MMDelta+1); its ">= NxtLowMMCnt" abort (SfxArrayV2.cpp:6150) only skips candidates that
// cannot change any of the three.  WHICH candidates are processed depends on order only through
// (a) the "already processed" set, (b) the 100th-candidate probe / MaxIter / node caps and (c) the
// early exit at the (MaxHits+1)-th exact match -- all three are evaluated here in lane order, which is
// the reference's processing order.  With MaxHits == 1 the stored hit is the first candidate at the
// final minimum, again in that order.
#pragma once
#include "../../include/bkx.h"
#include "bkx_index.cuh"

namespace bkx {

constexpr int kSeenCap = 96;         // "already processed" keys kept in shared memory per group
constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr int kGroup = 8;            // lanes per read

struct KParams {
  int max_subs, mmd, max_ns, strand_mode, max_hits, min_core_len, slides_per100, max_iter, max_nodes;
  int ml_mode, clamp_ml;
  int best;               // -N: LocateBestMatches instead of the staged AlignReads (general kernel only)
  int xdedup;             // fast kernel: "already processed" decided by comparing the earlier cores with the window (no key set)
  int prefetch;           // fast kernel: prefetch the next core's prefix-table entry into L2 (1) or not (0)
  int scan_iters;         // fast kernel: cores a lane may run down per step looking for a non-empty bucket (0: one core per step)
  bkx_multi_hit* multi;   // -r5: max_hits slots per read of this launch, or nullptr
};

// The reads of a launch once more, 2-bit packed (base i of the launch's concatenation at bits [2((i + phase) % 32), +2) of
// word (i + phase) / 32): the fast kernel takes a read from here when flags says it holds nothing but A C G T.
struct Packed2Src {
  const uint64_t* words = nullptr;   // 8-byte aligned, 16 readable bytes beyond the last base; nullptr: not given
  const uint8_t* flags = nullptr;    // per read: != 0 -- take the read from the one-byte-per-base copy; nullptr: no such read
  uint32_t phase = 0;
};

// What ProcCoredApprox stores for one read once AlignReads has returned (Aligner.cpp:9239-9245, 9310-9479): shared by
// both kernels.  `inst`, `low`, `nxt` are the search state, the hit_* values describe the first hit found at `low`.
static __device__ __noinline__ bkx_read_result make_result(const DevIndex& I, const KParams& P, int hr, int inst, int low,
                                                       int nxt, int L, int hit_strand, int hit_ent, uint64_t hit_p,
                                                       int hit_mm, uint32_t seeds, uint32_t cands) {
  bkx_read_result res;
  res.nar = BKX_NAR_NOHIT; res.strand = 0; res.num_hits = 0; res.low_mm = 0; res.nxt_low_mm = 0;
  res.low_hit_instances = 0; res.chrom_id = 0; res.match_loci = 0; res.match_len = 0; res.mismatches = 0;
  res.flags = 0; res.seeds = seeds; res.cands = cands; res.reserved = 0;
  if (inst > P.max_hits) inst = P.max_hits + 1;                                   // :9241
  if (P.clamp_ml && hr == BKX_HR_HITINSTS) { inst = P.max_hits; hr = BKX_HR_HITS; }  // :9243
  res.hit_rslt = (uint8_t)hr;
  switch (hr) {
    case BKX_HR_HITS:
      if (inst == 1 || P.ml_mode == BKX_ML_DEFAULT || P.ml_mode == BKX_ML_ALL) {  // unique, or every locus is wanted
        res.nar = BKX_NAR_ACCEPTED;
        res.num_hits = (P.ml_mode == BKX_ML_ALL) ? (uint8_t)min(inst, 255) : 1;   // beyond 255 loci: low_hit_instances has the count
        res.strand = hit_strand ? '-' : '+';
        res.chrom_id = __ldg(I.ent_id + hit_ent);
        res.match_loci = (uint32_t)(hit_p - __ldg(I.ent_start + hit_ent));
        res.match_len = (uint16_t)L;
        res.mismatches = (uint8_t)hit_mm;
        res.low_hit_instances = (P.ml_mode == BKX_ML_ALL) ? (int16_t)inst : 1;
      } else {                                      // -r1 / -r3 / -r4: counted, not placed (:9383-9397)
        res.nar = BKX_NAR_MULTIALIGN;
        res.low_hit_instances = (int16_t)inst;
      }
      res.low_mm = (int8_t)low;
      res.nxt_low_mm = (int8_t)nxt;
      break;
    case BKX_HR_MMDELTA:
    case BKX_HR_HITINSTS:
      res.nar = (hr == BKX_HR_MMDELTA) ? BKX_NAR_MMDELTA : BKX_NAR_MULTIALIGN;
      res.strand = '?';
      res.match_len = (uint16_t)L;
      res.low_hit_instances = (int16_t)inst;
      res.low_mm = (int8_t)low;
      res.nxt_low_mm = (int8_t)nxt;
      break;
    case BKX_HR_RMMDELTA:
      res.nxt_low_mm = (int8_t)nxt;
      break;
    default:
      break;
  }
  return res;
}

// overflow hash sets for groups whose strand/phase sees more than kSeenCap keys (high-copy repeats)
struct HashPool {
  uint64_t* tables;      // n_tables x slots u64: (epoch << 32) | key
  uint32_t* locks;       // 0 = free
  uint32_t* epochs;      // per table, monotonically increasing
  uint32_t n_tables;
  uint32_t slots;        // power of two
};

template <int G>
struct Grp {
  static constexpr unsigned kLaneMask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  // read, both strands, in shared memory
  uint64_t* s2[2];
  uint32_t* sx[2];
  uint32_t* seen;
  int* pre;
  int L;
  bool hasN;
  int gl;             // lane within the group
  unsigned gmask;     // this group's lanes within the warp
  int gshift;
  // "already processed" set of the current strand
  int seen_n;
  bool overflow;
  uint64_t* hash;
  uint32_t hmask, epoch;
  int table_id;
  int nodes;
  // phase state (LowHitInstances, LowMMCnt, NxtLowMMCnt) and first hit
  int inst, low, nxt;
  int hit_strand, hit_ent, hit_mm;
  uint64_t hit_p;
  bkx_multi_hit* multi;   // this read's -r5 slots or nullptr
  uint32_t seeds, cands;
  // -N (LocateBestMatches): the mismatch limit as tightened by a full list, "a locus was turned away from a full list"
  int bm_max;
  bool bm_sloughed;

  __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(gmask, p) >> gshift) & kLaneMask; }
  template <typename T>
  __device__ __forceinline__ T bcast(T v, int src) const { return __shfl_sync(gmask, v, src, G); }
  __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
  __device__ __forceinline__ int gmin(int v) const {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v = min(v, __shfl_xor_sync(gmask, v, o, G));
    return v;
  }
  __device__ __forceinline__ uint64_t gmax64(uint64_t v) const {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) { uint64_t t = __shfl_xor_sync(gmask, v, o, G); v = t > v ? t : v; }
    return v;
  }
  __device__ __forceinline__ int gsum(int v) const {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
    return v;
  }
};

template <int G>
__device__ __forceinline__ uint64_t read_word(const Grp<G>& c, int s, int pos) {
  int w = pos >> 5;
  unsigned sh = (unsigned)(pos & 31) * 2;
  uint64_t a = c.s2[s][w];
  if (sh == 0) return a;
  return (a >> sh) | (c.s2[s][w + 1] << (64 - sh));
}

template <int G>
__device__ __forceinline__ int rsym(const Grp<G>& c, int s, int i) {
  if ((c.sx[s][i >> 5] >> (i & 31)) & 1) return 4;
  return (int)((c.s2[s][i >> 5] >> ((i & 31) * 2)) & 3);
}

// probe-vs-suffix comparison of SfxArrayV2.cpp:7792-7811: -1 / 0 / +1, target EOS => -1.
template <int G>
__device__ __noinline__ int cmp_core_slow(const DevIndex& I, const Grp<G>& c, int s, int ofs, int len, uint64_t g) {
  for (int i = 0; i < len; ++i) {
    int gs = gsym(I, g + i);
    if (gs == 7) return -1;
    int rs = rsym(c, s, ofs + i);
    if (rs > gs) return 1;
    if (rs < gs) return -1;
  }
  return 0;
}

template <int G>
__device__ __forceinline__ int cmp_core(const DevIndex& I, const Grp<G>& c, int s, int ofs, int len, uint64_t g) {
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
template <int G>
__device__ __forceinline__ void locate_core(const DevIndex& I, const Grp<G>& c, int s, int ofs, int len,
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
template <int G>
__device__ __forceinline__ int hamming(const DevIndex& I, const Grp<G>& c, int s, uint64_t p, int max_mm) {
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
template <int G>
__device__ __forceinline__ bool seen_contains(const Grp<G>& c, uint32_t key) {
  if (!c.overflow) {
    bool f = false;
    for (int i = 0; i < c.seen_n; ++i) f |= (c.seen[i] == key);
    return f;
  }
  uint64_t want = ((uint64_t)c.epoch << 32) | key;
  uint32_t h = (key * 2654435761u) & c.hmask;
  for (;;) {
    uint64_t v = __ldcg(c.hash + h);  // L2 view: other lanes of this group insert with atomics
    if (v == want) return true;
    if ((uint32_t)(v >> 32) != c.epoch) return false;
    h = (h + 1) & c.hmask;
  }
}

template <int G>
__device__ __forceinline__ void hash_insert(Grp<G>& c, uint32_t key) {
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

// borrow one overflow table from the pool (lane 0 spins on the lock words; holders never wait)
template <int G>
__device__ __forceinline__ void hash_acquire(Grp<G>& c, const HashPool& hp, uint32_t salt) {
  int id = -1;
  uint32_t ep = 0;
  if (c.gl == 0) {
    uint32_t t = salt % hp.n_tables;
    for (;;) {
      if (atomicCAS(hp.locks + t, 0u, 1u) == 0u) break;
      t = (t + 1 == hp.n_tables) ? 0 : t + 1;
    }
    __threadfence();
    id = (int)t;
    ep = __ldcg(hp.epochs + t) + 1;
    __stcg(hp.epochs + t, ep);
  }
  c.table_id = c.bcast(id, 0);
  c.epoch = c.bcast(ep, 0);
  c.hash = hp.tables + (size_t)c.table_id * hp.slots;
  c.hmask = hp.slots - 1;
}

template <int G>
__device__ __forceinline__ void hash_release(Grp<G>& c, const HashPool& hp) {
  if (c.table_id < 0) return;
  c.sync();
  if (c.gl == 0) {
    __threadfence();
    atomicExch(hp.locks + c.table_id, 0u);
  }
  c.table_id = -1;
}

// all lanes call; lanes with ins==true add their (distinct) key
template <int G>
__device__ __forceinline__ void seen_insert(Grp<G>& c, const HashPool& hp, bool ins, uint32_t key) {
  unsigned m = c.ballot(ins);
  int cnt = __popc(m);
  if (cnt == 0) return;
  if (!c.overflow && c.seen_n + cnt <= kSeenCap) {
    if (ins) c.seen[c.seen_n + __popc(m & ((1u << c.gl) - 1))] = key;
    c.seen_n += cnt;
    c.sync();
    return;
  }
  if (!c.overflow) {  // migrate the shared-memory list into a global hash set
    c.overflow = true;
    hash_acquire(c, hp, key ^ (uint32_t)c.seen_n ^ (uint32_t)(c.hit_p >> 3));
    c.sync();
    for (int i = c.gl; i < c.seen_n; i += G) hash_insert(c, c.seen[i]);
    c.sync();
  }
  if (ins) hash_insert(c, key);
  c.seen_n += cnt;
  __threadfence_block();
  c.sync();
}

// ---- one step of the interval walk: up to G SA entries, one per lane, in processing order ------
// valid/cofs/sidx are lane-private.  capped => all lanes belong to one core whose interval ends at
// hi_idx and iter_cnt counts its new candidates so far (the 100th-candidate probe and MaxIter cap
// of SfxArrayV2.cpp:5857-5875 apply).  Returns the lane at which processing stopped (or -1).
template <int G, bool BEST>
__device__ __forceinline__ int walk_step(const DevIndex& I, const KParams& P, const HashPool& hp, Grp<G>& c, int s,
                                         int max_mm, bool valid, int cofs, uint64_t sidx, bool capped, int& iter_cnt,
                                         uint64_t hi_idx, bool& stop_core, bool& stop_strand, bool& stop_all) {
  const unsigned lt = (1u << c.gl) - 1;
  uint64_t loci = valid ? sa_get(I, sidx) : 0;
  bool ok = valid && loci >= (uint64_t)cofs;
  uint64_t p = loci - (uint64_t)cofs;
  int ent = -1;
  if (ok && BEST) {
    // LocateBestMatches only asks for the window to end inside the concatenation (SfxArrayV2.cpp:6839-6841); a window that
    // runs over a chromosome end is counted as a candidate and dies in the Hamming loop at the terminator
    ok = p + (uint64_t)c.L <= I.n;
    if (ok) ent = find_entry(I, p);
  } else if (ok) {
    ent = find_entry(I, p);
    ok = ent >= 0 && (p + (uint64_t)c.L - 1) <= __ldg(I.ent_end + ent);
  }
  uint32_t key = (uint32_t)(1u + (uint32_t)loci - (uint32_t)cofs);
  bool dup = ok && seen_contains(c, key);
  unsigned okm = c.ballot(ok);
  unsigned same = (__match_any_sync(c.gmask, key) >> c.gshift) & Grp<G>::kLaneMask;
  if (ok && (same & okm & lt)) dup = true;  // an earlier lane of this step already claims the key
  bool isnew = ok && !dup;
  unsigned newm = c.ballot(isnew);
  int stop_lane = -1;
  // caps, in processing order
  if (newm) {
    int rank = iter_cnt + __popc(newm & (lt | (1u << c.gl)));  // 1-based index among new candidates
    int cut = G;
    if (capped) {
      bool at100 = isnew && rank == 100 && sidx < hi_idx && (hi_idx - sidx + 2) > (uint64_t)P.max_iter;
      bool atmax = isnew && rank == P.max_iter;
      unsigned cm = c.ballot(at100 || atmax);
      if (cm) { cut = __ffs(cm) - 1; stop_core = true; }
    }
    // identifier-node budget (cMaxNumIdentNodes): the candidate that fills it is still processed
    int allowed = P.max_nodes - c.nodes;
    unsigned upto = (cut >= G - 1) ? Grp<G>::kLaneMask : ((2u << cut) - 1);
    if (__popc(newm & upto) >= allowed) {
      int nl = (int)__fns(newm, 0, allowed);  // lane of the allowed-th new candidate
      if (nl < cut) cut = nl;
      stop_core = true;
      stop_strand = true;
    }
    if (cut < G) {
      stop_lane = cut;
      if (c.gl > cut) isnew = false;
      newm = c.ballot(isnew);
    }
  }
  int mm = 255;
  if (isnew) mm = hamming(I, c, s, p, max_mm);
  bool acc = isnew && mm <= max_mm;
  if constexpr (BEST) {
    // ---- -N: keep the max_hits loci with the fewest mismatches, equal ones in discovery order (SfxArrayV2.cpp:6936-6986).
    // The candidates of this step are taken in lane order; lane 0 owns the list (this read's slots in global memory).
    const int nnew = __popc(newm);
    seen_insert(c, hp, isnew, key);
    c.nodes += nnew;
    iter_cnt += nnew;
    c.cands += (uint32_t)nnew;
    unsigned accm = c.ballot(acc && ent >= 0);
    bkx_multi_hit h;
    h.chrom_id = 0; h.match_loci = 0; h.match_len = (uint16_t)c.L; h.strand = s ? '-' : '+'; h.mismatches = (uint8_t)mm;
    if (acc && ent >= 0) {
      h.chrom_id = __ldg(I.ent_id + ent);
      h.match_loci = (uint32_t)(p - __ldg(I.ent_start + ent));
    }
    while (accm) {
      const int l = __ffs(accm) - 1;
      accm &= accm - 1;
      const int lmm = c.bcast(mm, l);
      const uint32_t l_chrom = c.bcast(h.chrom_id, l), l_loci = c.bcast(h.match_loci, l);
      const uint64_t l_p = c.bcast(p, l);
      const int l_ent = c.bcast(ent, l);
      if (lmm > c.bm_max) continue;                    // the tightened limit: its Hamming loop would have given up
      if (c.inst == P.max_hits) c.bm_sloughed = true;
      // position: behind every kept locus with at most this many mismatches; none if the list is full of such
      int pos = -1;
      if (c.multi) {
        int at = 0;
        if (c.gl == 0) {
          at = c.inst;
          for (int b = 0; b < c.inst; ++b)
            if ((int)c.multi[b].mismatches > lmm) { at = b; break; }
          if (at < P.max_hits) {
            int last = min(c.inst, P.max_hits - 1);          // the last one falls off a full list
            for (int b = last; b > at; --b) c.multi[b] = c.multi[b - 1];
            bkx_multi_hit nh;
            nh.chrom_id = l_chrom; nh.match_loci = l_loci; nh.match_len = (uint16_t)c.L; nh.strand = s ? '-' : '+';
            nh.mismatches = (uint8_t)lmm;
            c.multi[at] = nh;
          }
        }
        at = c.bcast(at, 0);
        pos = at < P.max_hits ? at : -1;
      } else {
        // no list wanted (-r1): only the number kept and the first of the best matter
        pos = (c.inst == 0 || lmm < c.hit_mm) ? 0 : (c.inst < P.max_hits ? c.inst : -1);
      }
      if (pos >= 0) {
        if (pos == 0) { c.hit_p = l_p; c.hit_ent = l_ent; c.hit_mm = lmm; c.hit_strand = s; }
        if (c.inst < P.max_hits) c.inst += 1;
        else if (c.multi) {   // inserted into a full list: the limit drops to the worst one kept
          int worst = 0;
          if (c.gl == 0) worst = (int)c.multi[c.inst - 1].mismatches;
          c.bm_max = c.bcast(worst, 0);
        }
      }
    }
    c.sync();
    return stop_lane;
  }
  // ordered early exit: the (MaxHits+1)-th exact match ends the whole search (SfxArrayV2.cpp:6206-6214)
  unsigned zm = c.ballot(acc && mm == 0);
  if (zm) {
    int before = (c.low == 0) ? c.inst : 0;
    if (before + __popc(zm) > P.max_hits) {
      int need = P.max_hits + 1 - before;
      int el = (int)__fns(zm, 0, need);
      if (c.gl > el) { isnew = false; acc = false; }
      newm = c.ballot(isnew);
      stop_all = true;
      stop_core = true;
      stop_lane = el;
    }
  }
  int nnew = __popc(newm);
  seen_insert(c, hp, isnew, key);
  c.nodes += nnew;
  iter_cnt += nnew;
  c.cands += (uint32_t)nnew;
  // merge (min, count of min, second distinct min, first lane at min) into the running state
  unsigned accm = c.ballot(acc);
  if (accm) {
    int bmin = c.gmin(acc ? mm : 255);
    unsigned minm = c.ballot(acc && mm == bmin);
    int v2 = c.gmin((acc && mm > bmin) ? mm : 255);
    int bcnt = __popc(minm);
    if (c.multi && bmin <= c.low) {  // -r5: the hit list, in discovery order = lane order within the step (:6157-6205)
      const int at = (bmin < c.low ? 0 : c.inst) + __popc(minm & ((1u << c.gl) - 1));
      if (acc && mm == bmin && at < P.max_hits) {
        bkx_multi_hit h;
        h.chrom_id = __ldg(I.ent_id + ent);
        h.match_loci = (uint32_t)(p - __ldg(I.ent_start + ent));
        h.match_len = (uint16_t)c.L;
        h.strand = s ? '-' : '+';
        h.mismatches = (uint8_t)mm;
        c.multi[at] = h;
      }
    }
    if (bmin < c.low) {
      int fl = __ffs(minm) - 1;
      c.nxt = min(c.low, v2);
      c.low = bmin;
      c.inst = bcnt;
      c.hit_p = c.bcast(p, fl);
      c.hit_ent = c.bcast(ent, fl);
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

// Closed form of the slide loop of SfxArrayV2.cpp:5836-5847 for delta >= CL (always true on this
// path: staged phases use delta == CL, the final phase CoreDelta = max(.., CoreLen)):
// cores sit at 0, delta, .., K*delta; core K is the one whose step gets shortened to
// r = L-(K*delta+CL); one more flush-right core at L-CL follows iff r > CL/3.
struct CoreLayout {
  int K, n, last_ofs, delta;
  __device__ __forceinline__ CoreLayout(int L, int CL, int d, int max_slides) {
    delta = d;
    K = (L - CL - d >= 0) ? (L - CL - d) / d + 1 : 0;
    int r = L - (K * d + CL);
    n = K + 1 + ((r > CL / 3) ? 1 : 0);
    if (n > max_slides) n = max_slides;
    if (L < CL) n = 0;  // the loop's `ofs <= L-CL` test fails at once
    last_ofs = L - CL;
  }
  __device__ __forceinline__ int ofs(int i) const { return i <= K ? i * delta : last_ofs; }
};

// One phase = LocateCoreMultiples(max_mm, CL, delta).  Returns tHRslt.
template <int G, bool BEST>
__device__ __forceinline__ int run_phase(const DevIndex& I, const KParams& P, const HashPool& hp, Grp<G>& c,
                                         int max_mm, int CL, int delta, int max_slides) {
  constexpr bool best = BEST;   // -N: this one call is the whole search (LocateBestMatches, SfxArrayV2.cpp:6654-7019)
  if (!best) {
    if (c.inst > P.max_hits && c.low == 0) return BKX_HR_HITINSTS;
    if (c.inst >= 1 && c.low == 0 && (c.nxt - c.low) < P.mmd) return BKX_HR_MMDELTA;
    if (c.inst <= 0 || c.low < 0 || c.nxt < 0) {
      c.inst = 0;
      c.low = c.nxt = max_mm + P.mmd + 1;
    }
  } else {
    c.inst = 0;
    c.bm_max = max_mm;
    c.bm_sloughed = false;
  }
  const int inst0 = c.inst, low0 = c.low, nxt0 = c.nxt;
  const CoreLayout lay(c.L, CL, delta, max_slides);
  const int n_cores = lay.n;
  const int s_first = (P.strand_mode == BKX_STRAND_CRICK) ? 1 : 0;
  const int n_strands = (P.strand_mode == BKX_STRAND_BOTH) ? 2 : 1;
  const int total = n_cores * n_strands;
  bool stop_all = false, stop_strand = false;
  int cur_strand = -1;
  for (int base = 0; base < total && !stop_all; base += G) {
    // ---- seeds: lane = (strand, core) of the combined list
    const int my = base + c.gl;
    const bool have = my < total;
    const int my_s = have ? s_first + my / n_cores : 0;
    const int my_c = have ? my % n_cores : 0;
    const int cofs = lay.ofs(my_c);
    uint64_t first = 0, cnt = 0;
    if (have && !(stop_strand && my_s == cur_strand)) locate_core(I, c, my_s, cofs, CL, first, cnt);
    const int nc = min(G, total - base);
    // ---- walk, strand by strand inside the chunk
    int lane0 = 0;
    while (lane0 < nc && !stop_all) {
      const int s = s_first + (base + lane0) / n_cores;
      const int core0 = (base + lane0) % n_cores;
      const int lanes = min(nc - lane0, n_cores - core0);  // lanes of this strand in this chunk
      if (s != cur_strand) {  // strand start: reset the "already processed" set (SfxArrayV2.cpp:5834-5835)
        hash_release(c, hp);
        cur_strand = s;
        c.seen_n = 0;
        c.overflow = false;
        c.nodes = 0;
        stop_strand = false;
      }
      if (stop_strand) { lane0 += lanes; continue; }
      const bool mine = c.gl >= lane0 && c.gl < lane0 + lanes;
      const uint64_t mycnt = mine ? cnt : 0;
      const uint64_t cmax = c.gmax64(mycnt);
      int cores_done = lanes;  // cores of this strand/chunk whose LocateFirstExact the reference issues
      if (cmax == 0) {
        // nothing to walk
      } else if (cmax <= 100 && !best) {
        // flattened: every interval is short enough that no cap can trigger
        int incl = (int)mycnt;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
          int t = __shfl_up_sync(c.gmask, incl, o, G);
          if (c.gl >= o) incl += t;
        }
        const int tot_e = c.bcast(incl, G - 1);
        c.pre[c.gl + 1] = incl;
        if (c.gl == 0) c.pre[0] = 0;
        c.sync();
        for (int e0 = 0; e0 < tot_e; e0 += G) {
          const int e = e0 + c.gl;
          const bool valid = e < tot_e;
          int ci = lane0;
          if (valid) {  // lane owning flattened entry e: largest ci with pre[ci] <= e
            int lo = lane0, hi = lane0 + lanes - 1;
            while (lo < hi) {
              int mid = (lo + hi + 1) >> 1;
              if (c.pre[mid] <= e) lo = mid; else hi = mid - 1;
            }
            ci = lo;
          }
          const uint64_t cf = c.bcast(first, ci);
          const int co = c.bcast(cofs, ci);
          const uint64_t sidx = cf + (uint64_t)(e - c.pre[ci]);
          int iter_dummy = 0;
          bool sc = false;
          int sl = walk_step<G, BEST>(I, P, hp, c, s, max_mm, valid, co, sidx, false, iter_dummy, 0, sc, stop_strand, stop_all);
          if (stop_all || stop_strand) {
            cores_done = c.bcast(ci, sl < 0 ? 0 : sl) - lane0 + 1;
            break;
          }
        }
        c.sync();
      } else {
        // some core of this strand/chunk has > 100 copies: strictly core by core
        for (int ci = lane0; ci < lane0 + lanes && !stop_all && !stop_strand; ++ci) {
          const uint64_t cf = c.bcast(first, ci);
          const uint64_t cc = c.bcast(cnt, ci);
          const int co = c.bcast(cofs, ci);
          if (cc == 0) continue;
          int iter_cnt = 0;
          bool stop_core = false;
          const uint64_t hi_idx = cf + cc - 1;
          for (uint64_t e0 = 0; e0 < cc && !stop_core; e0 += G) {
            uint64_t e = e0 + (uint64_t)c.gl;
            walk_step<G, BEST>(I, P, hp, c, s, best ? c.bm_max : max_mm, e < cc, co, cf + e, true, iter_cnt, hi_idx, stop_core,
                               stop_strand, stop_all);
          }
          // -N: a full list of exact matches that turned nothing away cannot improve (SfxArrayV2.cpp:6988-6992)
          if (best && c.inst == P.max_hits && c.bm_max == 0 && !c.bm_sloughed) stop_all = true;
          if (stop_all || stop_strand) cores_done = ci - lane0 + 1;
        }
      }
      c.seeds += (uint32_t)cores_done;
      lane0 += lanes;
    }
  }
  hash_release(c, hp);
  if (best) return c.inst >= 1 ? BKX_HR_HITS : BKX_HR_NONE;   // Aligner.cpp:9211-9217
  // return-code logic, SfxArrayV2.cpp:6237-6261
  if (c.low == low0 && c.inst == inst0) {
    if (nxt0 > c.nxt) return (c.nxt - c.low) < P.mmd ? BKX_HR_MMDELTA : BKX_HR_RMMDELTA;
    return BKX_HR_NONE;
  }
  if (c.inst >= 1 && (c.nxt - c.low) < P.mmd) return BKX_HR_MMDELTA;
  if (c.inst > P.max_hits) return BKX_HR_HITINSTS;
  return BKX_HR_HITS;
}

}  // namespace bkx
