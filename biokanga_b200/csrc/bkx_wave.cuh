// Wave path of the read aligner: the search of bkx_fast.cuh turned inside out.
//
// The lane-per-read kernel carries one read through its phases, cores, interval walk and Hamming loop in ONE thread: every
// DRAM line of a read is fetched behind the previous one, and a warp runs with a third of its lanes because its 32
// reads stand at different places of that state machine (profiles/r02_final_align_fast_ncu.md).  Here the same work is
// laid out by KIND instead of by read -- all reads that are in search phase `allow` go through two streaming kernels
// together, once per phase ("round"):
//
//   wave_step   one thread per read: (a) what the read's previous phase found -- "already processed" (one placement reached
//               through several cores), lowest / next-lowest mismatch count and instance count -- and then either the
//               read's result (ProcCoredApprox's mapping, Aligner.cpp:9239-9479) or its next phase (AlignReads' staged
//               loop, SfxArrayV2.cpp:7666-7760); (b) the prefix-table lookups of all cores of the phase the read is in
//               now -- independent loads, four cores at a time -- and an ITEM (read, strand, core, bucket) for every
//               bucket that is not empty;
//   wave_probe  one thread per item: bound refinement in the bucket, interval walk, Hamming of every placement against
//               the read; a CANDIDATE record (placement, strand, mismatches) per placement into the read's row.
//
// What one phase of LocateCoreMultiples (SfxArrayV2.cpp:5693-6262) leaves behind -- LowHitInstances, LowMMCnt, NxtLowMMCnt
// and the hit when it is unique -- does not depend on the order in which placements are met, with one exception: the early
// exit once more than MaxHits exact placements are known (:6199).  A read that gets there, and every read the lane-per-read
// kernel would hand on as well (an N or a chromosome end in a window, an interval of more than kFastMaxCnt suffixes, more
// placements than its row holds, reads that did not arrive 2-bit packed), is put on a list and redone from scratch by
// align_fast_kernel; so the result of a read still never depends on which kernel produced it.
//
// Measured (profiles/r02_experiments.md): 27.5 ms against 29.5-30.4 ms per 20 M reads at configs[1], 46.8 ms against 63 ms
// at configs[3] (14 G symbols), but 3.7 ms against 2.8 ms per 2 M reads -- launches of kWaveAutoReads reads and more take
// this path.  What it took: no list append through a returning atomic (one per warp bounded the first version at 2.9 ns
// each -- items now go into chunks of the queue owned by a warp, and reads stay in place with a state byte instead of
// being compacted), items that carry the read's position, length and phase (no dependent loads in front of the
// suffix-array fetch), the read staged in shared memory.  Gathering the suffix-array element in a kernel of its own
// (sa_split) and 1536 threads per SM for wave_probe both lost.
//
// Only the default search is laid out this way (no multi-loci option, no -N): those keep the lane-per-read kernel.
#pragma once
#include "bkx_fast.cuh"

namespace bkx {

constexpr int kWaveMaxRounds = 66;        // staged phases 0..63 + the final one (+1)
constexpr int kWaveCntItems = 80;         // counters: item slots of wave i at [80 + i]
constexpr int kWaveCntFallback = 160;     //           reads handed to align_fast_kernel
constexpr int kWaveCounters = 176;
constexpr unsigned kWaveFailed = 127;     // candidate record: more mismatches than the phase allows
constexpr unsigned kWaveOff = 0xffu;      // state byte: the read is not (or no longer) on the wave path
constexpr uint32_t kWaveAutoReads = 6000000;   // launches of at least this many reads take the wave path by default
constexpr int kWaveChunk = 256;           // item slots a warp takes from the queue at a time (>= 4 x 32)
constexpr unsigned long long kWaveNoItem = ~0ull;   // items[].y of an empty slot

struct WaveBuf {
  unsigned int* cnt = nullptr;            // kWaveCounters counters
  uint8_t* ph = nullptr;                  // per read: allow (bits 0..6) | final phase (bit 7); kWaveOff: not on the wave path
  uint8_t* fb = nullptr;                  // per read: must be redone by align_fast_kernel
  uint2* acc = nullptr;                   // per read: seeds, candidates of its finished phases
  uint32_t* ncand = nullptr;              // per read: candidate records of the current phase
  uint64_t* cand = nullptr;               // per read: `row` records: placement (40 bits) | strand << 40 | mismatches << 41
  ulonglong4* items = nullptr;            // x: bucket start (40 bits) | strand << 40 | core << 41;  y: read | bucket size << 32;
                                          // z: the read's first base in the 2-bit stream (44 bits) | length << 44 | state << 56
  uint32_t* fb_ids = nullptr;
  uint64_t item_cap = 0;
  int row = 8;
  int half_loads = 1;                     // 64-byte L2 fetches for the table, suffix-array and genome gathers (BKX_WAVE_HALF=0: off)
  int sa_split = 0;                       // 1: the first suffix-array element of every item is gathered by its own kernel
                                          // (measured: 5.1 + 13.5 ms instead of 15.2 ms in one kernel -- off)
};

struct ReadRef { const uint64_t* words; uint64_t bo; int L; };

// 32 bases of the forward strand from pos (0 <= pos), zero beyond the read's end
__device__ __forceinline__ uint64_t rr_fwd(const ReadRef& q, int pos) {
  if (pos >= q.L) return 0;
  const uint64_t b = q.bo + (uint64_t)pos;
  const uint64_t* w = q.words + (b >> 5);
  const unsigned sh = (unsigned)(b & 31) * 2;
  uint64_t v = __ldg(w) >> sh;
  if (sh) v |= __ldg(w + 1) << (64 - sh);
  const int rem = q.L - pos;
  if (rem < 32) v &= (1ull << (2 * rem)) - 1;
  return v;
}
// the same of strand s (1: the reverse complement)
__device__ __forceinline__ uint64_t rr_word(const ReadRef& q, int s, int pos) {
  if (s == 0) return rr_fwd(q, pos);
  if (pos >= q.L) return 0;
  const int t = q.L - 32 - pos;   // forward position of the last base of this word
  const uint64_t fw = t >= 0 ? rr_fwd(q, t) : (rr_fwd(q, 0) << (2 * (-t)));
  uint64_t rc = rev2(~fw);
  const int rem = q.L - pos;
  if (rem < 32) rc &= (1ull << (2 * rem)) - 1;
  return rc;
}
// -1 / 0 / +1 of core (strand s, ofs, len) against the pure-ACGT suffix at g (fl_cmp of bkx_fast.cuh on a ReadRef)
__device__ __forceinline__ int rr_cmp(const DevIndex& I, const ReadRef& q, int s, int ofs, int len, uint64_t g) {
  uint64_t w = g >> 5;
  const unsigned sh = (unsigned)(g & 31) * 2;
  uint64_t prev = __ldg(I.g2 + w);
  for (int b = 0; b < len; b += 32) {
    const uint64_t next = __ldg(I.g2 + (++w));
    const uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    const uint64_t rw = rr_word(q, s, ofs + b);
    uint64_t x = rw ^ gw;
    const int rem = len - b;
    if (rem < 32) x &= (1ull << (2 * rem)) - 1;
    if (x) {
      const int pos = (__ffsll((long long)x) - 1) >> 1;
      return ((rw >> (2 * pos)) & 3) > ((gw >> (2 * pos)) & 3) ? 1 : -1;
    }
  }
  return 0;
}

// fl_cmp of bkx_fast.cuh with 64-byte genome loads (HALF), for the one-compare-per-item pattern of wave_probe
template <bool HALF>
__device__ __forceinline__ uint64_t wv_g2(const DevIndex& I, uint64_t w) {
  if constexpr (HALF) return ldg_half_u64(I.g2 + w);
  else return __ldg(I.g2 + w);
}
template <bool HALF>
__device__ __forceinline__ uint64_t wv_sa(const DevIndex& I, uint64_t i) {
  if constexpr (HALF) return sa_get_half(I, i);
  else return sa_get(I, i);
}
template <bool HALF>
__device__ __forceinline__ int wv_cmp(const DevIndex& I, const FastLane& f, int s, int ofs, int len, uint64_t g) {
  uint64_t w = g >> 5;
  const unsigned sh = (unsigned)(g & 31) * 2;
  uint64_t prev = wv_g2<HALF>(I, w);
  for (int b = 0; b < len; b += 32) {
    const uint64_t next = wv_g2<HALF>(I, ++w);
    const uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    const uint64_t rw = fl_word(f, s, ofs + b);
    uint64_t x = rw ^ gw;
    const int rem = len - b;
    if (rem < 32) x &= (1ull << (2 * rem)) - 1;
    if (x) {
      const int pos = (__ffsll((long long)x) - 1) >> 1;
      return ((rw >> (2 * pos)) & 3) > ((gw >> (2 * pos)) & 3) ? 1 : -1;
    }
  }
  return 0;
}

// per-read search parameters (Aligner.cpp:9085-9095) and the phase they are in
struct WaveRead { int L, max_tot_mm, core_len, slides, core_delta; };
struct WavePhase { int CL, delta, mm_max, K, n_cores, last_ofs; };

__device__ __forceinline__ void wave_read(const KParams& P, int L, WaveRead& w) {
  w.L = L;
  w.max_tot_mm = P.max_subs == 0 ? 0 : max(1, (L * P.max_subs + 50) / 100);
  if (w.max_tot_mm > 63) w.max_tot_mm = 63;
  w.core_len = max(P.min_core_len, small_div(L, P.mmd == 1 ? w.max_tot_mm + 1 : w.max_tot_mm + 2));
  w.slides = max(1, (P.slides_per100 * L + 99) / 100);
  w.core_delta = max(small_div(L, w.slides) - 1, w.core_len);
}
__device__ __forceinline__ void wave_cores(const WaveRead& w, WavePhase& ph) {
  const int L = w.L;
  ph.K = (L - ph.CL - ph.delta >= 0) ? small_div(L - ph.CL - ph.delta, ph.delta) + 1 : 0;
  const int rr = L - (ph.K * ph.delta + ph.CL);
  ph.n_cores = ph.K + 1 + ((rr > ph.CL / 3) ? 1 : 0);
  if (ph.n_cores > w.slides) ph.n_cores = w.slides;
  if (L < ph.CL) ph.n_cores = 0;
  ph.last_ofs = L - ph.CL;
}
// the phase a state byte stands for (the state was produced by wave_enter, so the phase exists)
__device__ __forceinline__ void wave_phase(const KParams& P, const WaveRead& w, unsigned state, WavePhase& ph) {
  if (state & 0x80u) {
    ph.CL = w.core_len; ph.delta = w.core_delta; ph.mm_max = w.max_tot_mm;
  } else {
    const int allow = (int)(state & 0x7fu);
    ph.CL = small_div(w.L, allow + P.mmd); ph.delta = ph.CL; ph.mm_max = allow;
  }
  wave_cores(w, ph);
}
// AlignReads' staged loop: the phase for `allow` mismatches, or the final one; false when there is none
__device__ __forceinline__ bool wave_setup(const KParams& P, const WaveRead& w, int allow, bool& in_final, WavePhase& ph) {
  if (in_final) return false;
  bool staged = false;
  if (w.max_tot_mm > 0 && allow <= w.max_tot_mm) {
    const int cl = small_div(w.L, allow + P.mmd);
    if (cl > w.core_len) { staged = true; ph.CL = cl; ph.delta = cl; ph.mm_max = allow; }
  }
  if (!staged) {
    if (w.max_tot_mm > 0 && allow > w.max_tot_mm) return false;
    in_final = true;
    ph.CL = w.core_len; ph.delta = w.core_delta; ph.mm_max = w.max_tot_mm;
  }
  wave_cores(w, ph);
  return true;
}
// first phase with at least one core, starting at `allow` (fresh: a new read; otherwise the phase after state's).
// Returns the state byte, or -1: the read has no further phase (eHRnone)
__device__ __forceinline__ int wave_enter(const KParams& P, const WaveRead& w, int state, bool fresh) {
  int allow = fresh ? 0 : (state & 0x7f);
  bool in_final = fresh ? false : (state & 0x80) != 0;
  WavePhase ph;
  if (fresh) {
    for (;;) {
      if (!wave_setup(P, w, allow, in_final, ph)) return -1;
      if (ph.n_cores > 0) break;
      if (in_final) return -1;
      ++allow;
    }
  } else {
    for (;;) {
      if (in_final) return -1;
      ++allow;
      if (!wave_setup(P, w, allow, in_final, ph)) return -1;
      if (ph.n_cores > 0) break;
    }
  }
  return (allow & 0x7f) | (in_final ? 0x80 : 0);
}

// append to a device list: one atomic per warp
__device__ __forceinline__ unsigned wave_push(unsigned int* counter) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1));
}

}  // namespace bkx
