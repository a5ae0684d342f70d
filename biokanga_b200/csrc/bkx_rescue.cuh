// Orphan-mate recovery: one warp per orphan pair.
//
// Reference (file:line under /root/reference):
//   CAligner::ProcessPairedEnds, orphan branches   biokanga/Aligner.cpp:3218-3477
//   CSfxArrayV3::AlignPairedRead                    libbiokanga/SfxArrayV2.cpp:8247-8433
//   CSfxArrayV3::AdaptiveTrim                       libbiokanga/SfxArrayV2.cpp:5482-5682
//   CSfxArrayV3::IterateExactsRange                 libbiokanga/SfxArrayV2.cpp:3382-3474
//
// AlignPairedRead looks for the unaligned mate inside [anchor+MinDist, anchor+MaxDist] (or mirrored
// upstream).  Windows under 1000 loci are scanned linearly, one AdaptiveTrim per locus; here the
// genome window (<= 999+L bases, 2-bit codes + N/EOS mask) is staged ONCE in shared memory and the 32
// lanes each evaluate a locus per step -- the one place on this path where a tile is reused.
// Larger windows use the suffix array: exact matches of each core inside the window are evaluated.
//
// AdaptiveTrim is only ever called here with MinTrimLen == SeqLen (MinChimericLen is 0 on this path),
// where it reduces to a predicate on the mismatch bitmap M of the full-length alignment:
//   accept  <=>  M has a zero run >= 8 (cMinATExactLen), bits 0..2 and L-3..L-1 are zero (MinFlank 3),
//                and popc(M) * 100 < (MaxMM + 1) * L   (the reference's double ratio, exact in integers:
//                both operands < 2^18), with MaxMM == 0 demanding popc(M) == 0;
//   and it returns popc(M) as the mismatch count.  (Derivation in DESIGN.md section 3.)
#pragma once
#include "bkx_align.cuh"

namespace bkx {

constexpr int kRescueWarps = 4;
constexpr int kRescueThreads = kRescueWarps * 32;
constexpr int kRescueMaxLen = 2000;   // cMaxSeqLen

// per-warp shared memory (u64 words): read codes RW, read N mask RW/2 (u32 per 32 bases), window codes WW,
// window N/EOS bitmap WW/2+1
__host__ __device__ inline int rescue_rw(int Lmax) { return (Lmax + 31) / 32 + 1; }
__host__ __device__ inline int rescue_ww(int Lmax) { return (1000 + Lmax + 63) / 32 + 3; }
__host__ __device__ inline size_t rescue_warp_bytes(int Lmax) {
  return ((size_t)rescue_rw(Lmax) + (rescue_rw(Lmax) + 1) / 2 + rescue_ww(Lmax) + rescue_ww(Lmax) / 2 + 2) * 8;
}

// compress the even bits of x (one flag per 2-bit group) into the low 32 bits
__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
  x &= 0x5555555555555555ull;
  x = (x | (x >> 1)) & 0x3333333333333333ull;
  x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
  x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
  x = (x | (x >> 16)) & 0x00000000ffffffffull;
  return (uint32_t)x;
}

// 32 two-bit codes starting at base `pos` of a packed array in shared memory
__device__ __forceinline__ uint64_t sm_word2(const uint64_t* a, int pos) {
  int w = pos >> 5;
  unsigned sh = (unsigned)(pos & 31) * 2;
  uint64_t v = a[w];
  if (sh) v = (v >> sh) | (a[w + 1] << (64 - sh));
  return v;
}
// 32 one-bit flags starting at base `pos` of a bitmap held 64 flags per u64 word
__device__ __forceinline__ uint32_t sm_bits(const uint64_t* a, int pos) {
  int w = pos >> 6;
  unsigned sh = (unsigned)(pos & 63);
  uint64_t v = a[w] >> sh;
  if (sh > 32) v |= a[w + 1] << (64 - sh);
  return (uint32_t)v;
}

// The full-length AdaptiveTrim predicate, streamed over the 32-bit words of the mismatch bitmap.
struct ATFull {
  int mm = 0;
  uint32_t prev_z = 0, prev_m = 0, cur_m = 0;
  bool run8 = false, head_bad = false;
  int nw = 0;
  __device__ __forceinline__ void push(uint32_t m, int i, int L) {  // m already masked to the valid bases
    mm += __popc(m);
    if (i == 0 && (m & 7u)) head_bad = true;  // first three bases must match (MinFlankMatches 3)
    uint32_t z = ~m;
    int rem = L - 32 * i;
    if (rem < 32) z &= (1u << rem) - 1;
    // 39-bit view: top 7 flags of the previous word followed by this word -> any run of 8 matches
    uint64_t v = ((uint64_t)z << 7) | (uint64_t)(prev_z >> 25);
    uint64_t a = v & (v >> 1);
    a &= a >> 2;
    a &= a >> 4;
    if (a) run8 = true;
    prev_z = z;
    prev_m = cur_m;
    cur_m = m;
    nw = i + 1;
  }
  // mismatch count if accepted, else -1
  __device__ __forceinline__ int result(int L, int max_mm_rate) const {
    if (L < 25 || L > 2048) return -1;  // cMinATSeqLen / cMaxATSeqLen
    if (head_bad || !run8) return -1;
    if (max_mm_rate == 0 ? (mm != 0) : (mm * 100 >= (max_mm_rate + 1) * L)) return -1;
    for (int q = L - 3; q < L; ++q) {  // last three bases must match
      int w = q >> 5;
      uint32_t m = (w == nw - 1) ? cur_m : prev_m;
      if ((m >> (q & 31)) & 1u) return -1;
    }
    return mm;
  }
};

}  // namespace bkx
