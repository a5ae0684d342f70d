// Fast path of the read aligner: ONE LANE PER READ, persistent lanes, step-synchronous state machine.
//
// Every lane carries its own read through the reference's sequential algorithm (AlignReads phases ->
// LocateCoreMultiples cores -> interval walk -> Hamming; libbiokanga/SfxArrayV2.cpp:7666-7760,
// 5693-6262) exactly as the CPU does it, one core ("seed") per step; the 32 lanes of a warp execute
// the same step body on 32 different reads, so the dependent loads of 32 reads are in flight together
// and every warp instruction serves up to 32 reads.
//
// The fast path only keeps the cases whose sequential semantics need no bookkeeping beyond a handful
// of registers; anything else DEFERS the read, untouched, to the general group kernel
// (bkx_align.cuh), which redoes it from scratch:
//   * reads longer than kFastMaxLen, reads holding an N that still pass the N filter;
//   * a core whose SA interval holds more than kFastMaxCnt suffixes (repeats: the 100th-candidate
//     probe / MaxIter caps can never trigger below that);
//   * more than kFastHashCap distinct candidate loci in one strand of one phase (the first kFastSeen live in
//     shared memory, the rest in a lane-private epoch-tagged hash set in HBM);
//   * a genome window that touches an N or a chromosome end (symbol-wise compare needed).
// So the result of a read never depends on which kernel produced it.
#pragma once
#include "bkx_align.cuh"

namespace bkx {

constexpr int kFastMaxLen = 320;   // bases; 10 words + pad per strand
constexpr int kFastMaxCnt = 64;    // SA interval size handled in the fast path (must stay below 100, see above)
constexpr int kFastSeen = 24;      // "already processed" keys per lane kept in shared memory
constexpr int kFastHashSlots = 1024;  // per-lane overflow set in HBM: (epoch << 32 | key) slots, open addressing
constexpr int kFastHashCap = 512;     // keys per strand/phase before the read is deferred
constexpr int kFastWarps = 8;      // warps per block
constexpr int kFastThreads = kFastWarps * 32;

__host__ __device__ inline size_t fast_smem_bytes(int W) {
  // per warp: 2 strands x W words x 32 lanes x 8 B (lane-interleaved) + (kFastSeen keys + the hash-set epoch) x 32
  // lanes x 4 B
  return (size_t)kFastWarps * ((size_t)2 * W * 32 * 8 + (size_t)(kFastSeen + 1) * 32 * 4);
}

struct FastLane {
  const uint64_t* w2[2];  // this lane's packed read: word w of strand s at w2[s][w * 32]
  uint32_t* seen;         // this lane's seen keys: key i at seen[i * 32]; seen[kFastSeen * 32] = hash-set epoch
  int L;
};

__device__ __forceinline__ uint64_t fl_word(const FastLane& f, int s, int pos) {
  int w = pos >> 5;
  unsigned sh = (unsigned)(pos & 31) * 2;
  uint64_t a = f.w2[s][w * 32];
  if (sh == 0) return a;
  return (a >> sh) | (f.w2[s][(w + 1) * 32] << (64 - sh));
}

// -1 / 0 / +1 of core (strand s, ofs, len) against the pure-ACGT suffix at g (caller checked the span)
__device__ __forceinline__ int fl_cmp(const DevIndex& I, const FastLane& f, int s, int ofs, int len, uint64_t g) {
  uint64_t w = g >> 5;
  unsigned sh = (unsigned)(g & 31) * 2;
  uint64_t prev = __ldg(I.g2 + w);
  for (int b = 0; b < len; b += 32) {
    uint64_t next = __ldg(I.g2 + (++w));
    uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    uint64_t rw = fl_word(f, s, ofs + b);
    uint64_t x = rw ^ gw;
    int rem = len - b;
    if (rem < 32) x &= (1ull << (2 * rem)) - 1;
    if (x) {
      int pos = (__ffsll((long long)x) - 1) >> 1;
      return ((rw >> (2 * pos)) & 3) > ((gw >> (2 * pos)) & 3) ? 1 : -1;
    }
  }
  return 0;
}

// result codes of one fast step
enum { FS_OK = 0, FS_DEFER = 1 };

}  // namespace bkx
