// Device-resident index layout and the small inline accessors every kernel shares.
//
// HBM layout (one copy per GPU; replaces the mmap'd tsSfxBlock, libbiokanga/SfxArrayV2.h:98-104):
//   g2   2 bits/base, 32 bases per u64, base i of the concatenation at bits [2*(i%32), +2) of word i/32
//        (A0 C1 G2 T3; N is stored as 0 and EOS as 3, both flagged in gx)          n/4 bytes
//   gx   1 bit/base "not ACGT" (N or EOS)                                             n/8 bytes
//   gxc  1 bit per 64-base block: block holds any flagged base (L2 resident)          n/512 bytes
//   sa   the reference's suffix array as written by `biokanga index`: u32 per element (sa_lo); for 5-byte elements
//        (n >= 4e9) the 5-byte elements back to back, exactly as in the file (sa5): one element = one 128-byte line
//        fetch (a u32 plane + a u8 plane cost two)                                    4n / 5n bytes
//   pt   k-mer prefix table: pt[x] = number of suffixes whose first-k symbols sort below the ACGT
//        k-mer x (first base most significant); 4^k+1 entries of u32.  For n >= 2^32 the u32 entries are relative to
//        their block of 4096 entries, whose absolute starts (u64, 4^k/4096 of them: L2 resident) sit in pt_hi
//   ent  chromosome table sorted by start offset + a coarse block -> entry lookup table
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bkx {

struct DevIndex {
  const uint64_t* g2;
  const uint64_t* gx;
  const uint32_t* gxc;
  const uint32_t* sa_lo;  // 4-byte elements, or the low plane of borrowed planes
  const uint8_t* sa_hi;   // high plane of borrowed planes (bkx_open_index_planes below 2^32 symbols), else nullptr
  const uint8_t* sa5;     // 5-byte elements back to back (8-byte aligned, 16 bytes of slack); set instead of sa_lo
  const uint32_t* pt32;   // prefix table; relative to pt_hi[x >> kPtBlockShift] when pt_hi is set
  const uint64_t* pt_hi;
  const uint64_t* ent_start;
  const uint64_t* ent_end;
  const uint32_t* ent_id;
  const uint32_t* ent_of_id;  // entry id -> index into the sorted ent_* arrays (0xffffffff if unknown)
  uint32_t max_ent_id;
  const uint32_t* ent_lut;  // ent_lut[ofs >> lut_shift] = first entry whose end_ofs >= (block start)
  uint32_t lut_shift;
  uint64_t n;  // ConcatSeqLen
  uint32_t n_ent;
  int k;
};

constexpr int kPtBlockShift = 12;   // two-level prefix table: 4096 entries per block

// Loads that ask L2 for 64 bytes of a missing line instead of all 128 (ld.global.nc.L2::64B; profiles/probes/gather_modes.cu:
// 63.8 instead of 127.3 DRAM bytes per random gather, same gather rate).  For kernels that sit on the DRAM bandwidth with
// single-element gathers -- the wave path's table and suffix-array lookups (5.0 TB/s before, profiles/r02_wave_ncu.md).
__device__ __forceinline__ uint32_t ldg_half_u32(const uint32_t* p) {
  uint32_t v;
  asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t ldg_half_u64(const uint64_t* p) {
  uint64_t v;
  asm("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_half_u8(const uint8_t* p) {
  uint32_t v;
  asm("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ uint64_t sa_get(const DevIndex& I, uint64_t i) {
  if (I.sa5) {   // 40 bits at byte 5i: one aligned 64-bit word, two when the element runs over its end
    const uint64_t a = i * 5;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(I.sa5) + (a >> 3);
    const unsigned sh = (unsigned)(a & 7) * 8;
    uint64_t v = __ldg(w) >> sh;
    if (sh > 24) v |= __ldg(w + 1) << (64 - sh);
    return v & 0xffffffffffull;
  }
  uint64_t v = __ldg(I.sa_lo + i);
  if (I.sa_hi) v |= (uint64_t)__ldg(I.sa_hi + i) << 32;
  return v;
}

__device__ __forceinline__ uint64_t pt_get(const DevIndex& I, uint64_t x) {
  uint64_t v = __ldg(I.pt32 + x);
  if (I.pt_hi) v += __ldg(I.pt_hi + (x >> kPtBlockShift));
  return v;
}

// sa_get / pt_get with the 64-byte loads above
__device__ __forceinline__ uint64_t sa_get_half(const DevIndex& I, uint64_t i) {
  if (I.sa5) {
    const uint64_t a = i * 5;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(I.sa5) + (a >> 3);
    const unsigned sh = (unsigned)(a & 7) * 8;
    uint64_t v = ldg_half_u64(w) >> sh;
    if (sh > 24) v |= ldg_half_u64(w + 1) << (64 - sh);
    return v & 0xffffffffffull;
  }
  uint64_t v = ldg_half_u32(I.sa_lo + i);
  if (I.sa_hi) v |= (uint64_t)ldg_half_u8(I.sa_hi + i) << 32;
  return v;
}
__device__ __forceinline__ uint64_t pt_get_half(const DevIndex& I, uint64_t x) {
  uint64_t v = ldg_half_u32(I.pt32 + x);
  if (I.pt_hi) v += __ldg(I.pt_hi + (x >> kPtBlockShift));
  return v;
}

// 32 bases of the concatenation starting at base g (little-endian base order).
__device__ __forceinline__ uint64_t gword(const DevIndex& I, uint64_t g) {
  uint64_t w = g >> 5;
  unsigned sh = (unsigned)(g & 31) * 2;
  uint64_t a = __ldg(I.g2 + w);
  if (sh == 0) return a;
  uint64_t b = __ldg(I.g2 + w + 1);
  return (a >> sh) | (b << (64 - sh));
}

// 4-bit symbol of the reference at concatenation offset i: 0..3, 4 = N, 7 = EOS (also past the end).
__device__ __forceinline__ int gsym(const DevIndex& I, uint64_t i) {
  if (i >= I.n) return 7;
  int code = (int)((__ldg(I.g2 + (i >> 5)) >> ((i & 31) * 2)) & 3);
  int ex = (int)((__ldg(I.gx + (i >> 6)) >> (i & 63)) & 1);
  return ex ? (code == 3 ? 7 : 4) : code;
}

// true if any base in [g, g+len) is N/EOS or lies past the end of the concatenation.  len <= 2000 (cMaxSeqLen): the
// 64-base blocks of the span are at most 33 bits of the coarse map, taken as one 64-bit window (gxc has two words of slack).
__device__ __forceinline__ bool span_has_exc(const DevIndex& I, uint64_t g, uint32_t len) {
  if (g + len > I.n) return true;
  const uint64_t b0 = g >> 6, b1 = (g + len - 1) >> 6;
  const uint64_t w = b0 >> 5;
  const uint64_t v = ((uint64_t)__ldg(I.gxc + w) | ((uint64_t)__ldg(I.gxc + w + 1) << 32)) >> (unsigned)(b0 & 31);
  const unsigned nb = (unsigned)(b1 - b0) + 1;   // 1..33 blocks, all inside the 64 - (b0 & 31) >= 33 bits of the window
  return (v & ((1ull << nb) - 1)) != 0;
}

// a / b for 0 <= a, 1 <= b, both below 2^15: exact through one float division ((a + 0.5) / b is at least 0.5 / b away from an
// integer, the division's error is far below that)
__device__ __forceinline__ int small_div(int a, int b) { return (int)__fdividef((float)a + 0.5f, (float)b); }

// reverse the order of the 32 two-bit groups of a word (first base becomes most significant).
__device__ __forceinline__ uint64_t rev2(uint64_t w) {
  uint64_t r = __brevll(w);
  return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}

// spread the 32 bits of x to the even bit positions of a 64-bit word.
__device__ __forceinline__ uint64_t spread32(uint32_t x) {
  uint64_t v = x;
  v = (v | (v << 16)) & 0x0000ffff0000ffffull;
  v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
  v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}

// chromosome (entry) index containing concatenation offset ofs, or -1 (MapChunkHit2Entry,
// libbiokanga/SfxArrayV2.cpp:2530-2575).
__device__ __forceinline__ int find_entry(const DevIndex& I, uint64_t ofs) {
  if (ofs >= I.n) return -1;
  uint32_t e = __ldg(I.ent_lut + (ofs >> I.lut_shift));
  while (e < I.n_ent && __ldg(I.ent_end + e) < ofs) ++e;
  if (e >= I.n_ent || __ldg(I.ent_start + e) > ofs) return -1;  // ofs sits on a terminator
  return (int)e;
}

}  // namespace bkx
