// CUDA kernels of the bkx library (sm_100a): index preparation, read alignment, paired-end pairing.
#include "bkx_align.cuh"
#include "bkx_fast.cuh"
#include "bkx_wave.cuh"
#include "bkx_rescue.cuh"
#include "bkx_kernels.h"

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <mutex>
#include <vector>

namespace bkx {

// ------------------------------------------------------------------------------------------------
// Index preparation
// ------------------------------------------------------------------------------------------------
// One thread per 64-base block of the reference's 1-byte/base concatenation: writes two g2 words, one
// gx word, and raises the block's coarse bit.  bad_count receives the number of symbols that are not
// one of A C G T N EOS (soft-mask bit set, InDel/Undefined codes): such files are rejected.
__global__ void pack_genome_kernel(const uint8_t* __restrict__ seq, uint64_t n, uint64_t* __restrict__ g2,
                                   uint64_t* __restrict__ gx, uint32_t* __restrict__ gxc,
                                   unsigned long long* __restrict__ bad_count) {
  uint64_t nblk = (n + 63) >> 6;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t base = b << 6;
    uint64_t w0 = 0, w1 = 0, x = 0;
    unsigned bad = 0;
    if (base + 64 <= n && ((uintptr_t)(seq + base) & 15) == 0) {
      const uint4* v = reinterpret_cast<const uint4*>(seq + base);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u = __ldg(v + q);
        uint32_t ws[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            unsigned sym = (ws[j] >> (8 * t)) & 0xff;
            int i = q * 16 + j * 4 + t;
            unsigned code = sym & 3, ex = 0;
            if (sym > 3) { ex = 1; code = (sym == 7) ? 3 : 0; bad += (sym != 4 && sym != 7); }
            if (i < 32) w0 |= (uint64_t)code << (2 * i); else w1 |= (uint64_t)code << (2 * (i - 32));
            x |= (uint64_t)ex << i;
          }
        }
      }
    } else {
      for (int i = 0; i < 64 && base + i < n; ++i) {
        unsigned sym = seq[base + i];
        unsigned code = sym & 3, ex = 0;
        if (sym > 3) { ex = 1; code = (sym == 7) ? 3 : 0; bad += (sym != 4 && sym != 7); }
        if (i < 32) w0 |= (uint64_t)code << (2 * i); else w1 |= (uint64_t)code << (2 * (i - 32));
        x |= (uint64_t)ex << i;
      }
    }
    g2[2 * b] = w0;
    g2[2 * b + 1] = w1;
    gx[b] = x;
    if (x) atomicOr(gxc + (b >> 5), 1u << (b & 31));
    if (bad) atomicAdd(bad_count, (unsigned long long)bad);
  }
}

// 5-byte suffix elements -> u32 low plane + u8 high plane.
__global__ void split_sa5_kernel(const uint8_t* __restrict__ sa5, uint64_t n, uint32_t* __restrict__ lo,
                                 uint8_t* __restrict__ hi) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t* p = sa5 + i * 5;
    lo[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    hi[i] = p[4];
  }
}

// Table key of the suffix at position i: its first k symbols as a base-4 number (first base most
// significant); a suffix whose j-th symbol (j<k) is N/EOS/past-the-end takes the key
// (prefix_j << 2(k-j)) | 11..1, i.e. it is filed at the very end of its j-symbol prefix -- exactly
// where the reference's symbol order (A<C<G<T<N<EOS) sorts it.
__device__ __forceinline__ uint64_t suffix_key(const DevIndex& I, uint64_t i, int k) {
  uint64_t w = i >> 5;
  unsigned sh = (unsigned)(i & 31) * 2;
  uint64_t a = I.g2[w];
  uint64_t gw = sh ? ((a >> sh) | (I.g2[w + 1] << (64 - sh))) : a;
  // exception bits of [i, i+k)
  uint64_t xw = i >> 6;
  unsigned xs = (unsigned)(i & 63);
  uint64_t xa = I.gx[xw] >> xs;
  if (xs) xa |= I.gx[xw + 1] << (64 - xs);
  uint64_t kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1);
  xa &= kmask;
  // positions past the end behave as EOS
  if (i + (uint64_t)k > I.n) xa |= (~0ull << (I.n - i)) & kmask;
  uint64_t key = rev2(gw) >> (64 - 2 * k);
  if (xa) {
    int j = __ffsll((long long)xa) - 1;
    key |= (1ull << (2 * (k - j))) - 1;
  }
  return key;
}

__global__ void kmer_hist_kernel(DevIndex I, int k, uint32_t* __restrict__ cnt) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < I.n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t key = suffix_key(I, i, k);
    atomicAdd(cnt + key + 1, 1u);
  }
}

// Two-level table (n >= 2^32): the counts of each block of 4096 entries are summed (pass 1), the block sums are scanned
// into the absolute block starts pt_hi (CUB), and each block is scanned in place into entries relative to its start
// (pass 2).  One thread block of 256 threads per table block, 16 entries per thread.
__global__ void pt_block_sums_kernel(const uint32_t* __restrict__ cnt, uint64_t entries, uint64_t* __restrict__ sums) {
  __shared__ unsigned long long warp_sum[8];
  const uint64_t nblk = (entries + (1u << kPtBlockShift) - 1) >> kPtBlockShift;
  for (uint64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
    const uint64_t base = b << kPtBlockShift;
    unsigned long long v = 0;
    for (int q = 0; q < 16; ++q) {
      const uint64_t x = base + (uint64_t)q * 256 + threadIdx.x;
      if (x < entries) v += cnt[x];
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < 8; ++w) t += warp_sum[w];
      sums[b] = t;
    }
    __syncthreads();
  }
}

__global__ void pt_block_scan_kernel(uint32_t* __restrict__ cnt, uint64_t entries) {
  __shared__ uint32_t part[256];
  const uint64_t nblk = (entries + (1u << kPtBlockShift) - 1) >> kPtBlockShift;
  for (uint64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
    const uint64_t base = (b << kPtBlockShift) + (uint64_t)threadIdx.x * 16;   // 16 consecutive entries per thread
    uint32_t v[16];
    uint32_t run = 0;
    for (int q = 0; q < 16; ++q) {
      const uint64_t x = base + q;
      run += x < entries ? cnt[x] : 0u;
      v[q] = run;   // inclusive within the thread's 16
    }
    part[threadIdx.x] = run;
    __syncthreads();
    // exclusive scan of the 256 per-thread totals (Hillis-Steele in shared memory)
    for (int o = 1; o < 256; o <<= 1) {
      const uint32_t add = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0u;
      __syncthreads();
      part[threadIdx.x] += add;
      __syncthreads();
    }
    const uint32_t before = part[threadIdx.x] - run;
    for (int q = 0; q < 16; ++q) {
      const uint64_t x = base + q;
      if (x < entries) cnt[x] = before + v[q];
    }
    __syncthreads();
  }
}

// Self-check of a finished index: every suffix-array element must sit inside the prefix-table bucket of the suffix it
// names.  Ties the three big arrays together (genome words -> key, table, suffix array): a damaged page in any of them
// shows up as elements outside their bucket.  The coarse exception map is checked against the fine one as well.
__global__ void verify_index_kernel(DevIndex I, int k, unsigned long long* __restrict__ n_bad) {
  unsigned long long bad = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < I.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = sa_get(I, i);
    if (p >= I.n) { ++bad; continue; }
    if (p + (uint64_t)k > I.n) continue;  // the last k suffixes: the reference's comparator runs off its buffer there
    const uint64_t key = suffix_key(I, p, k);
    if (!(pt_get(I, key) <= i && i < pt_get(I, key + 1))) ++bad;
  }
  // ... and the coarse exception map must say exactly which 64-base blocks hold an N / EOS (pack_genome_kernel sets
  // bit b of gxc whenever gx[b] != 0)
  const uint64_t nblk = (I.n + 63) >> 6;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (uint64_t)gridDim.x * blockDim.x) {
    const bool any = __ldg(I.gx + b) != 0;
    const bool bit = ((__ldg(I.gxc + (b >> 5)) >> (unsigned)(b & 31)) & 1u) != 0;
    if (any != bit) ++bad;
  }
  for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

cudaError_t launch_verify_index(const DevIndex& I, int k, unsigned long long* n_bad, cudaStream_t st) {
  verify_index_kernel<<<148 * 16, 256, 0, st>>>(I, k, n_bad);
  return cudaGetLastError();
}

cudaError_t launch_pack_genome(const uint8_t* seq, uint64_t n, uint64_t* g2, uint64_t* gx, uint32_t* gxc,
                               unsigned long long* bad, cudaStream_t st) {
  uint64_t nblk = (n + 63) >> 6;
  int grid = (int)((nblk + 255) / 256 < 148 * 16 ? (nblk + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  pack_genome_kernel<<<grid, 256, 0, st>>>(seq, n, g2, gx, gxc, bad);
  return cudaGetLastError();
}

// 4-bit packed reads (base 2i in the low nibble of byte i) -> one byte per base; `phase` = 1 when the first wanted
// base sits in a high nibble.  Four bases (one u32 store) per thread.
__global__ void unpack4_kernel(const uint8_t* __restrict__ packed, unsigned phase, uint64_t n, uint8_t* __restrict__ out) {
  const uint64_t quads = (n + 3) >> 2;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = 4 * q + phase;  // nibble index of the first base of this quad
    const uint64_t b = j >> 1;
    uint32_t w = (uint32_t)__ldg(packed + b) | ((uint32_t)__ldg(packed + b + 1) << 8) | ((uint32_t)__ldg(packed + b + 2) << 16);
    w >>= (j & 1) * 4;
    const uint32_t v = (w & 0xf) | ((w & 0xf0) << 4) | ((w & 0xf00) << 8) | ((w & 0xf000) << 12);
    if (4 * q + 4 <= n) {
      *(uint32_t*)(out + 4 * q) = v;
    } else {
      for (uint64_t i = 4 * q; i < n; ++i) out[i] = (uint8_t)(v >> (8 * (i - 4 * q)));
    }
  }
}

cudaError_t launch_unpack4(const uint8_t* packed, unsigned phase, uint64_t n_bases, uint8_t* out, cudaStream_t st) {
  const uint64_t quads = (n_bases + 3) >> 2;
  int grid = (int)std::min<uint64_t>((quads + 255) / 256, 148 * 16);
  unpack4_kernel<<<grid, 256, 0, st>>>(packed, phase, n_bases, out);
  return cudaGetLastError();
}

// ---- compact host interface (bkx_align_reads_packed2): 2 bits per base in, 16 bytes per read out -------------------------
// 2-bit packed reads (base i at bits [2(i%4), +2) of byte i/4) -> one byte per base.  `phase` = position of the first
// wanted base inside its byte.  Sixteen bases (one 16-byte store) per thread; `packed` and `out` start on 16-byte
// boundaries and are over-allocated by the caller, so whole words are read and written at both ends.
__global__ void unpack2_kernel(const uint32_t* __restrict__ packed, unsigned phase, uint64_t n, uint8_t* __restrict__ out) {
  const uint64_t groups = (n + 15) >> 4;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t w = __funnelshift_r(__ldg(packed + g), __ldg(packed + g + 1), 2 * phase);
    uint4 v;
    uint32_t* vo = &v.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t b = (w >> (8 * q)) & 0xffu;
      vo[q] = (b & 3u) | ((b & 0xcu) << 6) | ((b & 0x30u) << 12) | ((b & 0xc0u) << 18);
    }
    *reinterpret_cast<uint4*>(out + 16 * g) = v;
  }
}

// the bases that are not A C G T (stored as code 0 in the 2-bit stream) get their etSeqBase code back
__global__ void scatter_exceptions_kernel(const uint64_t* __restrict__ pos, const uint8_t* __restrict__ code, uint32_t n_exc,
                                          uint64_t first_base, uint64_t n_bases, uint8_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_exc; i += gridDim.x * blockDim.x) {
    const uint64_t p = __ldg(pos + i) - first_base;
    if (p < n_bases) out[p] = __ldg(code + i);
  }
}

// read offsets of a slice of fixed-length reads: offs[i] = i * len, i = 0..n_reads
__global__ void fixed_offsets_kernel(uint64_t* __restrict__ offs, uint32_t n_reads, uint32_t len) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_reads; i += (uint64_t)gridDim.x * blockDim.x)
    offs[i] = i * len;
}

// 32-byte records -> the 16-byte form that crosses PCIe (bkx_read_result16, include/bkx.h)
__global__ void compact_results_kernel(const bkx_read_result* __restrict__ in, uint32_t n, bkx_read_result16* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(in + i);
    const uint4 a = __ldg(src), b = __ldg(src + 1);
    bkx_read_result r;
    memcpy(&r, &a, 16);
    memcpy(reinterpret_cast<char*>(&r) + 16, &b, 16);
    bkx_read_result16 c;
    c.nar_hr = (uint8_t)((r.nar & 0x1f) | (r.hit_rslt << 5));
    const unsigned sc = r.strand == '+' ? 1u : r.strand == '-' ? 2u : r.strand == '?' ? 3u : 0u;
    c.strand_flags = (uint8_t)(sc | ((r.flags & 3u) << 2));
    c.num_hits = r.num_hits;
    c.mismatches = r.mismatches;
    c.low_mm = r.low_mm;
    c.nxt_low_mm = r.nxt_low_mm;
    c.low_hit_instances = r.low_hit_instances;
    c.chrom_id = r.chrom_id;
    c.match_loci = r.match_loci;
    uint4 o;
    memcpy(&o, &c, 16);
    *reinterpret_cast<uint4*>(out + i) = o;
  }
}

// flags[read] = 1 for every read that holds a non-ACGT base: the read of exception k is found by bisection of the slice's
// offsets (offs[0] = 0 .. offs[n_reads] = bases of the slice)
__global__ void flag_exception_reads_kernel(const uint64_t* __restrict__ pos, uint32_t n_exc, uint64_t first_base,
                                            const uint64_t* __restrict__ offs, uint32_t n_reads, uint8_t* __restrict__ flags) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_exc; i += gridDim.x * blockDim.x) {
    const uint64_t p = __ldg(pos + i) - first_base;
    if (p >= __ldg(offs + n_reads)) continue;
    uint32_t lo = 0, hi = n_reads;   // last read with offs[read] <= p
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(offs + mid) <= p) lo = mid; else hi = mid;
    }
    flags[lo] = 1;
  }
}

cudaError_t launch_flag_exception_reads(const uint64_t* pos, uint32_t n_exc, uint64_t first_base, const uint64_t* offs,
                                        uint32_t n_reads, uint8_t* flags, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(flags, 0, n_reads, st);
  if (e != cudaSuccess || n_exc == 0) return e;
  int grid = (int)std::min<uint32_t>((n_exc + 255) / 256, 148 * 8);
  flag_exception_reads_kernel<<<grid, 256, 0, st>>>(pos, n_exc, first_base, offs, n_reads, flags);
  return cudaGetLastError();
}

cudaError_t launch_unpack2(const uint8_t* packed, unsigned phase, uint64_t n_bases, uint8_t* out, cudaStream_t st) {
  const uint64_t groups = (n_bases + 15) >> 4;
  int grid = (int)std::min<uint64_t>((groups + 255) / 256, 148 * 16);
  if (grid < 1) grid = 1;
  unpack2_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(packed), phase, n_bases, out);
  return cudaGetLastError();
}

cudaError_t launch_scatter_exceptions(const uint64_t* pos, const uint8_t* code, uint32_t n_exc, uint64_t first_base,
                                      uint64_t n_bases, uint8_t* out, cudaStream_t st) {
  if (n_exc == 0) return cudaSuccess;
  int grid = (int)std::min<uint32_t>((n_exc + 255) / 256, 148 * 8);
  scatter_exceptions_kernel<<<grid, 256, 0, st>>>(pos, code, n_exc, first_base, n_bases, out);
  return cudaGetLastError();
}

cudaError_t launch_fixed_offsets(uint64_t* offs, uint32_t n_reads, uint32_t len, cudaStream_t st) {
  int grid = (int)std::min<uint32_t>(n_reads / 256 + 1, 148 * 8);
  fixed_offsets_kernel<<<grid, 256, 0, st>>>(offs, n_reads, len);
  return cudaGetLastError();
}

// offs[0] = 0, offs[i + 1] = lens[0] + .. + lens[i]
struct U16ToU64 {
  __host__ __device__ __forceinline__ uint64_t operator()(const uint16_t& v) const { return (uint64_t)v; }
};
cudaError_t launch_len_offsets(const uint16_t* lens, uint32_t n_reads, uint64_t* offs, void* tmp, size_t* tmp_bytes, cudaStream_t st) {
  cub::TransformInputIterator<uint64_t, U16ToU64, const uint16_t*> it(lens, U16ToU64());
  if (!tmp) return cub::DeviceScan::InclusiveSum(nullptr, *tmp_bytes, it, offs + 1, (int)n_reads, st);
  cudaError_t e = cudaMemsetAsync(offs, 0, 8, st);
  if (e != cudaSuccess) return e;
  return cub::DeviceScan::InclusiveSum(tmp, *tmp_bytes, it, offs + 1, (int)n_reads, st);
}

cudaError_t launch_compact_results(const bkx_read_result* in, uint32_t n, bkx_read_result16* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  int grid = (int)std::min<uint32_t>((n + 255) / 256, 148 * 16);
  compact_results_kernel<<<grid, 256, 0, st>>>(in, n, out);
  return cudaGetLastError();
}

cudaError_t launch_split_sa5(const uint8_t* sa5, uint64_t n, uint32_t* lo, uint8_t* hi, cudaStream_t st) {
  split_sa5_kernel<<<148 * 8, 256, 0, st>>>(sa5, n, lo, hi);
  return cudaGetLastError();
}

// pt[x] = number of suffixes whose key sorts below x: a histogram of the keys (at x + 1) and its inclusive prefix sum.
// block_starts == nullptr: one u32 sum over the whole table (n < 2^32).  Otherwise the two-level form described at
// pt_block_sums_kernel: table entries relative to their block, block_starts[b] = absolute value before block b.
cudaError_t build_prefix_table(const DevIndex& I, int k, uint32_t* table, uint64_t* block_starts, cudaStream_t st) {
  const uint64_t entries = (1ull << (2 * k)) + 1;
  cudaError_t e = cudaMemsetAsync(table, 0, entries * 4, st);
  if (e != cudaSuccess) return e;
  kmer_hist_kernel<<<148 * 16, 256, 0, st>>>(I, k, table);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  if (!block_starts) {
    e = cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, table, table, (long long)entries, st);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tmp, tmp_bytes);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, table, table, (long long)entries, st);
  } else {
    const uint64_t nblk = (entries + (1u << kPtBlockShift) - 1) >> kPtBlockShift;
    const int grid = (int)std::min<uint64_t>(nblk, 148 * 16);
    pt_block_sums_kernel<<<grid, 256, 0, st>>>(table, entries, block_starts);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    unsigned long long* bs = (unsigned long long*)block_starts;
    e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bs, bs, (long long)nblk, st);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tmp, tmp_bytes);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, bs, bs, (long long)nblk, st);
    if (e == cudaSuccess) {
      pt_block_scan_kernel<<<grid, 256, 0, st>>>(table, entries);
      e = cudaGetLastError();
    }
  }
  cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(tmp);
  return e != cudaSuccess ? e : e2;
}

// borrowed planes -> 5-byte elements back to back
__global__ void merge_sa5_kernel(const uint32_t* __restrict__ lo, const uint8_t* __restrict__ hi, uint64_t n, uint8_t* __restrict__ sa5) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = lo[i];
    uint8_t* p = sa5 + i * 5;
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
    p[4] = hi ? hi[i] : (uint8_t)0;
  }
}
cudaError_t launch_merge_sa5(const uint32_t* lo, const uint8_t* hi, uint64_t n, uint8_t* sa5, cudaStream_t st) {
  merge_sa5_kernel<<<148 * 8, 256, 0, st>>>(lo, hi, n, sa5);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Read alignment
// ------------------------------------------------------------------------------------------------
constexpr int kGroupsPerBlock = kBlockThreads / kGroup;

// The dynamic shared-memory opt-in of a kernel (cudaFuncAttributeMaxDynamicSharedMemorySize) belongs to the CURRENT
// DEVICE, so what has been configured is remembered per device; within a device it is only ever raised (several indexes
// and several host threads -- `bkx-align --gpus N` drives one thread per GPU -- share the kernels), under a lock.
struct SmemOptIn {
  std::mutex mtx;
  size_t configured[64] = {};
};
template <typename K>
static cudaError_t ensure_smem(K kernel, size_t smem, SmemOptIn& cfg) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::lock_guard<std::mutex> lk(cfg.mtx);
  if (smem > cfg.configured[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cfg.configured[dev] = smem;
  }
  return cudaSuccess;
}
static SmemOptIn g_smem_general, g_smem_general_best, g_smem_fast, g_smem_fast_mlx, g_smem_fast_scan, g_smem_fast_mlx_scan, g_smem_rescue;

__host__ __device__ inline size_t group_smem_bytes(int W) {
  size_t per = (size_t)W * 8 * 2 + (size_t)W * 4 * 2 + kSeenCap * 4 + (kGroup + 2) * 4;
  return (per + 15) & ~(size_t)15;
}
size_t align_smem_bytes(int W) { return group_smem_bytes(W) * kGroupsPerBlock; }

struct BlockStats {
  unsigned int nar[BKX_NAR_COUNT];
  unsigned int plus, minus;
  unsigned int multi, multi_loci;   // -r1: reads whose search ended eHRhits with several loci (classed ML); the loci
                                    // of all multi-loci reads (-r1 and -r5)
  unsigned int acc_multi;           // -r5: accepted reads carrying several loci
  unsigned long long seeds, cands;
};

__device__ __forceinline__ void stats_add_basic(BlockStats& bs, const bkx_read_result& res) {
  atomicAdd(&bs.nar[res.nar], 1u);
  if (res.nar == BKX_NAR_ACCEPTED) atomicAdd(res.strand == '+' ? &bs.plus : &bs.minus, 1u);
  atomicAdd(&bs.seeds, (unsigned long long)res.seeds);
  atomicAdd(&bs.cands, (unsigned long long)res.cands);
}
__device__ __forceinline__ void stats_add(BlockStats& bs, const bkx_read_result& res) {
  stats_add_basic(bs, res);
  if (res.nar == BKX_NAR_MULTIALIGN && res.hit_rslt == BKX_HR_HITS) {
    atomicAdd(&bs.multi, 1u);
    atomicAdd(&bs.multi_loci, (unsigned int)res.low_hit_instances);
  } else if (res.nar == BKX_NAR_ACCEPTED && res.num_hits > 1) {   // -r5: the count is in low_hit_instances (num_hits stops at 255)
    atomicAdd(&bs.acc_multi, 1u);
    atomicAdd(&bs.multi_loci, (unsigned int)res.low_hit_instances);
  }
}

// the per-thread counters ProcCoredApprox merges at Aligner.cpp:9507-9516
__device__ __forceinline__ void stats_flush(const BlockStats& bs, bkx_align_stats* stats) {
  for (int i = threadIdx.x; i < BKX_NAR_COUNT; i += blockDim.x)
    if (bs.nar[i]) atomicAdd((unsigned long long*)&stats->nar[i], (unsigned long long)bs.nar[i]);
  if (threadIdx.x == 0) {
    unsigned long long reads = 0;
    for (int i = 0; i < BKX_NAR_COUNT; ++i) reads += bs.nar[i];
    const unsigned long long acc = bs.nar[BKX_NAR_ACCEPTED];
    atomicAdd((unsigned long long*)&stats->plus_hits, (unsigned long long)bs.plus);
    atomicAdd((unsigned long long*)&stats->minus_hits, (unsigned long long)bs.minus);
    atomicAdd((unsigned long long*)&stats->num_sloughed_ns, (unsigned long long)bs.nar[BKX_NAR_NS]);
    atomicAdd((unsigned long long*)&stats->tot_non_aligned,
              (unsigned long long)bs.nar[BKX_NAR_NOHIT] + bs.nar[BKX_NAR_MULTIALIGN] - bs.multi);
    atomicAdd((unsigned long long*)&stats->tot_accepted_unique, acc - bs.acc_multi);
    atomicAdd((unsigned long long*)&stats->tot_accepted_multi, (unsigned long long)bs.multi + bs.acc_multi);
    atomicAdd((unsigned long long*)&stats->tot_accepted_aligned, acc + bs.multi);
    atomicAdd((unsigned long long*)&stats->tot_loci_aligned, acc - bs.acc_multi + bs.multi_loci);
    atomicAdd((unsigned long long*)&stats->tot_not_accepted_delta, (unsigned long long)bs.nar[BKX_NAR_MMDELTA]);
    atomicAdd((unsigned long long*)&stats->seeds, bs.seeds);
    atomicAdd((unsigned long long*)&stats->cands, bs.cands);
    atomicAdd((unsigned long long*)&stats->reads, reads);
  }
}

// Per-read driver: ProcCoredApprox body (Aligner.cpp:9027-9504) + AlignReads phase loop
// (SfxArrayV2.cpp:7666-7760).  One kGroup-lane group per read; reads are claimed from a global cursor.
template <bool BEST>
__global__ void __launch_bounds__(kBlockThreads, 3) align_reads_kernel(
    DevIndex I, KParams P, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offs, uint32_t n_reads,
    int W, bkx_read_result* __restrict__ out, bkx_align_stats* __restrict__ stats, unsigned int* __restrict__ cursor,
    HashPool hp, const uint32_t* __restrict__ ids, const unsigned int* __restrict__ n_ids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BlockStats bs;
  constexpr int G = kGroup;
  if (ids) n_reads = *n_ids;  // second pass: only the reads the fast kernel deferred
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (int)(sizeof(BlockStats) / 4); i += blockDim.x) ((unsigned int*)&bs)[i] = 0;
  __syncthreads();

  unsigned char* my = smem_raw + group_smem_bytes(W) * (threadIdx.x / G);
  Grp<G> c;
  c.s2[0] = (uint64_t*)my;
  c.s2[1] = c.s2[0] + W;
  c.sx[0] = (uint32_t*)(c.s2[1] + W);
  c.sx[1] = c.sx[0] + W;
  c.seen = c.sx[1] + W;
  c.pre = (int*)(c.seen + kSeenCap);
  c.gl = lane & (G - 1);
  c.gshift = lane & ~(G - 1);
  c.gmask = Grp<G>::kLaneMask << c.gshift;
  c.table_id = -1;
  c.hash = nullptr;
  c.hmask = 0;
  c.epoch = 0;
  c.hit_p = 0;

  for (;;) {
    unsigned int r = 0;
    if (c.gl == 0) r = atomicAdd(cursor, 1u);
    r = c.bcast(r, 0);
    if (r >= n_reads) break;
    if (ids) r = ids[r];
    const uint64_t o0 = __ldg(offs + r);
    const int L = (int)(__ldg(offs + r + 1) - o0);
    const uint8_t* rd = bases + o0;
    c.L = L;
    c.multi = P.multi ? P.multi + (size_t)r * P.max_hits : nullptr;
    // ---- unpack, N filter (Aligner.cpp:9041-9063), 2-bit pack both strands: one 32-base word per lane
    int nN = 0;
    bool bad = false;
    const int words = (L + 31) >> 5;  // words <= W-1: word `words` is the zero pad the shifts read
    for (int t = c.gl; t < 2 * (words + 1); t += G) {
      const int s = t > words ? 1 : 0;
      const int w = t - s * (words + 1);
      uint64_t code2 = 0;
      uint32_t nm = 0;
      const int i0 = w * 32;
      const int cntb = min(32, L - i0);
      for (int j = 0; j < cntb; ++j) {
        unsigned b = __ldg(rd + (s ? (L - 1 - (i0 + j)) : (i0 + j))) & 0x07;
        bad |= (b > 4);
        unsigned isn = (b == 4);
        unsigned code = (b < 4) ? (s ? 3 - b : b) : 0;
        code2 |= (uint64_t)code << (2 * j);
        nm |= isn << j;
      }
      c.s2[s][w] = code2;
      c.sx[s][w] = nm;
      if (s == 0) nN += __popc(nm);
    }
    nN = c.gsum(nN);
    bad = c.ballot(bad) != 0;
    c.sync();
    c.hasN = nN > 0;
    int max_ns_seq = 0;
    if (P.max_ns) max_ns_seq = max((L * P.max_ns) / 100, P.max_ns);

    bkx_read_result res;
    res.nar = BKX_NAR_NOHIT; res.hit_rslt = 0; res.strand = 0; res.num_hits = 0; res.low_mm = 0; res.nxt_low_mm = 0;
    res.low_hit_instances = 0; res.chrom_id = 0; res.match_loci = 0; res.match_len = 0; res.mismatches = 0;
    res.flags = 0; res.seeds = 0; res.cands = 0; res.reserved = 0;
    c.seeds = c.cands = 0;

    if (bad || nN > max_ns_seq || L < 1) {
      // the reference stops at the first offending base; a code > N or too many Ns both give eNARNs
      res.nar = BKX_NAR_NS;
    } else {
      // per-read search parameters, Aligner.cpp:9085-9095
      int max_tot_mm = P.max_subs == 0 ? 0 : max(1, (L * P.max_subs + 50) / 100);
      if (max_tot_mm > 63) max_tot_mm = 63;
      int core_len = max(P.min_core_len, L / (P.mmd == 1 ? max_tot_mm + 1 : max_tot_mm + 2));
      int slides = max(1, (P.slides_per100 * L + 99) / 100);
      int core_delta = max(L / slides - 1, core_len);
      c.inst = 0; c.low = 0; c.nxt = 0;
      c.hit_strand = 0; c.hit_ent = -1; c.hit_mm = 0; c.hit_p = 0;
      int hr = 0, allow = 0;
      if constexpr (BEST) {   // -N: one un-staged pass that keeps the best loci (Aligner.cpp:9197-9218); LowMMCnt / NxtLowMMCnt stay 0
        hr = run_phase<G, true>(I, P, hp, c, max_tot_mm, core_len, core_delta, slides);
        c.low = 0; c.nxt = 0;
      } else {
        if (max_tot_mm > 0) {
          for (allow = 0; allow <= max_tot_mm; ++allow) {
            int cl = L / (allow + P.mmd);
            if (cl <= core_len) break;
            hr = run_phase<G, false>(I, P, hp, c, allow, cl, cl, slides);
            if (hr != 0) break;
          }
        }
        if (hr == 0 && allow <= max_tot_mm) hr = run_phase<G, false>(I, P, hp, c, max_tot_mm, core_len, core_delta, slides);
      }
      res = make_result(I, P, hr, c.inst, c.low, c.nxt, L, c.hit_strand, c.hit_ent, c.hit_p, c.hit_mm, c.seeds, c.cands);
    }
    if (c.gl == 0) {
      out[r] = res;
      stats_add(bs, res);
    }
    c.sync();
  }
  __syncthreads();
  if (stats) stats_flush(bs, stats);
}

cudaError_t launch_align(const DevIndex& I, const KParams& P, const uint8_t* bases, const uint64_t* offs,
                         uint32_t n_reads, int W, bkx_read_result* out, bkx_align_stats* stats, unsigned int* cursor,
                         const HashPool& hp, const uint32_t* ids, const unsigned int* n_ids, int grid, cudaStream_t st) {
  size_t smem = align_smem_bytes(W);
  cudaError_t e = P.best ? ensure_smem(align_reads_kernel<true>, smem, g_smem_general_best)
                         : ensure_smem(align_reads_kernel<false>, smem, g_smem_general);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  if (P.best)
    align_reads_kernel<true><<<grid, kBlockThreads, smem, st>>>(I, P, bases, offs, n_reads, W, out, stats, cursor, hp, ids, n_ids);
  else
    align_reads_kernel<false><<<grid, kBlockThreads, smem, st>>>(I, P, bases, offs, n_reads, W, out, stats, cursor, hp, ids, n_ids);
  return cudaGetLastError();
}

int align_blocks_per_sm(int W) {
  size_t smem = align_smem_bytes(W);
  ensure_smem(align_reads_kernel<false>, smem, g_smem_general);
  ensure_smem(align_reads_kernel<true>, smem, g_smem_general_best);
  int nb = 0, nb2 = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, align_reads_kernel<false>, kBlockThreads, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, align_reads_kernel<true>, kBlockThreads, smem);
  return nb < nb2 ? nb : nb2;
}

// ------------------------------------------------------------------------------------------------
// Fast path: one lane per read (see bkx_fast.cuh)
// ------------------------------------------------------------------------------------------------
// Lane-private overflow set of the fast kernel: kFastHashSlots (epoch << 32 | key) slots per lane in HBM, open
// addressing, emptied by moving to a fresh epoch.  Kept out of line: only reads with many candidate loci get here.
__device__ __forceinline__ uint64_t* fh_table(uint64_t* lane_hash) {
  return lane_hash + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kFastHashSlots;
}
__device__ __noinline__ bool fh_test_insert(uint64_t* lane_hash, uint32_t epoch, uint32_t key) {
  uint64_t* t = fh_table(lane_hash);
  const uint64_t want = ((uint64_t)epoch << 32) | key;
  uint32_t h = (key * 2654435761u) & (kFastHashSlots - 1);
  for (;; h = (h + 1) & (kFastHashSlots - 1)) {
    uint64_t v = t[h];
    if (v == want) return true;
    if ((uint32_t)(v >> 32) != epoch) break;
  }
  t[h] = want;
  return false;
}
__device__ __noinline__ void fh_spill(uint64_t* lane_hash, uint32_t epoch, const uint32_t* seen, int n, uint32_t key) {
  for (int i = 0; i < n; ++i) fh_test_insert(lane_hash, epoch, seen[i * 32]);
  fh_test_insert(lane_hash, epoch, key);
}

// MLX: any multi-loci option (-r1 / -X) is on.  A template parameter, not a run-time test: letting the result code vary
// at run time inside finish() costs the default instantiation its spill-free 64-register allocation (-3 % reads/s).
// SCAN: the step is split in two.  Most cores of a read are absent from the genome and die at the prefix-table lookup (an
// empty bucket); with one core per step those cheap lanes sat idle while a few lanes of the warp searched and walked a
// bucket (11 of 32 lanes active on average, profiles/r01_final_align_fast_ncu.md).  Now every lane first runs down its
// cores -- two table lookups in flight at a time -- until it stands on a non-empty bucket (or its read is finished), and
// only then the warp enters the search / walk / Hamming part, with most lanes having work there.  Skipping over empty
// buckets changes nothing the reference can observe: they yield no candidate, only the seed count moves on.
template <bool MLX, bool SCAN>
__global__ void __launch_bounds__(kFastThreads, 4) align_fast_kernel(
    DevIndex I, KParams P, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offs, uint32_t n_reads,
    int W, bkx_read_result* __restrict__ out, bkx_align_stats* __restrict__ stats, unsigned int* __restrict__ cursor,
    uint32_t* __restrict__ hard_ids, unsigned int* __restrict__ n_hard, uint64_t* __restrict__ lane_hash,
    uint32_t epoch_base, Packed2Src p2, const uint32_t* __restrict__ ids, const unsigned int* __restrict__ n_ids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BlockStats bs;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  for (int i = threadIdx.x; i < (int)(sizeof(BlockStats) / 4); i += blockDim.x) ((unsigned int*)&bs)[i] = 0;
  __syncthreads();

  uint64_t* region = (uint64_t*)smem_raw + (size_t)wib * ((size_t)2 * W * 32 + (size_t)(kFastSeen + 1) * 16);
  uint64_t* wr0 = region + lane;                   // this lane's forward words (stride 32)
  uint64_t* wr1 = region + (size_t)W * 32 + lane;  // reverse-complement words
  FastLane f;
  f.w2[0] = wr0;
  f.w2[1] = wr1;
  f.seen = (uint32_t*)(region + (size_t)2 * W * 32) + lane;
  f.seen[kFastSeen * 32] = epoch_base;  // bumped every time this lane spills a strand's keys into its hash set
  f.L = 0;
  const int k = I.k;
  const int s_first = (P.strand_mode == BKX_STRAND_CRICK) ? 1 : 0;
  const int s_last = (P.strand_mode == BKX_STRAND_WATSON) ? 0 : 1;

  // ---- per-lane state of the read in flight
  bool active = false, exhausted = false;
  uint32_t r = 0;
  int L = 0, max_tot_mm = 0, core_len = 0, core_delta = 0, slides = 0;
  int allow = 0;
  bool in_final = false;
  int mm_max = 0, CL = 0, delta = 0, K = 0, n_cores = 0, last_ofs = 0;
  int s = 0, ci = 0, seen_n = 0;
  int inst = 0, low = 0, nxt = 0, hit_strand = 0, hit_ent = -1, hit_mm = 0;
  uint64_t hit_p = 0;
  uint32_t seeds = 0, cands = 0;

  // returns false if the read has no further phase (=> eHRnone)
  auto setup_phase = [&]() -> bool {
    if (!in_final) {
      bool staged = false;
      if (max_tot_mm > 0 && allow <= max_tot_mm) {
        int cl = small_div(L, allow + P.mmd);
        if (cl > core_len) { staged = true; CL = cl; delta = cl; mm_max = allow; }
      }
      if (!staged) {
        if (max_tot_mm > 0 && allow > max_tot_mm) return false;  // staged loop ran out: no final phase
        in_final = true;
        CL = core_len; delta = core_delta; mm_max = max_tot_mm;
      }
    } else {
      return false;
    }
    K = (L - CL - delta >= 0) ? small_div(L - CL - delta, delta) + 1 : 0;
    int rr = L - (K * delta + CL);
    n_cores = K + 1 + ((rr > CL / 3) ? 1 : 0);
    if (n_cores > slides) n_cores = slides;
    if (L < CL) n_cores = 0;
    last_ofs = L - CL;
    inst = 0;
    low = nxt = mm_max + P.mmd + 1;
    s = s_first;
    ci = 0;
    seen_n = 0;
    return true;
  };

  auto finish = [&](int hr) {  // same mapping as make_result (bkx_align.cuh), kept inline: a call here costs 12 % of the kernel
    bkx_read_result res;
    res.nar = BKX_NAR_NOHIT; res.strand = 0; res.num_hits = 0; res.low_mm = 0;
    res.nxt_low_mm = 0; res.low_hit_instances = 0; res.chrom_id = 0; res.match_loci = 0; res.match_len = 0;
    res.mismatches = 0; res.flags = 0; res.seeds = seeds; res.cands = cands; res.reserved = 0;
    int ii = inst > P.max_hits ? P.max_hits + 1 : inst;
    bool multi = false;
    int nh = 1;
    if constexpr (MLX) {
      if (P.clamp_ml && hr == BKX_HR_HITINSTS) { ii = P.max_hits; hr = BKX_HR_HITS; }
      multi = hr == BKX_HR_HITS && ii > 1 && P.ml_mode != BKX_ML_ALL && P.ml_mode != BKX_ML_DEFAULT;
      if (P.ml_mode == BKX_ML_ALL) nh = ii;
    }
    res.hit_rslt = (uint8_t)hr;
    if (hr == BKX_HR_HITS && !multi) {
      res.nar = BKX_NAR_ACCEPTED;
      res.num_hits = (uint8_t)min(nh, 255);   // -r5 beyond 255 loci: low_hit_instances holds the count
      res.strand = hit_strand ? '-' : '+';
      res.chrom_id = __ldg(I.ent_id + hit_ent);
      res.match_loci = (uint32_t)(hit_p - __ldg(I.ent_start + hit_ent));
      res.match_len = (uint16_t)L;
      res.mismatches = (uint8_t)hit_mm;
      res.low_hit_instances = (int16_t)nh;
      res.low_mm = (int8_t)low;
      res.nxt_low_mm = (int8_t)nxt;
    } else if (hr == BKX_HR_MMDELTA || hr == BKX_HR_HITINSTS || hr == BKX_HR_HITS) {
      res.nar = (hr == BKX_HR_MMDELTA) ? BKX_NAR_MMDELTA : BKX_NAR_MULTIALIGN;
      if (!multi) { res.strand = '?'; res.match_len = (uint16_t)L; }
      res.low_hit_instances = (int16_t)ii;
      res.low_mm = (int8_t)low;
      res.nxt_low_mm = (int8_t)nxt;
      if (multi) { atomicAdd(&bs.multi, 1u); atomicAdd(&bs.multi_loci, (unsigned int)ii); }
    }
    if constexpr (MLX) {
      if (res.nar == BKX_NAR_ACCEPTED && nh > 1) { atomicAdd(&bs.acc_multi, 1u); atomicAdd(&bs.multi_loci, (unsigned int)nh); }
    }
    out[r] = res;
    stats_add_basic(bs, res);
    active = false;
  };

  auto defer = [&]() {
    unsigned int at = atomicAdd(n_hard, 1u);
    hard_ids[at] = r;
    active = false;
  };

  // the end of a phase (both strands done, or the global early exit): next phase, or the read's result
  auto phase_end = [&]() {
    if (inst == 0) {
      bool more = true;
      for (;;) {
        if (in_final) { more = false; break; }
        ++allow;
        if (!setup_phase()) { more = false; break; }
        if (n_cores > 0) break;
      }
      if (!more) { inst = 0; low = 0; nxt = 0; finish(BKX_HR_NONE); }
    } else if ((nxt - low) < P.mmd) {
      finish(BKX_HR_MMDELTA);
    } else if (inst > P.max_hits) {
      finish(BKX_HR_HITINSTS);
    } else {
      finish(BKX_HR_HITS);
    }
  };
  // SCAN state: `have` -- the lane stands on core (s, ci) whose bucket [blo, bhi) is not empty; `pend_end` -- its phase
  // is over and phase_end() is due
  bool have = false, pend_end = false;
  uint64_t blo = 0, bhi = 0;
  auto bucket_of = [&](int strand, int ofs, uint64_t& lo, uint64_t& hi) {
    const uint64_t key = rev2(fl_word(f, strand, ofs)) >> (64 - 2 * k);
    if (CL >= k) {
      lo = pt_get(I, key);
      hi = pt_get(I, key + 1);
    } else {
      const int sh = 2 * (k - CL);
      const uint64_t pfx = key >> sh;
      lo = pt_get(I, pfx << sh);
      hi = pt_get(I, (pfx + 1) << sh);
    }
  };

  for (;;) {
    // ---- (1) idle lanes claim the next reads (one atomic per warp)
    const bool want = !active && !exhausted;
    const unsigned need = __ballot_sync(0xffffffffu, want);
    if (need) {
      const int leader = __ffs(need) - 1;
      unsigned int base_r = 0;
      if (lane == leader) base_r = atomicAdd(cursor, (unsigned int)__popc(need));
      base_r = __shfl_sync(0xffffffffu, base_r, leader);
      bool fresh = false;   // this lane just received a read that fits the fast path
      int nN = 0;
      bool bad = false;
      if (want) {
        r = base_r + __popc(need & lt);
        if (r >= (ids ? __ldg(n_ids) : n_reads)) {   // ids: only the reads the wave path handed on (bkx_wave.cuh)
          exhausted = true;
        } else {
          if (ids) r = __ldg(ids + r);
          active = true;
          seeds = cands = 0;
          L = (int)(__ldg(offs + r + 1) - __ldg(offs + r));
          if (L > kFastMaxLen || ((L + 31) >> 5) + 1 > W) defer();
          else fresh = true;
        }
      }
      // ---- reads that arrive 2-bit packed (and hold nothing but A C G T) are taken as they are: a few unaligned 64-bit
      //      extracts per lane instead of the cooperative packing below (which was 15 % of the kernel's instructions)
      const bool direct = fresh && p2.words != nullptr && !(p2.flags && __ldg(p2.flags + r));
      if (direct) {
        const uint64_t bo = __ldg(offs + r) + p2.phase;
        const uint64_t* src = p2.words + (bo >> 5);
        const unsigned sh = (unsigned)(bo & 31) * 2;
        const int words = (L + 31) >> 5;
        uint64_t prev = __ldg(src);
        for (int w = 0; w < words; ++w) {
          const uint64_t next = __ldg(src + w + 1);
          uint64_t v = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
          prev = next;
          const int len = L - 32 * w;
          if (len < 32) v &= (1ull << (2 * len)) - 1;
          wr0[w * 32] = v;
        }
        wr0[words * 32] = 0;
      }
      // ---- the warp packs the other new reads together: each lane takes 4 bases (one coalesced 32-bit load),
      //      a multiply gathers their 2-bit codes into a byte, three shuffles assemble the 64-bit words
      unsigned pk = __ballot_sync(0xffffffffu, fresh && !direct);
      while (pk) {
        const int j = __ffs(pk) - 1;
        pk &= pk - 1;
        const uint32_t rj = __shfl_sync(0xffffffffu, r, j);
        const int Lj = __shfl_sync(0xffffffffu, L, j);
        const uintptr_t a0 = (uintptr_t)(bases + __ldg(offs + rj));
        const uint32_t* ap = (const uint32_t*)(a0 & ~(uintptr_t)3);
        const unsigned bsh = (unsigned)(a0 & 3) * 8;
        const int ngroups = (Lj + 3) >> 2;
        int cntN = 0;
        bool badl = false;
        for (int g0 = 0; g0 < ngroups; g0 += 32) {
          const int g = g0 + lane;
          uint32_t x = 0;
          if (g < ngroups) {
            x = __funnelshift_r(__ldg(ap + g), __ldg(ap + g + 1), bsh) & 0x07070707u;
            const int rem = Lj - 4 * g;
            if (rem < 4) x &= (1u << (8 * rem)) - 1u;
          }
          const uint32_t nf = x & 0x04040404u;               // N (4) or an invalid code (5..7)
          cntN += __popc(nf);
          badl |= (((nf >> 2) & (x | (x >> 1))) & 0x01010101u) != 0;
          const uint32_t codes = (((x & 0x03030303u) * 0x00041041u) >> 18) & 0xffu;  // 4 bases -> 8 bits
          uint64_t v = (uint64_t)codes << (8 * (lane & 7));
          v |= __shfl_xor_sync(0xffffffffu, v, 1);
          v |= __shfl_xor_sync(0xffffffffu, v, 2);
          v |= __shfl_xor_sync(0xffffffffu, v, 4);
          const int w = (g0 >> 3) + (lane >> 3);
          if ((lane & 7) == 0 && w < W) region[w * 32 + j] = v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) cntN += __shfl_xor_sync(0xffffffffu, cntN, o);
        const unsigned badm = __ballot_sync(0xffffffffu, badl);
        const int wordsj = (Lj + 31) >> 5;
        if (lane == 0) region[wordsj * 32 + j] = 0;   // zero pad word read by the unaligned extracts
        if (lane == j) { nN = cntN; bad = badm != 0; }
      }
      __syncwarp();
      if (fresh) {
        const int words = (L + 31) >> 5;
        int max_ns_seq = 0;
        if (P.max_ns) max_ns_seq = max((L * P.max_ns) / 100, P.max_ns);
        if (bad || nN > max_ns_seq || L < 1) {
          bkx_read_result res;
          res.nar = BKX_NAR_NS; res.hit_rslt = 0; res.strand = 0; res.num_hits = 0; res.low_mm = 0; res.nxt_low_mm = 0;
          res.low_hit_instances = 0; res.chrom_id = 0; res.match_loci = 0; res.match_len = 0; res.mismatches = 0;
          res.flags = 0; res.seeds = 0; res.cands = 0; res.reserved = 0;
          out[r] = res;
          stats_add_basic(bs, res);
          active = false;
        } else if (nN > 0) {
          defer();  // N-bearing reads need the symbol-wise compare of the general kernel
        } else {
          // reverse complement from the packed forward words
          f.L = L;
          for (int w = 0; w < words; ++w) {
            int t = L - 32 * (w + 1);  // forward position of the last base of this rc word
            uint64_t fw;
            if (t >= 0) fw = fl_word(f, 0, t);
            else fw = wr0[0] << (2 * (-t));
            uint64_t rc = rev2(~fw);
            int len = min(32, L - 32 * w);
            if (len < 32) rc &= (1ull << (2 * len)) - 1;
            wr1[w * 32] = rc;
          }
          wr1[words * 32] = 0;
          // per-read search parameters, Aligner.cpp:9085-9095
          max_tot_mm = P.max_subs == 0 ? 0 : max(1, (L * P.max_subs + 50) / 100);
          if (max_tot_mm > 63) max_tot_mm = 63;
          core_len = max(P.min_core_len, small_div(L, P.mmd == 1 ? max_tot_mm + 1 : max_tot_mm + 2));
          slides = max(1, (P.slides_per100 * L + 99) / 100);
          core_delta = max(small_div(L, slides) - 1, core_len);
          allow = 0;
          in_final = false;
          hit_ent = -1;
          hit_p = 0;
          inst = low = nxt = 0;
          while (active) {  // first phase with at least one core
            if (!setup_phase()) { inst = 0; low = 0; nxt = 0; finish(BKX_HR_NONE); break; }
            if (n_cores > 0) break;
            if (in_final) { inst = 0; low = 0; nxt = 0; finish(BKX_HR_NONE); break; }
            ++allow;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, active)) {
      if (__all_sync(0xffffffffu, exhausted)) break;
      continue;
    }
    if constexpr (SCAN) {
      // ---- (2a) every lane runs down its cores until it stands on a non-empty bucket
      const int scan_iters = P.scan_iters;
      for (int it = 0; it < scan_iters; ++it) {
        if (active && pend_end) { pend_end = false; phase_end(); }
        if (active && !have) {
          // this core and the one after it (same phase): both lookups are issued before either is looked at
          const int c0 = ci <= K ? ci * delta : last_ofs;
          int s1 = s, ci1 = ci + 1;
          bool has1 = true;
          if (ci1 >= n_cores) {
            if (s < s_last) { s1 = s + 1; ci1 = 0; } else { has1 = false; s1 = s; ci1 = ci; }
          }
          const int c1 = ci1 <= K ? ci1 * delta : last_ofs;
          uint64_t lo0, hi0, lo1, hi1;
          bucket_of(s, c0, lo0, hi0);
          bucket_of(s1, c1, lo1, hi1);
          ++seeds;
          if (lo0 < hi0) { have = true; blo = lo0; bhi = hi0; }
          else if (!has1) pend_end = true;
          else {
            if (s1 != s) seen_n = 0;
            s = s1; ci = ci1;
            ++seeds;
            if (lo1 < hi1) { have = true; blo = lo1; bhi = hi1; }
            else if (ci + 1 < n_cores) ++ci;
            else if (s < s_last) { ++s; ci = 0; seen_n = 0; }
            else pend_end = true;
          }
        }
        if (!__any_sync(0xffffffffu, active && !have)) break;
      }
    }
    // ---- (2) one core of the current strand/phase for every active lane (SCAN: for the lanes standing on a bucket)
    if (SCAN ? (active && have) : active) {
      const int cofs = ci <= K ? ci * delta : last_ofs;
      bool dfr = false, stop_all = false;
      // SA interval of the core: prefix-table bucket, then lower / upper bound
      if constexpr (!SCAN) {
        ++seeds;
        if (P.prefetch) {
          // the next core of this phase is known now: ask L2 for its table entry, so that the next step's lookup -- 14 of the
          // ~30 DRAM line fetches of a read are such lookups -- finds it there
          int s1 = s, ci1 = ci + 1;
          if (ci1 >= n_cores) { s1 = s + 1; ci1 = 0; }
          if (s1 <= s_last) {
            const int c1 = ci1 <= K ? ci1 * delta : last_ofs;
            uint64_t key1 = rev2(fl_word(f, s1, c1)) >> (64 - 2 * k);
            if (CL < k) { const int sh1 = 2 * (k - CL); key1 = (key1 >> sh1) << sh1; }
            const void* pa = (const void*)(I.pt32 + key1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
          }
        }
        bucket_of(s, cofs, blo, bhi);
      }
      have = false;
      uint64_t first = 0, cnt = 0;
      bool located = false;
      if (blo < bhi && CL <= k) {
        // core no longer than the table key: the bucket IS the interval, except for suffixes holding an N/EOS inside
        // the core span, which sort at the bucket's end -- so if the last element matches, every element does
        uint64_t g = sa_get(I, bhi - 1);
        if (!span_has_exc(I, g, (uint32_t)CL) && fl_cmp(I, f, s, cofs, CL, g) == 0) {
          first = blo;
          cnt = bhi - blo;
          located = true;
          if (cnt > (uint64_t)kFastMaxCnt) dfr = true;
        }
      }
      if (blo < bhi && !located) {
        uint64_t l = blo, h = bhi;
        bool h_equal = false;
        while (l < h) {
          uint64_t m = l + ((h - l) >> 1);
          uint64_t g = sa_get(I, m);
          if (span_has_exc(I, g, (uint32_t)CL)) { dfr = true; break; }
          int c = fl_cmp(I, f, s, cofs, CL, g);
          if (c > 0) l = m + 1; else { h = m; h_equal = (c == 0); }
        }
        if (!dfr && l < bhi && h_equal) {
          first = l;
          // upper bound by stepping: intervals kept in the fast path are short
          uint64_t ul = l + 1;
          while (ul < bhi) {
            if (ul - first >= (uint64_t)kFastMaxCnt) { dfr = true; break; }
            uint64_t g = sa_get(I, ul);
            if (span_has_exc(I, g, (uint32_t)CL)) { dfr = true; break; }
            if (fl_cmp(I, f, s, cofs, CL, g) != 0) break;
            ++ul;
          }
          cnt = ul - first;
        }
      }
      // interval walk, strictly sequential (SfxArrayV2.cpp:5857-6209)
      for (uint64_t e = 0; e < cnt && !dfr; ++e) {
        uint64_t loci = sa_get(I, first + e);
        if (loci < (uint64_t)cofs) continue;
        uint64_t p = loci - (uint64_t)cofs;
        int ent = find_entry(I, p);
        if (ent < 0 || (p + (uint64_t)L - 1) > __ldg(I.ent_end + ent)) continue;
        uint32_t kk = (uint32_t)(1u + (uint32_t)loci - (uint32_t)cofs);
        if (P.xdedup) {
          // "Already processed" without a set: this placement was reached before in this strand / phase exactly when an
          // earlier core of the read also matches the genome there (its interval then held this locus, and the fast path
          // walks every interval it keeps in full).  Tested on the packed words of a window the Hamming loop is about to
          // read anyway -- no key list to scan, no hash set in HBM to probe.  (Indexes of more than 2^32 symbols keep the
          // key set: the reference's 32-bit keys collide there, and that is reproduced.)
          if (span_has_exc(I, p, (uint32_t)L)) { dfr = true; break; }
          bool dup = false;
          for (int j = 0; j < ci && !dup; ++j) {
            const int oj = j <= K ? j * delta : last_ofs;
            dup = fl_cmp(I, f, s, oj, CL, p + (uint64_t)oj) == 0;
          }
          if (dup) continue;
        } else if (seen_n <= kFastSeen) {  // beyond that the strand's keys live in the lane's hash set
          bool dup = false;
          for (int i = 0; i < seen_n; ++i) dup |= (f.seen[i * 32] == kk);
          if (dup) continue;
          if (seen_n < kFastSeen) {
            f.seen[seen_n * 32] = kk;
          } else {  // spill this strand's keys into the lane's hash set in HBM
            const uint32_t epoch = ++f.seen[kFastSeen * 32];
            fh_spill(lane_hash, epoch, f.seen, seen_n, kk);
          }
        } else {
          if (fh_test_insert(lane_hash, f.seen[kFastSeen * 32], kk)) continue;
          if (seen_n >= kFastHashCap) { dfr = true; break; }
        }
        ++seen_n;
        ++cands;
        if (!P.xdedup && span_has_exc(I, p, (uint32_t)L)) { dfr = true; break; }
        // Hamming over packed words; rejected once > MaxTotMM or >= NxtLowMMCnt (SfxArrayV2.cpp:6148-6151)
        uint64_t w = p >> 5;
        unsigned sh = (unsigned)(p & 31) * 2;
        uint64_t prev = __ldg(I.g2 + w);
        int mm = 0;
        const int lim = min(mm_max, nxt - 1);
        for (int b = 0, wi = 0; b < L; b += 32, ++wi) {
          uint64_t next = __ldg(I.g2 + (++w));
          uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
          prev = next;
          uint64_t x = f.w2[s][wi * 32] ^ gw;
          uint64_t m = (x | (x >> 1)) & 0x5555555555555555ull;
          int rem = L - b;
          if (rem < 32) m &= (1ull << (2 * rem)) - 1;
          mm += __popcll(m);
          if (mm > lim) break;
        }
        if (mm > lim) continue;
        if (mm < low) {
          inst = 1; nxt = low; low = mm;
          hit_p = p; hit_ent = ent; hit_mm = mm; hit_strand = s;
        } else if (mm == low) {
          ++inst;
        } else {
          nxt = mm;  // low < mm < nxt
        }
        if constexpr (MLX) {  // -r5: keep every locus at the lowest mismatch count, in discovery order
          if (P.multi && mm == low && inst <= P.max_hits) {
            bkx_multi_hit h;
            h.chrom_id = __ldg(I.ent_id + ent);
            h.match_loci = (uint32_t)(p - __ldg(I.ent_start + ent));
            h.match_len = (uint16_t)L;
            h.strand = s ? '-' : '+';
            h.mismatches = (uint8_t)mm;
            P.multi[(size_t)r * P.max_hits + (inst - 1)] = h;
          }
        }
        if (inst > P.max_hits && low == 0) { stop_all = true; break; }
      }
      if (dfr) {
        defer();
        pend_end = false;
      } else {
        // advance: next core / strand / phase
        ++ci;
        bool phase_done = stop_all;
        if (!phase_done && ci >= n_cores) {
          if (s < s_last) { ++s; ci = 0; seen_n = 0; }
          else phase_done = true;
        }
        if (phase_done) {
          if constexpr (SCAN) pend_end = true;   // handled at the top of the next scan step
          else phase_end();
        }
      }
    }
  }
  __syncthreads();
  if (stats) stats_flush(bs, stats);
}

cudaError_t launch_align_fast(const DevIndex& I, const KParams& P, const uint8_t* bases, const uint64_t* offs,
                              uint32_t n_reads, int W, bkx_read_result* out, bkx_align_stats* stats,
                              unsigned int* cursor, uint32_t* hard_ids, unsigned int* n_hard, uint64_t* lane_hash,
                              uint32_t epoch_base, int grid, cudaStream_t st, const Packed2Src& p2, const uint32_t* ids,
                              const unsigned int* n_ids) {
  size_t smem = fast_smem_bytes(W);
  const bool mlx = P.ml_mode != 0 || P.clamp_ml != 0;
  const bool scan = P.scan_iters > 0;
  cudaError_t e = mlx ? (scan ? ensure_smem(align_fast_kernel<true, true>, smem, g_smem_fast_mlx_scan)
                              : ensure_smem(align_fast_kernel<true, false>, smem, g_smem_fast_mlx))
                      : (scan ? ensure_smem(align_fast_kernel<false, true>, smem, g_smem_fast_scan)
                              : ensure_smem(align_fast_kernel<false, false>, smem, g_smem_fast));
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(n_hard, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
#define BKX_LAUNCH_FAST(M, S)                                                                                              \
  align_fast_kernel<M, S><<<grid, kFastThreads, smem, st>>>(I, P, bases, offs, n_reads, W, out, stats, cursor, hard_ids, \
                                                            n_hard, lane_hash, epoch_base, p2, ids, n_ids)
  if (mlx) { if (scan) BKX_LAUNCH_FAST(true, true); else BKX_LAUNCH_FAST(true, false); }
  else { if (scan) BKX_LAUNCH_FAST(false, true); else BKX_LAUNCH_FAST(false, false); }
#undef BKX_LAUNCH_FAST
  return cudaGetLastError();
}

int fast_blocks_per_sm(int W) {
  size_t smem = fast_smem_bytes(W);
  ensure_smem(align_fast_kernel<false, false>, smem, g_smem_fast);
  ensure_smem(align_fast_kernel<true, false>, smem, g_smem_fast_mlx);
  ensure_smem(align_fast_kernel<false, true>, smem, g_smem_fast_scan);
  ensure_smem(align_fast_kernel<true, true>, smem, g_smem_fast_mlx_scan);
  int nb[4] = {0, 0, 0, 0};
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb[0], align_fast_kernel<false, false>, kFastThreads, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb[1], align_fast_kernel<true, false>, kFastThreads, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb[2], align_fast_kernel<false, true>, kFastThreads, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb[3], align_fast_kernel<true, true>, kFastThreads, smem);
  return std::min(std::min(nb[0], nb[1]), std::min(nb[2], nb[3]));
}

// ------------------------------------------------------------------------------------------------
// Wave path: the default search laid out by kind of work instead of by read (see bkx_wave.cuh)
// ------------------------------------------------------------------------------------------------
constexpr int kWaveThreads = 256;

__device__ __forceinline__ void wave_none_result(bkx_read_result& res, uint32_t seeds, uint32_t cands) {
  res.nar = BKX_NAR_NOHIT; res.hit_rslt = BKX_HR_NONE; res.strand = 0; res.num_hits = 0; res.low_mm = 0; res.nxt_low_mm = 0;
  res.low_hit_instances = 0; res.chrom_id = 0; res.match_loci = 0; res.match_len = 0; res.mismatches = 0;
  res.flags = 0; res.seeds = seeds; res.cands = cands; res.reserved = 0;
}

// The read of this lane, both strands, into its column of the block's shared memory (word w of strand s at wr[s][w * 32],
// as in align_fast_kernel): every later extract is a conflict-free shared-memory access instead of a scattered global
// load -- 32 lanes x 2-4 loads per extract kept the load/store unit, not DRAM, busy.
__device__ __forceinline__ void wave_stage(const ReadRef& q, uint64_t* wr0, uint64_t* wr1, bool fwd_only) {
  const int L = q.L, words = (L + 31) >> 5;
  const uint64_t* src = q.words + (q.bo >> 5);
  const unsigned sh = (unsigned)(q.bo & 31) * 2;
  uint64_t prev = __ldg(src);
  for (int w = 0; w < words; ++w) {
    const uint64_t next = __ldg(src + w + 1);
    uint64_t v = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
    prev = next;
    const int len = L - 32 * w;
    if (len < 32) v &= (1ull << (2 * len)) - 1;
    wr0[w * 32] = v;
  }
  wr0[words * 32] = 0;
  if (fwd_only) return;
  FastLane f;
  f.w2[0] = wr0; f.w2[1] = wr1; f.seen = nullptr; f.L = L;
  for (int w = 0; w < words; ++w) {
    const int t = L - 32 * (w + 1);   // forward position of the last base of this reverse-complement word
    const uint64_t fw = t >= 0 ? fl_word(f, 0, t) : (wr0[0] << (2 * (-t)));
    uint64_t rc = rev2(~fw);
    const int len = min(32, L - 32 * w);
    if (len < 32) rc &= (1ull << (2 * len)) - 1;
    wr1[w * 32] = rc;
  }
  wr1[words * 32] = 0;
}

// One thread per read, once per round: (a) what the read's phase of the previous round found -- then its result, or its
// next phase (round 0: whether the read can go down this path at all, and its first phase); (b) the prefix-table lookups
// of all cores of the phase it is in now, both strands, and an item per bucket that is not empty.  Items go to the queue
// in chunks of kWaveChunk slots owned by one warp: one global atomic per chunk (a returning atomic per append bounds a
// kernel at 2.9 ns each); the unused tail of a chunk is filled with empty items.
// mode 0: first round; 1: a round in between; 2: after the last round -- (a) only, a read still on the path is handed on
__global__ void __launch_bounds__(kWaveThreads) wave_step_kernel(DevIndex I, KParams P, const uint64_t* __restrict__ offs,
                                                                 uint32_t n_reads, Packed2Src p2, WaveBuf B, int round, int mode,
                                                                 int W, bkx_read_result* __restrict__ out,
                                                                 bkx_align_stats* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char wave_smem[];
  __shared__ BlockStats bs;
  for (int i = threadIdx.x; i < (int)(sizeof(BlockStats) / 4); i += blockDim.x) ((unsigned int*)&bs)[i] = 0;
  __syncthreads();
  FastLane f;   // this lane's column of the block's staging area: W words per strand
  f.w2[0] = (uint64_t*)wave_smem + (size_t)(threadIdx.x >> 5) * ((size_t)2 * W * 32) + (threadIdx.x & 31);
  f.w2[1] = f.w2[0] + (size_t)W * 32;
  f.seen = nullptr;
  f.L = 0;
  const int s_first = (P.strand_mode == BKX_STRAND_CRICK) ? 1 : 0;
  const int s_last = (P.strand_mode == BKX_STRAND_WATSON) ? 0 : 1;
  const int n_strands = s_last - s_first + 1;
  const bool wide = I.n > (1ull << 32);
  const int k = I.k;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned int* q_count = B.cnt + kWaveCntItems + round;
  const bool half = B.half_loads != 0;
  uint64_t cur = 0, end = 0;   // this warp's chunk of the item queue (warp-uniform)
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n_reads; base += stride) {
    const uint32_t r = (uint32_t)(base + lane);
    const bool in_range = base + lane < n_reads;
    unsigned state = kWaveOff;
    ReadRef q;
    q.words = p2.words; q.bo = 0; q.L = 0;
    WaveRead w;
    w.L = 0; w.max_tot_mm = 0; w.core_len = 0; w.slides = 0; w.core_delta = 0;
    bool handed_on = false;
    if (in_range) {
      if (mode == 0) {
        const uint64_t o0 = __ldg(offs + r);
        const uint64_t len = __ldg(offs + r + 1) - o0;
        if (len < 1 || len > (uint64_t)kFastMaxLen || (int)((len + 31) >> 5) + 1 > W || (p2.flags && __ldg(p2.flags + r))) {
          handed_on = true;
        } else {
          q.bo = o0 + p2.phase; q.L = (int)len;
          wave_read(P, q.L, w);
          const int first = wave_enter(P, w, 0, true);
          B.fb[r] = 0;
          B.ncand[r] = 0;
          B.acc[r] = make_uint2(0u, 0u);
          if (first < 0) {
            bkx_read_result res;
            wave_none_result(res, 0, 0);
            out[r] = res;
            stats_add_basic(bs, res);
          } else {
            state = (unsigned)first;
          }
        }
        B.ph[r] = (uint8_t)state;
      } else {
        state = B.ph[r];
        if (state != kWaveOff) {
          // ---- (a) the phase the read was in
          const unsigned old_state = state;
          state = kWaveOff;
          const uint64_t o0 = __ldg(offs + r);
          q.bo = o0 + p2.phase; q.L = (int)(__ldg(offs + r + 1) - o0);
          wave_read(P, q.L, w);
          if (B.fb[r]) {
            handed_on = true;
          } else {
            WavePhase ph;
            wave_phase(P, w, old_state, ph);
            const unsigned nc = B.ncand[r];
            int inst = 0, low = ph.mm_max + P.mmd + 1, nxt = low;
            unsigned distinct = 0;
            uint64_t hit = 0;
            bool redo = false;
            if (nc) {
              const uint64_t* __restrict__ row = B.cand + (size_t)r * B.row;
              const uint64_t kmask = (1ull << 41) - 1;   // placement + strand
              for (unsigned i = 0; i < nc; ++i) {
                const uint64_t c = row[i];
                const uint64_t key = c & kmask;
                bool dup = false;
                for (unsigned j = 0; j < i; ++j) {
                  const uint64_t kj = row[j] & kmask;
                  if (kj == key) dup = true;
                  // the reference's "already processed" keys are 32 bits wide (SfxArrayV2.cpp:6093): beyond 2^32 symbols
                  // two placements of one strand can share a key, and which of them is dropped depends on the order they
                  // are met in
                  else if (wide && ((kj ^ key) >> 40) == 0 && (uint32_t)kj == (uint32_t)key) redo = true;
                }
                if (dup) continue;
                ++distinct;
                const int mm = (int)((c >> 41) & 127);
                if (mm == (int)kWaveFailed) continue;
                if (mm < low) { inst = 1; nxt = low; low = mm; hit = c; }
                else if (mm == low) ++inst;
                else if (mm < nxt) nxt = mm;
              }
              B.ncand[r] = 0;
            }
            // more exact placements than MaxHits: the reference stops the phase there (SfxArrayV2.cpp:6199) and what it
            // has seen by then depends on the order
            if (redo || (inst > P.max_hits && low == 0)) {
              handed_on = true;
            } else {
              const uint2 acc = B.acc[r];
              const uint32_t seeds = acc.x + (uint32_t)(n_strands * ph.n_cores), cands = acc.y + distinct;
              bkx_read_result res;
              wave_none_result(res, seeds, cands);
              bool done = true;
              if (inst == 0) {
                const int next = wave_enter(P, w, (int)old_state, false);
                if (next >= 0) {
                  if (mode == 2) handed_on = true;
                  else { state = (unsigned)next; B.acc[r] = make_uint2(seeds, cands); }
                  done = false;
                }
              } else {
                // ProcCoredApprox's mapping for the default multi-loci mode (finish() of align_fast_kernel)
                const int hr = (nxt - low) < P.mmd ? BKX_HR_MMDELTA : inst > P.max_hits ? BKX_HR_HITINSTS : BKX_HR_HITS;
                res.hit_rslt = (uint8_t)hr;
                res.low_mm = (int8_t)low;
                res.nxt_low_mm = (int8_t)nxt;
                res.match_len = (uint16_t)q.L;
                if (hr == BKX_HR_HITS) {
                  const uint64_t p = hit & 0xffffffffffull;
                  const int ent = find_entry(I, p);
                  res.nar = BKX_NAR_ACCEPTED;
                  res.num_hits = 1;
                  res.strand = ((hit >> 40) & 1) ? '-' : '+';
                  res.chrom_id = __ldg(I.ent_id + ent);
                  res.match_loci = (uint32_t)(p - __ldg(I.ent_start + ent));
                  res.mismatches = (uint8_t)low;
                  res.low_hit_instances = 1;
                } else {
                  res.nar = (hr == BKX_HR_MMDELTA) ? BKX_NAR_MMDELTA : BKX_NAR_MULTIALIGN;
                  res.strand = '?';
                  res.low_hit_instances = (int16_t)(inst > P.max_hits ? P.max_hits + 1 : inst);
                }
              }
              if (done) {
                out[r] = res;
                stats_add_basic(bs, res);
              }
            }
          }
          B.ph[r] = (uint8_t)state;
        }
      }
      if (handed_on) B.fb_ids[wave_push(B.cnt + kWaveCntFallback)] = r;
    }
    if (mode == 2) continue;
    // ---- (b) the lookups of the phase the read is in now
    WavePhase ph;
    ph.n_cores = 0; ph.CL = 0; ph.K = 0; ph.delta = 0; ph.last_ofs = 0; ph.mm_max = 0;
    if (state != kWaveOff) {
      wave_phase(P, w, state, ph);
      wave_stage(q, (uint64_t*)f.w2[0], (uint64_t*)f.w2[1], s_first == s_last && s_first == 0);
      f.L = q.L;
    }
    const int maxc = __reduce_max_sync(0xffffffffu, ph.n_cores);
    for (int s = s_first; s <= s_last; ++s) {
      for (int c0 = 0; c0 < maxc; c0 += 4) {
        uint64_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // four cores' lookups are issued before any of them is looked at
          lo[j] = hi[j] = 0;
          const int ci = c0 + j;
          if (ci < ph.n_cores) {
            const int cofs = ci <= ph.K ? ci * ph.delta : ph.last_ofs;
            const uint64_t key = rev2(fl_word(f, s, cofs)) >> (64 - 2 * k);
            if (ph.CL >= k) {
              lo[j] = half ? pt_get_half(I, key) : pt_get(I, key);
              hi[j] = half ? pt_get_half(I, key + 1) : pt_get(I, key + 1);
            } else {
              const int sh = 2 * (k - ph.CL);
              const uint64_t pfx = key >> sh;
              lo[j] = pt_get(I, pfx << sh);
              hi[j] = pt_get(I, (pfx + 1) << sh);
            }
          }
        }
        int mine = 0;
        bool huge = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (hi[j] - lo[j] >= 0xffffffffull) { huge = true; hi[j] = lo[j]; }
          mine += lo[j] < hi[j];
        }
        if (huge) B.fb[r] = 1;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int all = __shfl_sync(0xffffffffu, incl, 31);
        if (all == 0) continue;
        if (cur + (uint64_t)all > end) {   // the warp's chunk is used up: close it, take the next
          for (uint64_t i = cur + lane; i < end; i += 32) B.items[i] = make_ulonglong4(0, kWaveNoItem, 0, 0);
          unsigned int nb = 0;
          if (lane == 0) nb = atomicAdd(q_count, (unsigned)kWaveChunk);
          nb = __shfl_sync(0xffffffffu, nb, 0);
          cur = nb;
          end = cur + kWaveChunk;
        }
        if (end > B.item_cap) {   // queue full: these reads are redone by align_fast_kernel
          if (mine) B.fb[r] = 1;
          cur = end = 0;
          continue;
        }
        uint64_t at = cur + (uint64_t)(incl - mine);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (lo[j] < hi[j]) {
            B.items[at++] = make_ulonglong4(lo[j] | ((uint64_t)s << 40) | ((uint64_t)(c0 + j) << 41),
                                            (uint64_t)r | ((hi[j] - lo[j]) << 32),
                                            q.bo | ((uint64_t)q.L << 44) | ((uint64_t)state << 56), 0);
          }
        }
        cur += (uint64_t)all;
      }
    }
  }
  for (uint64_t i = cur + lane; i < end; i += 32) B.items[i] = make_ulonglong4(0, kWaveNoItem, 0, 0);
  __syncthreads();
  if (stats) stats_flush(bs, stats);
}

// one thread per item: the suffix-array element the search of its bucket looks at first -- the bucket's last one when
// the bucket is the interval (core no longer than the table key), its middle one otherwise.  Nothing but that gather.
__global__ void __launch_bounds__(kWaveThreads) wave_sa_kernel(DevIndex I, KParams P, WaveBuf B, int round) {
  const uint64_t n_raw = B.cnt[kWaveCntItems + round];
  const uint64_t n_items = n_raw < B.item_cap ? n_raw : (B.item_cap / kWaveChunk) * kWaveChunk;
  const int k = I.k;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_items; t += stride) {
    const ulonglong4 it = B.items[t];
    if (it.y == kWaveNoItem) continue;
    const uint64_t blo = it.x & 0xffffffffffull, size = it.y >> 32;
    WaveRead w;
    wave_read(P, (int)((it.z >> 44) & 0xfff), w);
    WavePhase ph;
    wave_phase(P, w, (unsigned)(it.z >> 56), ph);
    const uint64_t at = ph.CL <= k ? blo + size - 1 : blo + (size >> 1);
    reinterpret_cast<uint64_t*>(B.items + t)[3] = sa_get(I, at);
  }
}

// one thread per item: the core's interval inside its bucket, the walk over it, Hamming of every placement
// (the body of align_fast_kernel's step (2), with the read taken from the 2-bit stream)
template <bool HALF>
__global__ void __launch_bounds__(kWaveThreads) wave_probe_kernel(DevIndex I, KParams P, Packed2Src p2, WaveBuf B, int round,
                                                                  int W) {
  extern __shared__ __align__(16) unsigned char wave_smem[];
  FastLane f;
  f.w2[0] = (uint64_t*)wave_smem + (size_t)(threadIdx.x >> 5) * ((size_t)2 * W * 32) + (threadIdx.x & 31);
  f.w2[1] = f.w2[0] + (size_t)W * 32;
  f.seen = nullptr;
  const uint64_t n_raw = B.cnt[kWaveCntItems + round];
  const uint64_t n_items = n_raw < B.item_cap ? n_raw : (B.item_cap / kWaveChunk) * kWaveChunk;
  const int k = I.k;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_items; t += stride) {
    const ulonglong4 it = B.items[t];
    if (it.y == kWaveNoItem) continue;
    const uint64_t blo = it.x & 0xffffffffffull, bhi = blo + (it.y >> 32);
    const int s = (int)((it.x >> 40) & 1), ci = (int)(it.x >> 41);
    const uint32_t r = (uint32_t)it.y;
    ReadRef q;
    q.words = p2.words; q.bo = it.z & ((1ull << 44) - 1); q.L = (int)((it.z >> 44) & 0xfff);
    const int L = q.L;
    f.L = L;
    wave_stage(q, (uint64_t*)f.w2[0], (uint64_t*)f.w2[1], s == 0);   // only strand s is looked at
    WaveRead w;
    wave_read(P, L, w);
    WavePhase ph;
    wave_phase(P, w, (unsigned)(it.z >> 56), ph);
    const int CL = ph.CL;
    const int cofs = ci <= ph.K ? ci * ph.delta : ph.last_ofs;
    bool dfr = false;
    uint64_t first = 0, cnt = 0;
    bool located = false;
    // the element the search of the bucket looks at first -- SA[at0] = g0 -- comes from wave_sa_kernel when that ran
    const uint64_t at0 = CL <= k ? bhi - 1 : blo + ((bhi - blo) >> 1);
    const uint64_t g0 = B.sa_split ? it.w : wv_sa<HALF>(I, at0);
    if (CL <= k) {
      // core no longer than the table key: the bucket IS the interval, except for suffixes holding an N/EOS inside the
      // core span, which sort at the bucket's end -- so if the last element matches, every element does
      if (!span_has_exc(I, g0, (uint32_t)CL) && wv_cmp<HALF>(I, f, s, cofs, CL, g0) == 0) {
        first = blo;
        cnt = bhi - blo;
        located = true;
        if (cnt > (uint64_t)kFastMaxCnt) dfr = true;
      }
    }
    if (!located) {
      uint64_t l = blo, h = bhi;
      bool h_equal = false;
      while (l < h) {
        const uint64_t m = l + ((h - l) >> 1);
        const uint64_t g = m == at0 ? g0 : wv_sa<HALF>(I, m);
        if (span_has_exc(I, g, (uint32_t)CL)) { dfr = true; break; }
        const int c = wv_cmp<HALF>(I, f, s, cofs, CL, g);
        if (c > 0) l = m + 1; else { h = m; h_equal = (c == 0); }
      }
      if (!dfr && l < bhi && h_equal) {
        first = l;
        uint64_t ul = l + 1;
        while (ul < bhi) {
          if (ul - first >= (uint64_t)kFastMaxCnt) { dfr = true; break; }
          const uint64_t g = wv_sa<HALF>(I, ul);
          if (span_has_exc(I, g, (uint32_t)CL)) { dfr = true; break; }
          if (wv_cmp<HALF>(I, f, s, cofs, CL, g) != 0) break;
          ++ul;
        }
        cnt = ul - first;
      }
    }
    for (uint64_t e = 0; e < cnt && !dfr; ++e) {
      const uint64_t loci = first + e == at0 ? g0 : wv_sa<HALF>(I, first + e);
      if (loci < (uint64_t)cofs) continue;
      const uint64_t p = loci - (uint64_t)cofs;
      const int ent = find_entry(I, p);
      if (ent < 0 || (p + (uint64_t)L - 1) > __ldg(I.ent_end + ent)) continue;
      const unsigned slot = atomicAdd(B.ncand + r, 1u);
      if (slot >= (unsigned)B.row || span_has_exc(I, p, (uint32_t)L)) { dfr = true; break; }
      // Hamming over packed words, given up once beyond the phase's allowance
      uint64_t gw_i = p >> 5;
      const unsigned sh = (unsigned)(p & 31) * 2;
      uint64_t prev = wv_g2<HALF>(I, gw_i);
      int mm = 0;
      for (int b = 0, wi = 0; b < L; b += 32, ++wi) {
        const uint64_t next = wv_g2<HALF>(I, ++gw_i);
        const uint64_t gw = sh ? ((prev >> sh) | (next << (64 - sh))) : prev;
        prev = next;
        const uint64_t x = f.w2[s][wi * 32] ^ gw;
        uint64_t m = (x | (x >> 1)) & 0x5555555555555555ull;
        const int rem = L - b;
        if (rem < 32) m &= (1ull << (2 * rem)) - 1;   // the read's last word is zero beyond its end: mask the genome side
        mm += __popcll(m);
        if (mm > ph.mm_max) break;
      }
      const uint64_t mmc = mm > ph.mm_max ? (uint64_t)kWaveFailed : (uint64_t)mm;
      B.cand[(size_t)r * B.row + slot] = p | ((uint64_t)s << 40) | (mmc << 41);
    }
    if (dfr) B.fb[r] = 1;
  }
}

// rounds of a launch: a read goes through at most one phase per allowance and the final one
static int wave_rounds(const KParams& P, uint32_t max_len) {
  const uint32_t len = max_len < (uint32_t)kFastMaxLen ? max_len : (uint32_t)kFastMaxLen;
  int mt = P.max_subs == 0 ? 0 : std::max(1, (int)((len * (uint32_t)P.max_subs + 50) / 100));
  if (mt > 63) mt = 63;
  // staged phases exist for allowances below MaxTotMM only (at MaxTotMM the staged core is the final core), then the final
  // one; whatever is still on the path after the last round is handed on by the closing step
  return std::min(mt + 1, kWaveMaxRounds - 1);
}
// kernels one launch_wave starts: a step and a probe kernel per round (and the suffix-array gather when that is split off),
// and the closing step
int wave_launches(const KParams& P, uint32_t max_len, bool sa_split) { return 1 + (sa_split ? 3 : 2) * wave_rounds(P, max_len); }

cudaError_t launch_wave(const DevIndex& I, const KParams& P, const uint64_t* offs, uint32_t n_reads, uint32_t max_len,
                        const Packed2Src& p2, const WaveBuf& B, bkx_read_result* out, bkx_align_stats* stats, int sms,
                        cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(B.cnt, 0, kWaveCounters * sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  const int grid = sms * (2048 / kWaveThreads);
  const int rounds = wave_rounds(P, max_len);
  // staging area: two strands x W words per thread (reads of up to kFastMaxLen bases: 11 words, 45 KB per block)
  const int W = (int)((std::min<uint32_t>(std::max<uint32_t>(max_len, 1), (uint32_t)kFastMaxLen) + 31) / 32) + 1;
  const size_t smem = (size_t)2 * W * 8 * kWaveThreads;
  static const bool trace = getenv("BKX_TRACE") != nullptr;   // diagnostic: time of every kernel of every round, serialising
  std::vector<cudaEvent_t> ev;
  auto mark = [&]() { if (trace) { cudaEvent_t e2; cudaEventCreate(&e2); cudaEventRecord(e2, st); ev.push_back(e2); } };
  mark();
  for (int round = 0; round < rounds; ++round) {
    wave_step_kernel<<<grid, kWaveThreads, smem, st>>>(I, P, offs, n_reads, p2, B, round, round == 0 ? 0 : 1, W, out, stats);
    mark();
    if (B.sa_split) wave_sa_kernel<<<grid, kWaveThreads, 0, st>>>(I, P, B, round);
    mark();
    if (B.half_loads) wave_probe_kernel<true><<<grid, kWaveThreads, smem, st>>>(I, P, p2, B, round, W);
    else wave_probe_kernel<false><<<grid, kWaveThreads, smem, st>>>(I, P, p2, B, round, W);
    mark();
  }
  wave_step_kernel<<<grid, kWaveThreads, smem, st>>>(I, P, offs, n_reads, p2, B, rounds, 2, W, out, stats);
  if (trace) {
    cudaStreamSynchronize(st);
    std::vector<unsigned int> c(kWaveCounters);
    cudaMemcpy(c.data(), B.cnt, kWaveCounters * 4, cudaMemcpyDeviceToHost);
    for (int round = 0; round < rounds; ++round) {
      float a = 0.f, b = 0.f, d = 0.f;
      cudaEventElapsedTime(&a, ev[3 * round], ev[3 * round + 1]);
      cudaEventElapsedTime(&b, ev[3 * round + 1], ev[3 * round + 2]);
      cudaEventElapsedTime(&d, ev[3 * round + 2], ev[3 * round + 3]);
      fprintf(stderr, "[bkx trace] wave %d: %u item slots; step %.3f ms, sa %.3f ms, probe %.3f ms\n", round,
              c[kWaveCntItems + round], a, b, d);
    }
    for (cudaEvent_t e2 : ev) cudaEventDestroy(e2);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Paired ends: AcceptProvPE / PEInsertSize and the per-pair state machine without orphan recovery
// (Aligner.cpp:2726-2850, 3107-3216, 3421-3477).  One thread per pair.  keep: AcceptThisChromID per chromosome id
// (-Z / -z, Aligner.cpp:2651-2710), NULL when no filter is set.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pe_insert_size(const bkx_pe_params& pe, uint8_t s1, uint32_t st1, uint32_t en1,
                                              uint8_t s2, uint32_t st2, uint32_t en2) {
  if ((pe.pair_strand && s1 != s2) || (!pe.pair_strand && s1 == s2)) return -1;
  int frag;
  if (pe.circularised) frag = (s1 == '+') ? 1 + (int)en1 - (int)st2 : 1 + (int)st2 - (int)en1;
  else frag = (s1 == '+') ? 1 + (int)en2 - (int)st1 : 1 + (int)en1 - (int)st2;
  if (frag < 0) return -1;
  if (frag < pe.pair_min_len) return -6;
  if (frag > pe.pair_max_len) return -7;
  return frag;
}

struct PEBlock { unsigned int v[8]; };

constexpr int kPairHistBins = 4096;

__global__ void pair_reads_kernel(bkx_pe_params pe, bkx_read_result* __restrict__ res, uint32_t n_pairs,
                                  bkx_pe_stats* __restrict__ stats, uint32_t* __restrict__ len_dist,
                                  uint32_t* __restrict__ orphan_list, unsigned int* __restrict__ n_orphans,
                                  const uint8_t* __restrict__ keep) {
  __shared__ PEBlock pb;
  // insert lengths concentrate on a few hundred values: counted per block in shared memory and added to the global
  // histogram once per block (one global atomic per accepted pair on those few addresses was what the kernel waited for)
  __shared__ unsigned int sh_hist[kPairHistBins];
  if (threadIdx.x < 8) pb.v[threadIdx.x] = 0;
  if (len_dist) for (int k = threadIdx.x; k < kPairHistBins; k += blockDim.x) sh_hist[k] = 0;
  __syncthreads();
  enum { UNAL = 0, ACCP = 1, ACCSE = 2, PPAIRED = 3, PUNP = 4, FILT = 5, UNDER = 6, OVER = 7 };
  const int mode = pe.pe_proc;
  unsigned int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // this thread's share of the counters: one shared-memory atomic each at the end
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
    // the two 32-byte records of a pair as four 16-byte loads (and stores below): the kernel streams 128 bytes per pair and
    // was limited by the number of load / store instructions in flight (lg_throttle, profiles/r02_pe_kernels_ncu.md)
    bkx_read_result f, r;
    {
      const uint4* src = reinterpret_cast<const uint4*>(res + 2 * (size_t)i);
      const uint4 a = src[0], b = src[1], c = src[2], d = src[3];
      memcpy(&f, &a, 16); memcpy(reinterpret_cast<char*>(&f) + 16, &b, 16);
      memcpy(&r, &c, 16); memcpy(reinterpret_cast<char*>(&r) + 16, &d, 16);
    }
    auto store_pair = [&]() {
      uint4 a, b, c, d;
      memcpy(&a, &f, 16); memcpy(&b, reinterpret_cast<const char*>(&f) + 16, 16);
      memcpy(&c, &r, 16); memcpy(&d, reinterpret_cast<const char*>(&r) + 16, 16);
      uint4* dst = reinterpret_cast<uint4*>(res + 2 * (size_t)i);
      dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
    };
    f.flags &= (uint8_t)~(BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED);
    r.flags &= (uint8_t)~(BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED);
    bool f_un = f.nar == BKX_NAR_NS || f.nar == BKX_NAR_NOHIT || f.nar == BKX_NAR_UNALIGNED;
    bool r_un = r.nar == BKX_NAR_NS || r.nar == BKX_NAR_NOHIT || r.nar == BKX_NAR_UNALIGNED;
    bool done = false;
    if (!(f.nar == BKX_NAR_ACCEPTED || r.nar == BKX_NAR_ACCEPTED)) {
      ++cnt[UNAL];
      done = true;
    } else if (mode == BKX_PE_UNIQUE && (f_un || r_un)) {
      f.num_hits = r.num_hits = 0;
      f.low_hit_instances = r.low_hit_instances = 0;
      if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PENOHIT;
      if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PENOHIT;
      ++cnt[PUNP];
      done = true;
    } else if (f.nar == BKX_NAR_ACCEPTED && r.nar == BKX_NAR_ACCEPTED) {
      int frag;
      if (!(f.num_hits == 1 && r.num_hits == 1)) frag = 0;
      else {
        const bool bf = keep ? __ldg(keep + f.chrom_id) != 0 : true;
        if (f.chrom_id != r.chrom_id) {  // Aligner.cpp:2771-2786: -3 both ends filtered, -4 the 5' end, -5 the 3' end
          const bool br = keep ? __ldg(keep + r.chrom_id) != 0 : true;
          frag = (bf && br) ? -2 : (!bf && !br) ? -3 : !bf ? -4 : -5;
        } else if (!bf) frag = -3;
        else frag = pe_insert_size(pe, f.strand, f.match_loci, f.match_loci + f.match_len - 1, r.strand, r.match_loci,
                                   r.match_loci + r.match_len - 1);
      }
      if (frag > 0) {
        f.flags |= BKX_FLG_PE_ALIGNED;
        r.flags |= BKX_FLG_PE_ALIGNED;
        if (len_dist) { if (frag < kPairHistBins) atomicAdd(&sh_hist[frag], 1u); else atomicAdd(len_dist + frag, 1u); }
        ++cnt[ACCP];
        done = true;
      } else {
        switch (frag) {
          case -1: f.nar = r.nar = BKX_NAR_PESTRAND; break;
          case -2: f.nar = r.nar = BKX_NAR_PECHROM; break;
          case -6: f.nar = r.nar = BKX_NAR_PEINSERTMIN; break;
          case -7: f.nar = r.nar = BKX_NAR_PEINSERTMAX; break;
          case -3:  // Aligner.cpp:3170
            ++cnt[FILT];
            f.num_hits = r.num_hits = 0;
            f.low_hit_instances = r.low_hit_instances = 0;
            f.nar = r.nar = BKX_NAR_CHROMFILT;
            done = true;
            break;
          case -4: f.nar = BKX_NAR_CHROMFILT; f.low_hit_instances = 0; f.num_hits = 0; break;
          case -5: r.nar = BKX_NAR_CHROMFILT; r.low_hit_instances = 0; r.num_hits = 0; break;
          default: break;
        }
        if (!done && mode == BKX_PE_UNIQUE) {
          f.num_hits = r.num_hits = 0;
          f.low_hit_instances = r.low_hit_instances = 0;
          if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PENOHIT;
          if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PENOHIT;
          ++cnt[PUNP];
          done = true;
        }
      }
    }
    if (!done) {
      ++cnt[PUNP];
      if ((mode == BKX_PE_ORPHAN || mode == BKX_PE_ORPHAN_SE) && orphan_list &&
          ((f.num_hits == 1 && !r_un) || (r.num_hits == 1 && !f_un))) {
        // orphan recovery is a separate (warp per orphan) kernel; it finishes this pair
        orphan_list[atomicAdd(n_orphans, 1u)] = i;
        store_pair();
        continue;
      }
      if (f.nar == BKX_NAR_CHROMFILT || r.nar == BKX_NAR_CHROMFILT) ++cnt[FILT];
      if (f.nar == BKX_NAR_PEINSERTMIN || r.nar == BKX_NAR_PEINSERTMIN) ++cnt[UNDER];
      if (f.nar == BKX_NAR_PEINSERTMAX || r.nar == BKX_NAR_PEINSERTMAX) ++cnt[OVER];
      if (!(mode == BKX_PE_ORPHAN_SE || mode == BKX_PE_UNIQUE_SE)) {
        f.num_hits = r.num_hits = 0;
        f.low_hit_instances = r.low_hit_instances = 0;
        if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PENOHIT;
        if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PENOHIT;
      } else {
        // an end counts when it has its one locus on a chromosome that passes the filter (Aligner.cpp:3442-3477)
        if (f.num_hits != 1 || (keep && !__ldg(keep + f.chrom_id))) {
          f.num_hits = 0; f.low_hit_instances = 0;
          if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PEUNALIGN;
        } else { f.nar = BKX_NAR_ACCEPTED; ++cnt[ACCSE]; }
        if (r.num_hits != 1 || (keep && !__ldg(keep + r.chrom_id))) {
          r.num_hits = 0; r.low_hit_instances = 0;
          if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PEUNALIGN;
        } else { r.nar = BKX_NAR_ACCEPTED; ++cnt[ACCSE]; }
      }
    }
    store_pair();
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (cnt[k]) atomicAdd(&pb.v[k], cnt[k]);
  __syncthreads();
  if (len_dist)
    for (int k = threadIdx.x; k < kPairHistBins; k += blockDim.x)
      if (sh_hist[k]) atomicAdd(len_dist + k, sh_hist[k]);
  if (threadIdx.x == 0 && stats) {
    atomicAdd((unsigned long long*)&stats->unaligned_pairs, (unsigned long long)pb.v[UNAL]);
    atomicAdd((unsigned long long*)&stats->accepted_num_paired, (unsigned long long)pb.v[ACCP]);
    atomicAdd((unsigned long long*)&stats->accepted_num_se, (unsigned long long)pb.v[ACCSE]);
    atomicAdd((unsigned long long*)&stats->partner_paired, (unsigned long long)pb.v[PPAIRED]);
    atomicAdd((unsigned long long*)&stats->partner_unpaired, (unsigned long long)pb.v[PUNP]);
    atomicAdd((unsigned long long*)&stats->num_filtered_by_chrom, (unsigned long long)pb.v[FILT]);
    atomicAdd((unsigned long long*)&stats->under_len_pairs, (unsigned long long)pb.v[UNDER]);
    atomicAdd((unsigned long long*)&stats->over_len_pairs, (unsigned long long)pb.v[OVER]);
  }
}

cudaError_t launch_pair(const bkx_pe_params& pe, bkx_read_result* res, uint32_t n_pairs, bkx_pe_stats* stats,
                        uint32_t* len_dist, uint32_t* orphan_list, unsigned int* n_orphans, const uint8_t* keep,
                        cudaStream_t st) {
  int grid = (int)((n_pairs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  pair_reads_kernel<<<grid, 256, 0, st>>>(pe, res, n_pairs, stats, len_dist, orphan_list, n_orphans, keep);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Orphan-mate recovery (see bkx_rescue.cuh): one warp per orphan pair
// ------------------------------------------------------------------------------------------------
struct RescueCtx {
  uint64_t* rd2;    // oriented mate, 2-bit codes
  uint32_t* rdx;    // oriented mate, N flags (u32 per 32 bases)
  uint64_t* win2;   // staged genome window, 2-bit codes
  uint64_t* winx;   // staged genome window, N/EOS bitmap (64 per word)
  int lane;
};

// mismatch bitmap word i (32 bases) of the oriented mate against the concatenation at p, from global memory
__device__ __forceinline__ uint32_t mm_word_global(const DevIndex& I, const RescueCtx& c, uint64_t p, int i, int L) {
  uint64_t g = p + 32ull * i;
  uint64_t gw = gword(I, g);
  // N/EOS flags of [g, g+32); past the end of the concatenation counts as EOS
  uint64_t xw = g >> 6;
  unsigned xs = (unsigned)(g & 63);
  uint64_t xa = __ldg(I.gx + xw) >> xs;
  if (xs > 32) xa |= __ldg(I.gx + xw + 1) << (64 - xs);
  uint32_t gx32 = (uint32_t)xa;
  if (g + 32 > I.n) gx32 |= (g >= I.n) ? 0xffffffffu : (0xffffffffu << (unsigned)(I.n - g));
  if (g + 32 > I.n) gw |= (g >= I.n) ? ~0ull : (~0ull << (2 * (unsigned)(I.n - g)));  // EOS code past the end
  uint64_t x = c.rd2[i] ^ gw;
  uint32_t m = compress_even(x | (x >> 1)) | (gx32 ^ c.rdx[i]);
  // a flagged genome base against a flagged read base only matches when both are N (code 0): EOS is code 3
  int rem = L - 32 * i;
  if (rem < 32) m &= (1u << rem) - 1;
  return m;
}

// AlignPairedRead (SfxArrayV2.cpp:8247-8433), MinChimericLen 0.  All lanes call; returns 1 and the hit in
// (out_loci, out_mm) when a mate alignment is found.  The oriented mate is already in c.rd2 / c.rdx.
__device__ __forceinline__ int align_paired_read(const DevIndex& I, const KParams& P, const bkx_pe_params& pe,
                                                 RescueCtx& c, bool has_n, bool b3, uint32_t chrom_id, uint32_t start_loci,
                                                 uint32_t end_loci, int L, uint32_t& out_loci, int& out_mm) {
  const int min_d = pe.pair_min_len, max_d = pe.pair_max_len, max_allowed = P.max_subs;
  if (min_d < L || min_d > max_d) return 0;
  if (chrom_id < 1 || chrom_id > I.max_ent_id) return 0;
  const uint32_t ei = __ldg(I.ent_of_id + chrom_id);
  if (ei == 0xffffffffu) return 0;
  const uint64_t cs = __ldg(I.ent_start + ei);
  const uint32_t targ_len = (uint32_t)(__ldg(I.ent_end + ei) - cs + 1);
  int targ_loci;
  if (b3) {
    targ_loci = (int)start_loci;
    if ((uint32_t)(targ_loci + min_d) > targ_len) return 0;
  } else {
    targ_loci = (int)end_loci;
    if (targ_loci < min_d || (uint32_t)targ_loci >= targ_len) return 0;
  }
  uint32_t sp, ep;
  if (b3) {
    sp = (uint32_t)(targ_loci + min_d);
    if (sp + (uint32_t)L >= targ_len) return 0;
    ep = (uint32_t)(targ_loci + max_d);
  } else {
    sp = end_loci < (uint32_t)max_d ? 0 : end_loci - (uint32_t)max_d;
    ep = end_loci - (uint32_t)min_d;
  }
  const int nw = (L + 31) >> 5;
  unsigned long long best = ~0ull;  // (mm << 40) | order
  if (ep - sp >= 1000) {
    // ---- suffix-array seeded: exact core hits inside the window (no cap, as the reference)
    int match_len = L - 1;
    const int core_subs = pe.rescue_core_subs_p1 > 0 ? pe.rescue_core_subs_p1 - 1 : P.max_subs;  // -6 runs: m_MaxSubs, Aligner.cpp:3256
    int max_tot_mm = core_subs == 0 ? 0 : max(1, (match_len * core_subs + 50) / 100);
    if (max_tot_mm > 63) max_tot_mm = 63;
    int core_len = max(P.min_core_len, L / (P.mmd == 1 ? max_tot_mm + 1 : max_tot_mm + 2));
    int core_delta = max(L / P.slides_per100 - 1, core_len);  // sic: per-100bp value (Aligner.cpp:3266)
    Grp<32> g;
    g.s2[0] = c.rd2; g.s2[1] = c.rd2; g.sx[0] = c.rdx; g.sx[1] = c.rdx;
    g.L = L; g.hasN = has_n; g.gl = c.lane; g.gmask = 0xffffffffu; g.gshift = 0;
    unsigned long long order = 0;
    for (int co = 0; co + core_len <= L; co += core_delta) {
      uint64_t first = 0, cnt = 0;
      if (c.lane == 0) locate_core<32>(I, g, 0, co, core_len, first, cnt);
      first = __shfl_sync(0xffffffffu, first, 0);
      cnt = __shfl_sync(0xffffffffu, cnt, 0);
      for (uint64_t j0 = 0; j0 < cnt; j0 += 32) {
        uint64_t j = j0 + (uint64_t)c.lane;
        if (j < cnt) {
          uint64_t loci = sa_get(I, first + j);
          int e2 = find_entry(I, loci);
          if (e2 >= 0 && (uint32_t)e2 == ei) {
            uint32_t hl = (uint32_t)(loci - cs);
            if (hl >= sp && hl <= ep && (uint32_t)co <= hl && (hl + (uint32_t)L - (uint32_t)co) < targ_len) {
              ATFull at;
              for (int i = 0; i < nw; ++i) at.push(mm_word_global(I, c, cs + hl - co, i, L), i, L);
              int mm = at.result(L, max_allowed);
              if (mm >= 0) {
                unsigned long long key = ((unsigned long long)mm << 40) | (order + j);
                if (key < best) { best = key; out_loci = hl - (uint32_t)co; }
              }
            }
          }
        }
      }
      order += cnt;
    }
  } else {
    // ---- linear scan with the window staged in shared memory
    const uint64_t g0 = cs + sp;                 // first base of the window (concatenation offset)
    const uint64_t w0 = g0 >> 5;                 // first staged 2-bit word
    const uint64_t x0 = g0 >> 6;                 // first staged flag word
    const uint64_t g_last = cs + ep + (uint64_t)L;  // one past the last base any locus touches
    const int n2 = (int)(((g_last + 31) >> 5) - w0) + 1;
    const int nx = (int)(((g_last + 63) >> 6) - x0) + 1;
    const uint64_t lim2 = (I.n + 31) >> 5, limx = (I.n + 63) >> 6;
    for (int i = c.lane; i < n2; i += 32) {
      uint64_t wi = w0 + i;
      uint64_t v2 = wi < lim2 + 2 ? __ldg(I.g2 + wi) : 0ull;
      uint64_t b2 = wi << 5;
      if (b2 + 32 > I.n) v2 |= (b2 >= I.n) ? ~0ull : (~0ull << (2 * (unsigned)(I.n - b2)));  // EOS code past the end
      c.win2[i] = v2;
    }
    for (int i = c.lane; i < nx; i += 32) {
      uint64_t wi = x0 + i;
      uint64_t v = wi < limx + 1 ? __ldg(I.gx + wi) : 0ull;
      // past the end of the concatenation behaves as EOS: flagged (codes there are 0, never equal-and-unflagged)
      uint64_t base = wi << 6;
      if (base + 64 > I.n) v |= (base >= I.n) ? ~0ull : (~0ull << (unsigned)(I.n - base));
      c.winx[i] = v;
    }
    __syncwarp();
    const int off2 = (int)(g0 - (w0 << 5)), offx = (int)(g0 - (x0 << 6));
    for (uint32_t b = 0; b <= ep - sp; b += 32) {
      uint32_t d = b + (uint32_t)c.lane;
      if (d <= ep - sp) {
        ATFull at;
        for (int i = 0; i < nw; ++i) {
          uint64_t gw = sm_word2(c.win2, off2 + (int)d + 32 * i);
          uint32_t gx32 = sm_bits(c.winx, offx + (int)d + 32 * i);
          uint64_t x = c.rd2[i] ^ gw;
          uint32_t m = compress_even(x | (x >> 1)) | (gx32 ^ c.rdx[i]);
          int rem = L - 32 * i;
          if (rem < 32) m &= (1u << rem) - 1;
          at.push(m, i, L);
        }
        int mm = at.result(L, max_allowed);
        if (mm >= 0) {
          unsigned long long key = ((unsigned long long)mm << 40) | d;
          if (key < best) { best = key; out_loci = sp + d; }
        }
      }
    }
    __syncwarp();
  }
  // first locus (in processing order) with the fewest mismatches; must beat MaxAllowedMM+1
  unsigned long long bmin = best;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, bmin, o);
    bmin = t < bmin ? t : bmin;
  }
  if (bmin == ~0ull) return 0;
  int src = __ffs(__ballot_sync(0xffffffffu, best == bmin)) - 1;
  out_loci = __shfl_sync(0xffffffffu, out_loci, src);
  out_mm = (int)(bmin >> 40);
  return out_mm <= max_allowed ? 1 : 0;
}

__global__ void __launch_bounds__(kRescueThreads) orphan_rescue_kernel(
    DevIndex I, KParams P, bkx_pe_params pe, bkx_read_result* __restrict__ res, const uint32_t* __restrict__ orphan_list,
    const unsigned int* __restrict__ n_orphans, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offs,
    int Lmax, bkx_pe_stats* __restrict__ stats, uint32_t* __restrict__ len_dist, unsigned int* __restrict__ cursor,
    const uint8_t* __restrict__ keep) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int pb[8];
  enum { UNAL = 0, ACCP = 1, ACCSE = 2, PPAIRED = 3, PUNP = 4, FILT = 5, UNDER = 6, OVER = 7 };
  if (threadIdx.x < 8) pb[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint64_t* base = (uint64_t*)(smem_raw + rescue_warp_bytes(Lmax) * wib);
  RescueCtx c;
  c.rd2 = base;
  c.rdx = (uint32_t*)(base + rescue_rw(Lmax));
  c.win2 = base + rescue_rw(Lmax) + (rescue_rw(Lmax) + 1) / 2;
  c.winx = c.win2 + rescue_ww(Lmax);
  c.lane = lane;
  const unsigned int total = *n_orphans;
  const int mode = pe.pe_proc;
  for (;;) {
    unsigned int q = 0;
    if (lane == 0) q = atomicAdd(cursor, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= total) break;
    const uint32_t i = orphan_list[q];
    bkx_read_result f = res[2 * i], r = res[2 * i + 1];
    const bool f_un = f.nar == BKX_NAR_NS || f.nar == BKX_NAR_NOHIT || f.nar == BKX_NAR_UNALIGNED;
    const bool r_un = r.nar == BKX_NAR_NS || r.nar == BKX_NAR_NOHIT || r.nar == BKX_NAR_UNALIGNED;
    bool paired = false;
    for (int side = 0; side < 2 && !paired; ++side) {
      // side 0: 5' end is the anchor, rescue PE2 (Aligner.cpp:3222-3310); side 1: 3' anchor, rescue PE1 (:3321-3410)
      const bkx_read_result& anc = side == 0 ? f : r;
      if (!(anc.num_hits == 1 && !(side == 0 ? r_un : f_un))) continue;
      if (keep && !__ldg(keep + anc.chrom_id)) {
        // anchor on a filtered chromosome: no recovery; the reference marks the 5' end in both arms (Aligner.cpp:3296-3302,
        // 3411-3417 -- the second one tests the 3' end and writes the 5' end)
        if (anc.nar == BKX_NAR_ACCEPTED) { f.num_hits = 0; f.low_hit_instances = 0; f.nar = BKX_NAR_CHROMFILT; }
        continue;
      }
      bool b3 = anc.strand == '+';
      bool anti;
      if (side == 0) anti = pe.pair_strand ? (anc.strand != '+') : (anc.strand == '+');
      else { anti = anc.strand == '+'; if (pe.pair_strand) { b3 = !b3; anti = !anti; } }
      if (pe.circularised) b3 = !b3;
      const uint32_t os = anc.match_loci, oe = anc.match_loci + anc.match_len - 1;
      const uint32_t mate = 2 * i + (side == 0 ? 1 : 0);
      const uint64_t o0 = __ldg(offs + mate);
      const int L = (int)(__ldg(offs + mate + 1) - o0);
      if (L < 1 || L > kRescueMaxLen || L > Lmax) continue;
      // pack the mate in the orientation it is expected to align in
      const uint8_t* rd = bases + o0;
      const int words = (L + 31) >> 5;
      int nN = 0;
      for (int w = 0; w <= words; ++w) {
        int p = w * 32 + lane;
        unsigned code = 0, isn = 0;
        if (p < L) {
          unsigned b = __ldg(rd + (anti ? (L - 1 - p) : p)) & 0x07;
          isn = (b >= 4);
          code = (b < 4) ? (anti ? 3 - b : b) : 0;
        }
        unsigned b0 = __ballot_sync(0xffffffffu, code & 1), b1 = __ballot_sync(0xffffffffu, code & 2);
        unsigned bn = __ballot_sync(0xffffffffu, isn);
        if (lane == 0) { c.rd2[w] = spread32(b0) | (spread32(b1) << 1); c.rdx[w] = bn; }
        nN += __popc(bn);
      }
      __syncwarp();
      uint32_t hl = 0;
      int hmm = 0;
      int rs = align_paired_read(I, P, pe, c, nN > 0, b3, anc.chrom_id, os, oe, L, hl, hmm);
      int frag = 0;
      const uint8_t hstrand = anti ? '-' : '+';
      if (rs == 1) {
        if (side == 0) frag = pe_insert_size(pe, anc.strand, os, oe, hstrand, hl, hl + (uint32_t)L - 1);
        else frag = pe_insert_size(pe, hstrand, hl, hl + (uint32_t)L - 1, anc.strand, os, oe);
        if (frag <= 0) rs = 0;
      }
      if (rs == 1) {
        bkx_read_result& m = side == 0 ? r : f;
        m.strand = hstrand; m.chrom_id = anc.chrom_id; m.match_loci = hl; m.match_len = (uint16_t)L;
        m.mismatches = (uint8_t)hmm; m.num_hits = 1; m.low_mm = (int8_t)hmm; m.low_hit_instances = 1;
        f.flags |= BKX_FLG_PE_ALIGNED; r.flags |= BKX_FLG_PE_ALIGNED;
        m.flags |= BKX_FLG_PE_RECOVERED;
        f.nar = r.nar = BKX_NAR_ACCEPTED;
        if (lane == 0) {
          if (len_dist) atomicAdd(len_dist + frag, 1u);
          atomicAdd(&pb[ACCP], 1u);
          atomicAdd(&pb[PPAIRED], 1u);
        }
        paired = true;
      }
    }
    if (!paired) {  // Aligner.cpp:3421-3477
      if (lane == 0) {
        if (f.nar == BKX_NAR_CHROMFILT || r.nar == BKX_NAR_CHROMFILT) atomicAdd(&pb[FILT], 1u);
        if (f.nar == BKX_NAR_PEINSERTMIN || r.nar == BKX_NAR_PEINSERTMIN) atomicAdd(&pb[UNDER], 1u);
        if (f.nar == BKX_NAR_PEINSERTMAX || r.nar == BKX_NAR_PEINSERTMAX) atomicAdd(&pb[OVER], 1u);
      }
      if (mode != BKX_PE_ORPHAN_SE) {
        f.num_hits = r.num_hits = 0;
        f.low_hit_instances = r.low_hit_instances = 0;
        if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PENOHIT;
        if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PENOHIT;
      } else {
        if (f.num_hits != 1 || (keep && !__ldg(keep + f.chrom_id))) { f.num_hits = 0; f.low_hit_instances = 0; if (f.nar == BKX_NAR_ACCEPTED) f.nar = BKX_NAR_PEUNALIGN; }
        else { f.nar = BKX_NAR_ACCEPTED; if (lane == 0) atomicAdd(&pb[ACCSE], 1u); }
        if (r.num_hits != 1 || (keep && !__ldg(keep + r.chrom_id))) { r.num_hits = 0; r.low_hit_instances = 0; if (r.nar == BKX_NAR_ACCEPTED) r.nar = BKX_NAR_PEUNALIGN; }
        else { r.nar = BKX_NAR_ACCEPTED; if (lane == 0) atomicAdd(&pb[ACCSE], 1u); }
      }
    }
    if (lane == 0) { res[2 * i] = f; res[2 * i + 1] = r; }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0 && stats) {
    atomicAdd((unsigned long long*)&stats->accepted_num_paired, (unsigned long long)pb[ACCP]);
    atomicAdd((unsigned long long*)&stats->accepted_num_se, (unsigned long long)pb[ACCSE]);
    atomicAdd((unsigned long long*)&stats->partner_paired, (unsigned long long)pb[PPAIRED]);
    atomicAdd((unsigned long long*)&stats->num_filtered_by_chrom, (unsigned long long)pb[FILT]);
    atomicAdd((unsigned long long*)&stats->under_len_pairs, (unsigned long long)pb[UNDER]);
    atomicAdd((unsigned long long*)&stats->over_len_pairs, (unsigned long long)pb[OVER]);
  }
}

cudaError_t launch_rescue(const DevIndex& I, const KParams& P, const bkx_pe_params& pe, bkx_read_result* res,
                          const uint32_t* orphan_list, const unsigned int* n_orphans, const uint8_t* bases,
                          const uint64_t* offs, int Lmax, bkx_pe_stats* stats, uint32_t* len_dist, unsigned int* cursor,
                          const uint8_t* keep, cudaStream_t st) {
  size_t smem = rescue_warp_bytes(Lmax) * kRescueWarps;
  cudaError_t e = ensure_smem(orphan_rescue_kernel, smem, g_smem_rescue);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  orphan_rescue_kernel<<<148 * 4, kRescueThreads, smem, st>>>(I, P, pe, res, orphan_list, n_orphans, bases, offs, Lmax,
                                                               stats, len_dist, cursor, keep);
  return cudaGetLastError();
}

}  // namespace bkx
